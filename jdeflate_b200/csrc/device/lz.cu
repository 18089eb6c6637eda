/*
 * lz.cu -- LZ77 stage of the deflate pipeline for sm_100a.
 *
 * Replaces the reference's serial parsers compress1 / compress2 with their
 * match finders getmatch1 / getmatch2 and hash maintenance skipbytes1/2,
 * slidehash (src/deflator.c:2335-2973, 1899-1911).
 *
 * Two kernels:
 *
 *  chain_kernel   builds, for every input position, the distance to the
 *                 previous position with the same 4-byte hash -- the links the
 *                 reference's insert-every-position policy produces in
 *                 mchain[] (the chain content does not depend on the parse,
 *                 only on the data).  Hash: big-endian 4 bytes * 0x1e35a7bd
 *                 (gethead/gethash, src/deflator.c:1930-1947), top 14 bits
 *                 (the reference keeps 16; measured cost of 14: +1.5 % chain
 *                 steps, +0.04 % size).  The head-table update is a serial
 *                 dependency, so throughput = tables resident per SM / step
 *                 latency: a 32 KiB table (2^14 x u16) lets six CTAs share an
 *                 SM.  A CTA is a producer warp (coalesced prefetched loads,
 *                 hashing, link write-back) and a consumer warp (head table),
 *                 32 positions per step, duplicates inside a step detected by
 *                 a store / read-back and resolved with match_any over the
 *                 lanes involved only.
 *
 *  lz_kernel      one CTA per 16 KiB segment.  The 32 KiB of history plus the
 *                 segment (bytes and chain links) are staged in shared memory;
 *                 every position is searched (bounded by the level's max
 *                 chain, early exit at `nice`, the reference's
 *                 strbgn[len]==pmatch[len] pre-filter) -- the search is
 *                 position-parallel because on the compression side the
 *                 history is the input itself.  The lazy / greedy selection
 *                 (good length, the reference's offset-aware accept rule
 *                 src/deflator.c:2860-2879) is evaluated per position, and the
 *                 one truly serial step -- following the chosen tokens from
 *                 the segment start -- is done by 128 speculative walkers
 *                 whose paths are stitched exactly (paths re-converge within a
 *                 few tokens).  Tokens are compacted with a block scan and
 *                 written coalesced; symbol histograms are accumulated with
 *                 shared-memory atomics.
 *
 * Algorithmic traffic of the stage: N bytes read.  Implementation traffic per
 * input byte: 2 B links written + (1+2)*3 B staged per segment (history halo)
 * + <= 4 B tokens written.
 */
#include "deflate.cuh"
#include <stdlib.h>

#ifndef HASH_BITS
#define HASH_BITS      14
#endif
#define HASH_MUL       0x1e35a7bdu

/* ---------------------------------------------------------------------------
 * chain_kernel
 * ------------------------------------------------------------------------- */

/*
 * Work item r covers positions [r*range, (r+1)*range) of the batch; ranges
 * never straddle chunks (range divides the chunk size).  A range that does not
 * start a chunk first replays the preceding 32 KiB without emitting links.
 *
 * The head-table update is a serial dependency (group g+1 must see the heads
 * group g stored), so one warp runs it with nothing else on its plate and a
 * second warp feeds it: warp 0 ("producer") streams the input with coalesced,
 * prefetched 128-byte loads, hashes 128 positions per round into a shared
 * ring and writes the finished links of the round before last to HBM; warp 1
 * ("consumer") turns hashes into links.  One __syncthreads per round.
 */
#define CH_THREADS   64
#define CH_BLOCK     128u                  /* positions per round */
#define CH_DUMMY     (1u << HASH_BITS)     /* 32 private slots for positions that are not hashed */

struct ChainSmem {
	uint16_t head[(1u << HASH_BITS) + 32];
	uint16_t hash[2][CH_BLOCK];
	uint16_t dist[2][CH_BLOCK];
};

/* producer: hash block `blk` (lane i holds word i of the block, `wnext` word i of the
 * next one); every lane assembles the 4 bytes of its position with two shuffles */
static __device__ __forceinline__ void
chain_hash_block(ChainSmem& S, uint32_t blk, uint32_t wcur, uint32_t wnext, uint32_t hashable, unsigned lane)
{
	const uint32_t rel0 = blk * CH_BLOCK;
#pragma unroll
	for (uint32_t j = 0; j < 4; j++) {
		const uint32_t idx = 8 * j + (lane >> 2);
		const uint32_t lo = __shfl_sync(JDB_FULL_MASK, wcur, idx);
		uint32_t hi = __shfl_sync(JDB_FULL_MASK, wcur, (idx + 1) & 31);
		if (j == 3) {
			const uint32_t hn = __shfl_sync(JDB_FULL_MASK, wnext, 0);
			if (idx == 31) hi = hn;
		}
		const uint32_t le = __funnelshift_r(lo, hi, (lane & 3u) * 8u);
		const uint32_t be = __byte_perm(le, 0, 0x0123);
		const uint32_t rel = rel0 + 32 * j + lane;
		const uint32_t h = rel < hashable ? (be * HASH_MUL) >> (32 - HASH_BITS) : CH_DUMMY + lane;
		S.hash[blk & 1][32 * j + lane] = (uint16_t) h;
	}
}

/* producer: links of block `blk`, finished by the consumer in the previous round, to HBM */
static __device__ __forceinline__ void
chain_emit_block(ChainSmem& S, uint32_t blk, uint16_t* __restrict__ out, uint32_t emit0, uint32_t span, unsigned lane)
{
	const uint32_t rel0 = blk * CH_BLOCK;
#pragma unroll
	for (uint32_t j = 0; j < 4; j++) {
		const uint32_t rel = rel0 + 32 * j + lane;
		if (rel >= emit0 && rel < span) out[rel] = S.dist[blk & 1][32 * j + lane];
	}
}

/* consumer: hashes of block `blk` -> links, head table update */
static __device__ __forceinline__ void
chain_link_block(ChainSmem& S, uint32_t blk, unsigned lane)
{
	const uint32_t rel0 = blk * CH_BLOCK;
	if (rel0 && (rel0 & (WND - 1)) == 0) {
		/* every 32768 positions retire entries that are out of the window
		 * so 16-bit positions never alias (cf. slidehash) */
		const uint32_t stale = (rel0 + 0x8000u) & 0xffffu;
		for (uint32_t i = lane; i < (1u << HASH_BITS); i += 32) {
			uint32_t d = (rel0 - S.head[i]) & 0xffffu;
			/* (d == 0: position rel0 itself is not in the table yet -- this is the "empty" value
			 * of the start, 0x8000, which from here on would read as a link to position 32768) */
			if (d >= WND || d == 0) S.head[i] = (uint16_t) stale;
		}
		__syncwarp();
	}
	uint32_t hh[4];
#pragma unroll
	for (uint32_t j = 0; j < 4; j++) hh[j] = S.hash[blk & 1][32 * j + lane];
#pragma unroll
	for (uint32_t j = 0; j < 4; j++) {
		const uint32_t rel = rel0 + 32 * j + lane;
		const uint32_t slot = hh[j];
		/* Two lanes with the same hash in one group of 32 are not rare in text
		 * (short words, runs).  Fast path: everybody reads the old head, everybody
		 * stores its own position, and a read-back tells whether any store lost,
		 * i.e. whether duplicates exist; only then the exact resolution runs. */
		const uint32_t old = S.head[slot];
		__syncwarp();
		S.head[slot] = (uint16_t) rel;
		__syncwarp();
		const uint32_t chk = S.head[slot];
		const bool lost = chk != (rel & 0xffffu);
		const uint32_t d = (rel - old) & 0xffffu;
		uint32_t dist = (d < WND && d <= rel) ? d : 0;
		const unsigned lostmask = __ballot_sync(JDB_FULL_MASK, lost);
		if (lostmask) {
			/* lanes involved: the losers and the winners they lost to (the read-back
			 * names the winner's position).  match_any costs per distinct value, so
			 * it runs over the involved lanes only. */
			const unsigned involved = lostmask |
				__reduce_or_sync(JDB_FULL_MASK, lost ? 1u << ((chk - rel0) & 31u) : 0u);
			if ((involved >> lane) & 1u) {
				const unsigned same = __match_any_sync(involved, slot);
				const unsigned lower = same & ((1u << lane) - 1u);
				if (lower) dist = lane - (31 - __clz(lower));
				if ((same >> lane) == 1u) S.head[slot] = (uint16_t) rel;     /* highest lane of its group */
			}
			__syncwarp();
		}
		if (slot >= CH_DUMMY) dist = 0;
		S.dist[blk & 1][32 * j + lane] = (uint16_t) dist;
	}
}

__global__ void __launch_bounds__(CH_THREADS)
chain_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes, uint32_t range,
             const uint32_t* __restrict__ chunk_len, uint16_t* __restrict__ prev, uint16_t* __restrict__ heads)
{
	JDB_DYN_SMEM(smem_raw);
	ChainSmem& S = *(ChainSmem*) smem_raw;
	const unsigned lane = threadIdx.x & 31u;
	const bool producer = threadIdx.x < 32;

	const uint64_t r0 = (uint64_t) blockIdx.x * range;
	if (r0 >= n) return;
	const uint64_t chunk0 = r0 / chunk_bytes * chunk_bytes;
	const uint64_t chunk1 = chunk_end(chunk_len, chunk0, chunk_bytes, n);
	if (r0 >= chunk1) return;                    /* ragged chunk: nothing in this range */
	uint64_t r1 = r0 + range;
	if (r1 > chunk1) r1 = chunk1;
	/* warm-up start, multiple of SEG.  With `heads` there is no warm-up: every range leaves its
	 * head table behind and chain_fix_kernel links the first occurrences of a range to the last
	 * ones of the range before it */
	const uint64_t start = heads ? r0 : r0 >= chunk0 + WND ? r0 - WND : chunk0;

	/* positions are handled relative to `start`; an entry holds the low 16
	 * bits, "empty" is anything that decodes to a distance >= 32768 */
	for (uint32_t i = threadIdx.x; i < (1u << HASH_BITS) + 32; i += CH_THREADS) S.head[i] = 0x8000u;

	const uint32_t* words = (const uint32_t*) (in + start);
	const uint64_t nwords = (n - start + 3) / 4;                 /* words that start before `n` */
	const uint32_t span = (uint32_t) (r1 - start);
	const uint32_t nblocks = (span + CH_BLOCK - 1) / CH_BLOCK;
	const uint32_t hashable = chunk1 - start >= 4 ? (uint32_t) (chunk1 - start - 3) : 0;   /* rel < hashable */
	const uint32_t emit0 = (uint32_t) (r0 - start);

#define LOADW(blk) ((uint64_t) (blk) * 32 + lane < nwords ? __ldg(words + (uint64_t) (blk) * 32 + lane) : 0u)
	/* four blocks in flight, in four NAMED registers: the round loop is unrolled by
	 * four so that no register is ever copied while its load is outstanding (a
	 * rotating w0 = w1 ... would wait for the newest load every round).  Eight in
	 * flight: 2.35 instead of 2.24 ms per 256 MiB (round 2). */
	uint32_t wa = 0, wb = 0, wc = 0, wd = 0;
	if (producer) { wa = LOADW(0); wb = LOADW(1); wc = LOADW(2); wd = LOADW(3); }
	__syncthreads();

#define CH_ROUND(T, WCUR, WNEXT) \
	do { \
		const uint32_t t_ = (T); \
		if (t_ < nblocks + 2) { \
			if (producer) { \
				if (t_ < nblocks) chain_hash_block(S, t_, WCUR, WNEXT, hashable, lane); \
				if (t_ >= 2) chain_emit_block(S, t_ - 2, prev + start, emit0, span, lane); \
				WCUR = LOADW(t_ + 4); \
			} else if (t_ >= 1 && t_ <= nblocks) { \
				chain_link_block(S, t_ - 1, lane); \
			} \
		} \
		__syncthreads(); \
	} while (0)

	for (uint32_t t = 0; t < nblocks + 2; t += 4) {
		CH_ROUND(t + 0, wa, wb);
		CH_ROUND(t + 1, wb, wc);
		CH_ROUND(t + 2, wc, wd);
		CH_ROUND(t + 3, wd, wa);
	}
#undef CH_ROUND
#undef LOADW
	if (heads) {
		/* distance from the end of the range back to the last position of every hash (0: none
		 * inside the window); the last round ended with a barrier */
		uint16_t* const hout = heads + (uint64_t) blockIdx.x * (1u << HASH_BITS);
		for (uint32_t i = threadIdx.x; i < (1u << HASH_BITS); i += CH_THREADS) {
			const uint32_t d = (span - S.head[i]) & 0xffffu;
			hout[i] = (uint16_t) (d < WND ? d : 0u);
		}
	}
}

/*
 * Second step of the chain build without warm-up: a position of range r that found no earlier
 * position with its hash inside the range (link 0) is linked to the last such position of range
 * r - 1, when that lies inside the window.  Only the first 32 KiB of a range can reach back.
 */
__global__ void __launch_bounds__(256)
chain_fix_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes, uint32_t range,
                 const uint32_t* __restrict__ chunk_len, uint16_t* __restrict__ prev, const uint16_t* __restrict__ heads)
{
	const uint64_t r0 = (uint64_t) blockIdx.x * range;
	if (r0 >= n) return;
	const uint64_t chunk0 = r0 / chunk_bytes * chunk_bytes;
	if (r0 == chunk0) return;                    /* nothing in front of a chunk */
	const uint64_t chunk1 = chunk_end(chunk_len, chunk0, chunk_bytes, n);
	if (r0 >= chunk1) return;
	uint64_t lim = r0 + WND < r0 + range ? r0 + WND : r0 + range;
	if (chunk1 < 4) return;
	if (lim > chunk1 - 3) lim = chunk1 - 3;      /* positions that have four bytes to hash */
	const uint16_t* const H = heads + (uint64_t) (blockIdx.x - 1) * (1u << HASH_BITS);
	/* eight positions per trip: their links (one 16-byte load; r0 is a multiple of SEG) and the
	 * 11 bytes that hold their hashes (12 bytes from an 8-byte aligned address) -- no load depends
	 * on a branch, so everything a trip needs is in flight at once */
	for (uint64_t p8 = r0 + 8u * threadIdx.x; p8 < lim; p8 += 8u * 256u) {
		const uint4 v = *(const uint4*) (prev + p8);
		const uint2 b01 = *(const uint2*) (in + p8);
		const uint32_t b2 = *(const uint32_t*) (in + p8 + 8);
		const uint32_t lk[4] = { v.x, v.y, v.z, v.w };
		const uint32_t by[4] = { b01.x, b01.y, b2, 0u };
		uint32_t dend[8];
#pragma unroll
		for (uint32_t k = 0; k < 8; k++) {
			const bool z = ((lk[k >> 1] >> (16u * (k & 1u))) & 0xffffu) == 0 && p8 + k < lim;
			const uint32_t le = __funnelshift_r(by[k >> 2], by[(k >> 2) + 1], 8u * (k & 3u));
			const uint32_t be = __byte_perm(le, 0, 0x0123);
			dend[k] = z ? H[(be * HASH_MUL) >> (32 - HASH_BITS)] : 0u;
		}
#pragma unroll
		for (uint32_t k = 0; k < 8; k++) {
			const uint32_t d = (uint32_t) (p8 + k - r0) + dend[k];
			if (dend[k] && d < WND) prev[p8 + k] = (uint16_t) d;
		}
	}
}

/* ---------------------------------------------------------------------------
 * lz_kernel
 * ------------------------------------------------------------------------- */

/*
 * A CTA handles two consecutive segments (8 KiB of positions each) that share one staged
 * window (the 32 KiB of history before the first plus both segments: bytes and links).  The
 * two halves of the CTA ("groups", 512 threads each, their own named barrier) run the phases
 * below independently of each other: most phases leave either the issue slots or the
 * shared-memory pipe idle (serial walkers, dependent loads), and the other group fills them.
 */
#define LZ_THREADS   1024
#define LZ_GROUPS    2
#define LZ_GTHREADS  (LZ_THREADS / LZ_GROUPS)      /* 512 */
#ifndef LZ_WALK_STEPS
#define LZ_WALK_STEPS 2
#endif
#define PER_THREAD   (SEG / LZ_GTHREADS)           /* 16 */
#ifndef WALK_BLOCK
#define WALK_BLOCK   64u       /* 128: 2-3 % slower (the walkers are a serial step; round 2) */
#endif
#define WALKERS      (SEG / WALK_BLOCK)            /* 128 */
#define WORDS_PER_WB (WALK_BLOCK / 32)             /* 2 */
#define DATA_BYTES   (WND + LZ_GROUPS * SEG + 320) /* history + segments + look-ahead/guard */
#define NOPOS        0xffffffffu

#ifdef JDB_SIMT_EMU
/* emulator-only instrumentation (tools/emu_lz_explore.py): chain steps per position */
extern "C" { uint32_t* jdb_emu_lz_iters = 0; uint16_t* jdb_emu_lz_steps = 0; }
#define LZ_STAT(x) x
#else
#define LZ_STAT(x)
#endif

/* experiment builds only (-DLZ_PROF): cycles of thread 0 per phase, summed over all CTAs,
 * printed when the process exits */
#if defined(LZ_PROF) && !defined(JDB_SIMT_EMU)
#include <stdio.h>
enum { LZP_STAGE, LZP_PASS1, LZP_PARSE, LZP_NEED, LZP_SEARCH, LZP_SHORT3, LZP_FINALPARSE, LZP_EMIT, LZP_PA, LZP_PB, LZP_PC, LZP_PD, LZP_N };
__device__ unsigned long long g_lz_prof[LZP_N + 1];
__device__ unsigned long long g_lz_cnt[8];
#define LZ_PROF_COUNTERS() uint32_t lzc_[6] = {0, 0, 0, 0, 0, 0}
#define LZ_PROF_COUNT(i, v) lzc_[i] += (uint32_t) (v)
#define LZ_PROF_FLUSH() do { for (int i_ = 0; i_ < 6; i_++) { uint32_t v_ = lzc_[i_]; if (i_ == 3) v_ = __reduce_add_sync(JDB_FULL_MASK, v_); if (lane == 0) atomicAdd(&g_lz_cnt[i_], (unsigned long long) v_); } } while (0)
#define LZ_PROF_INIT() long long lzp_t = clock64()
#define LZ_PROF_PARAM , long long& lzp_t
#define LZ_PROF_ARG , lzp_t
#define LZ_PROF_MARK(i) do { if (tid == 0) { const long long t_ = clock64(); atomicAdd(&g_lz_prof[i], (unsigned long long) (t_ - lzp_t)); lzp_t = t_; } } while (0)
static void lz_prof_dump()
{
	static const char* names[LZP_N] = { "stage", "pass1", "parse", "need", "search", "short3", "finalparse", "emit", "parse:take", "parse:walk", "parse:entry", "parse:fix" };
	unsigned long long h[LZP_N + 1];
	if (cudaMemcpyFromSymbol(h, g_lz_prof, sizeof(h)) != cudaSuccess) return;
	unsigned long long tot = 0;
	for (int i = 0; i < LZP_N; i++) tot += h[i];
	fprintf(stderr, "LZ_PROF segments %llu, cycles/segment %.0f:", h[LZP_N], h[LZP_N] ? (double) tot / h[LZP_N] : 0.0);
	for (int i = 0; i < LZP_N; i++) fprintf(stderr, " %s %.0f (%.1f%%)", names[i], h[LZP_N] ? (double) h[i] / h[LZP_N] : 0.0, tot ? 100.0 * h[i] / tot : 0.0);
	fprintf(stderr, "\n");
	unsigned long long c[8];
	if (cudaMemcpyFromSymbol(c, g_lz_cnt, sizeof(c)) != cudaSuccess || !h[LZP_N]) return;
	fprintf(stderr, "LZ_PROF per segment: warp-steps %.0f (walking lanes %.1f), extend phases %.0f (lanes %.1f), batches %.0f (positions %.1f)\n",
	        (double) c[0] / h[LZP_N], c[0] ? (double) c[1] / c[0] : 0.0, (double) c[2] / h[LZP_N], c[2] ? (double) c[3] / c[2] : 0.0,
	        (double) c[4] / h[LZP_N], c[4] ? (double) c[5] / c[4] : 0.0);
}
#else
#define LZ_PROF_INIT()
#define LZ_PROF_PARAM
#define LZ_PROF_ARG
#define LZ_PROF_MARK(i)
#define LZ_PROF_COUNTERS()
#define LZ_PROF_COUNT(i, v)
#define LZ_PROF_FLUSH()
#endif

struct LzParams {
	uint32_t good, nice, chain, lazy;
	uint32_t short3;                 /* probe short distances for 3-byte matches */
	uint32_t skip_segs;              /* leading segments that are preset dictionary: history only */
	uint32_t rounds;                 /* 0: search every position; k: k rounds of "search where a tentative parse goes" */
	uint32_t hist_min;               /* first byte of the batch a match of chunk 0 may reach (dictionary start) */
	uint32_t patience;               /* chain steps still allowed once a match has been found */
	uint32_t tlazy;                  /* the first tentative parse applies the lazy rule, too */
	uint32_t succ;                   /* successors of path positions are searched as well (the lazy rule reads them) */
	uint32_t prewalk;                /* links followed in pass 1 to tell how long a position's chain is (<= 8) */
	uint32_t skip_div;               /* no search at all when fewer than 1/skip_div of the positions matched in pass 1 */
	uint32_t r2min;                  /* no second round when the first improved fewer than 1/r2min of the positions (0: always two) */
};

struct LzGroup {
	uint32_t m[SEG];                 /* per position result: len << 16 | dist, 0 = none */
	/* parse state */
	uint32_t spec[SEG / 32];         /* positions on the speculative paths    */
	uint32_t fix[SEG / 32];          /* positions added by the stitching pass; during a search: successor-only positions */
	uint32_t take[SEG / 32];         /* positions whose match a parser arriving there takes */
	uint32_t cand[SEG / 32];         /* pass 1: positions with a chain worth walking */
	uint32_t cls[2][SEG / 32];       /* pass 1: ... and, bit-sliced, the class 0..3 = expected length of the walk */
	uint32_t need[SEG / 32];         /* positions that get the full search now */
	uint32_t done[SEG / 32];         /* positions that were searched in an earlier round */
	uint16_t land[WALKERS];          /* where each walker left its block      */
	uint16_t merge[WALKERS];
	uint16_t entry[WALKERS];         /* where the true path enters the block, then where it leaves it */
	uint32_t warp_sum[LZ_GTHREADS / 32];
	uint32_t nomatch;                /* positions without a match (3-byte probe switch) */
	uint32_t conflict;               /* first block whose speculative landing turned out wrong */
	uint32_t nmatch1;                /* positions whose first candidate matched */
	uint32_t improved;               /* positions whose match the last search round improved */
	uint32_t pad_;                   /* sizeof(LzGroup) stays a multiple of 8: land[] is read as uint2 */
	uint32_t ccount[2];              /* positions queued per class, 2 x 16 bits each */
	uint32_t next_batch;
};
/* the symbol histogram of the emit phase lives where need[] and done[] were */
static_assert(2 * (SEG / 32) >= NSYM, "histogram overlay");

struct LzSmem {
	uint16_t prev[WND + LZ_GROUPS * SEG];
	uint8_t  data[DATA_BYTES];
	LzGroup  g[LZ_GROUPS];
};

static_assert(sizeof(LzGroup) % 8 == 0, "the second group's arrays are read with 8-byte vector loads");
static_assert(sizeof(LzSmem) <= 232448, "LzSmem must fit the 227 KB of shared memory a CTA can have");

/* barrier of one group (named barrier 1 + group) */
static __device__ __forceinline__ void lz_gsync(uint32_t grp)
{
#ifdef JDB_SIMT_EMU
	simt_named_barrier(1 + grp, LZ_GTHREADS);
#else
	asm volatile("bar.sync %0, %1;" :: "r"(1 + grp), "n"(LZ_GTHREADS) : "memory");
#endif
}

__constant__ uint8_t c_short_dist[12] = { 1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64 };

static __device__ __forceinline__ uint32_t ilog2_u32(uint32_t v) { return 31 - __clz(v); }

/* the reference's lazy accept rule (src/deflator.c:2865-2879): does the match (nlen, ndist) found
 * one position later displace the match (len, dist)?  nlen >= len is the caller's business. */
static __device__ __forceinline__ bool
lazy_displaces(uint32_t len, uint32_t dist, uint32_t nlen, uint32_t ndist)
{
	const int32_t delta = (int32_t) nlen - (int32_t) len;
	if (delta > 4) return true;
	const int32_t l1 = (int32_t) ilog2_u32(dist), l2 = (int32_t) ilog2_u32(ndist);
	return (delta << 2) + (l1 - l2) >= 2;
}

/* class of a chain walk of about `est` steps */
static __device__ __forceinline__ uint32_t lz_class(uint32_t est) { return est <= 2 ? 0u : est <= 6 ? 1u : est <= 24 ? 2u : 3u; }

/* the token decision for a fresh position p (segment coordinates): returns
 * the next fresh position */
static __device__ __forceinline__ uint32_t
next_pos(const LzGroup& S, uint32_t p)
{
	return ((S.take[p >> 5] >> (p & 31u)) & 1u) ? p + (S.m[p] >> 16) : p + 1;
}

/*
 * The parse over the per-position matches in S.m.  The only serial thing about it is
 * following the chosen tokens from the segment start, and paths that start at different
 * positions re-converge within a few tokens, so:
 *   (1) per position: would a parser arriving here take the match (lazy rule)?
 *   (2) one speculative walker per 64 positions follows the decisions from the start of
 *       its block until it leaves the block (land[]);
 *   (3) one thread chains the landings: the true path enters block w where the last
 *       entered block before it landed -- provided the path through that block joined
 *       the block's speculative path, which (4) checks for all blocks in parallel: a
 *       thread per block walks from the entry to the first speculative position;
 *   (5) only when some block was crossed without joining (a match longer than what was
 *       left of the block) one thread redoes the rest serially.
 * Afterwards position p is on the path iff (spec bit && p >= merge[block]) || fix bit.
 * All threads call it.
 */
static __device__ __forceinline__ void
lz_parse_phase(LzGroup& S, const uint32_t grp, const LzParams& prm, const uint32_t tid, const uint32_t seg_len, const bool lazy LZ_PROF_PARAM)
{
	const uint32_t nblk = (seg_len + WALK_BLOCK - 1) / WALK_BLOCK;
	/* ---- per position: would a parser arriving here take the match? ----
	 * The answers go into a bitmap (one ballot per 32 positions) that lets the
	 * walkers below cross literal runs in one step. */
	for (uint32_t k = 0; k < PER_THREAD; k++) {
		if (k * LZ_GTHREADS >= seg_len) break;
		const uint32_t p = tid + k * LZ_GTHREADS;
		const uint32_t v = p < seg_len ? S.m[p] : 0u;
		const uint32_t len = v >> 16;
		bool take = len != 0;
		if (lazy && take && len < prm.good && p + 1 < seg_len) {
			const uint32_t w = S.m[p + 1];
			const uint32_t nlen = w >> 16;
			if (nlen >= len && lazy_displaces(len, v & 0xffffu, nlen, w & 0xffffu)) take = false;
		}
		const unsigned tb = __ballot_sync(JDB_FULL_MASK, take);
		if ((tid & 31u) == 0) { S.take[p >> 5] = tb; S.spec[p >> 5] = 0; S.fix[p >> 5] = 0; }
	}
	if (tid == 0) S.conflict = NOPOS;
	lz_gsync(grp);
	LZ_PROF_MARK(LZP_PA);

	/* ---- speculative walkers: one per 64 positions ---- */
	if (tid < nblk) {
		const uint32_t b0 = tid * WALK_BLOCK, b1 = b0 + WALK_BLOCK;
		const uint32_t lim = b1 < seg_len ? b1 : seg_len;
		uint32_t p = b0;
		/* the TAKE bitmap word of the current 32 positions lives in a register: a
		 * literal run inside the word is crossed in one step without touching
		 * m[], a taken match costs one m[] load for its length; the path bits of
		 * the word are collected in a register too (words of a block belong to
		 * one walker: plain stores) */
		uint32_t curw = p >> 5, tw = S.take[curw], sw = 0;
		while (p < lim) {
			const uint32_t w = p >> 5;
			if (w != curw) {
				S.spec[curw] = sw;
				curw = w;
				tw = S.take[w];
				sw = 0;
			}
			const uint32_t sh = p & 31u;
			const uint32_t bits = tw >> sh;
			if (bits & 1u) {
				sw |= 1u << sh;
				p += S.m[p] >> 16;
			} else {
				uint32_t nlit = bits ? (uint32_t) (__ffs((int) bits) - 1) : 32u - sh;
				if (nlit > lim - p) nlit = lim - p;
				sw |= (nlit >= 32u ? 0xffffffffu : ((1u << nlit) - 1u)) << sh;
				p += nlit;
			}
		}
		S.spec[curw] = sw;
		S.land[tid] = (uint16_t) p;
	}
	lz_gsync(grp);
	LZ_PROF_MARK(LZP_PB);

	/* ---- where the true path enters every block, if every block joins its walker ---- */
	if (tid == 0) {
		uint32_t g = 0;
		for (uint32_t w = 0; w < nblk; w += 4) {
			/* four landings per load; slots past nblk are never read back */
			const uint2 l4 = *(const uint2*) &S.land[w];
			const uint32_t l[4] = { l4.x & 0xffffu, l4.x >> 16, l4.y & 0xffffu, l4.y >> 16 };
			uint32_t e[4];
#pragma unroll
			for (uint32_t u = 0; u < 4; u++) {
				e[u] = g;
				if (g < (w + u + 1) * WALK_BLOCK) g = l[u];
			}
			*(uint2*) &S.entry[w] = make_uint2(e[0] | (e[1] << 16), e[2] | (e[3] << 16));
		}
	}
	lz_gsync(grp);
	LZ_PROF_MARK(LZP_PC);

	/* ---- per block: from the entry to the first speculative position ---- */
	if (tid < nblk) {
		const uint32_t b0 = tid * WALK_BLOCK, b1 = b0 + WALK_BLOCK;
		const uint32_t lim = b1 < seg_len ? b1 : seg_len;
		const uint32_t t = S.entry[tid];
		const uint32_t ld = S.land[tid];
		uint32_t mg, out;
		if (t >= lim) { mg = lim; out = t; }                 /* block jumped over */
		else if (t == b0) { mg = b0; out = ld; }
		else {
			uint32_t p = t;
			while (p < lim && !((S.spec[p >> 5] >> (p & 31)) & 1u)) {
				S.fix[p >> 5] |= 1u << (p & 31);             /* words of this block: no other writer */
				p = next_pos(S, p);
			}
			if (p < lim) { mg = p; out = ld; }
			else {
				mg = lim; out = p;
				if (p != ld) atomicMin(&S.conflict, tid);
			}
		}
		S.merge[tid] = (uint16_t) mg;
		S.entry[tid] = (uint16_t) out;                       /* where the path leaves the block */
	}
	lz_gsync(grp);
	LZ_PROF_MARK(LZP_PD);

	/* ---- rare: a block was crossed without joining its walker; the blocks after the
	 * first such block were entered on a wrong assumption: redo them serially ---- */
	if (S.conflict != NOPOS) {
		if (tid == 0) {
			const uint32_t c = S.conflict;
			uint32_t t = S.entry[c];
			for (uint32_t w = c + 1; w < nblk; w++) {
				const uint32_t b0 = w * WALK_BLOCK, b1 = b0 + WALK_BLOCK;
				const uint32_t lim = b1 < seg_len ? b1 : seg_len;
				for (uint32_t i = 0; i < WORDS_PER_WB; i++) S.fix[w * WORDS_PER_WB + i] = 0;
				if (t >= lim) { S.merge[w] = (uint16_t) lim; continue; }
				if (t == b0) { S.merge[w] = (uint16_t) b0; t = S.land[w]; continue; }
				uint32_t p = t;
				while (p < lim && !((S.spec[p >> 5] >> (p & 31)) & 1u)) {
					S.fix[p >> 5] |= 1u << (p & 31);
					p = next_pos(S, p);
				}
				if (p < lim) { S.merge[w] = (uint16_t) p; t = S.land[w]; }
				else { S.merge[w] = (uint16_t) lim; t = p; }
			}
		}
		lz_gsync(grp);
	}
}

/* the path bits of bitmap word w after lz_parse_phase */
static __device__ __forceinline__ uint32_t
lz_path_word(const LzGroup& S, uint32_t w)
{
	const uint32_t mg = S.merge[w / WORDS_PER_WB];
	uint32_t pb = S.spec[w];
	if (mg >= (w + 1) * 32) pb = 0;
	else if (mg > w * 32) pb &= 0xffffffffu << (mg - w * 32);
	return pb | S.fix[w];
}

__global__ void __launch_bounds__(LZ_THREADS, 1)
lz_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes,
          const uint32_t* __restrict__ chunk_len, const uint16_t* __restrict__ prev, LzParams prm,
          uint32_t* __restrict__ tok, uint32_t* __restrict__ seg_ntok, uint32_t* __restrict__ seg_hist)
{
	JDB_DYN_SMEM(smem_raw);
	LzSmem& W = *(LzSmem*) smem_raw;             /* the staged window, shared by both groups */
	const uint32_t ctid = threadIdx.x;
	const uint32_t grp = ctid / LZ_GTHREADS;
	const uint32_t tid = ctid % LZ_GTHREADS;     /* thread index within the group */
	const uint32_t lane = tid & 31u;
	LzGroup& S = W.g[grp];

	/* the pair of segments of this CTA.  Chunk sizes are multiples of the pair -- both segments lie
	 * in one chunk and share its history -- or exactly one segment (small records, one per slot:
	 * jdb200_deflate_batch): then the pair is two chunks, each group has its own, and all they share
	 * is the staged window. */
	const uint64_t pair0 = (uint64_t) blockIdx.x * (LZ_GROUPS * SEG);
	const bool split = chunk_bytes < LZ_GROUPS * SEG;
	const uint64_t chunkA0 = pair0 / chunk_bytes * chunk_bytes;
	const uint64_t chunkA1 = chunk_end(chunk_len, chunkA0, chunk_bytes, n);
	uint64_t chunkB0 = chunkA0, chunkB1 = chunkA1;
	if (split && pair0 + SEG < n) {
		chunkB0 = pair0 + SEG;
		chunkB1 = chunk_end(chunk_len, chunkB0, chunk_bytes, n);
	}
	/* the end of the data this CTA looks at */
	uint64_t pair1 = pair0 + LZ_GROUPS * SEG;
	if (!split) { if (pair1 > chunkA1) pair1 = chunkA1; }
	else pair1 = chunkB1 > chunkB0 ? chunkB1 : chunkA1;
	if (pair0 >= pair1) return;                  /* ragged chunk(s): empty segment slots (the whole CTA leaves) */
	const uint64_t hist0 = pair0 >= chunkA0 + WND ? pair0 - WND : chunkA0;  /* first staged byte */
	const uint64_t stage1 = split ? pair1 : chunkA1;                        /* staged bytes end here at the latest */
	LZ_PROF_INIT();

	/* ---- stage bytes and links (16-byte vectors; `in` and hist0 are 16-aligned) ---- */
	if (blockIdx.x * LZ_GROUPS + LZ_GROUPS > prm.skip_segs) {
		const uint32_t nbytes = (uint32_t) ((stage1 - hist0) < (uint64_t) DATA_BYTES ? (stage1 - hist0) : DATA_BYTES);
		const uint4* src = (const uint4*) (in + hist0);
		uint4* dst = (uint4*) W.data;
		const uint32_t nv = nbytes / 16;
		for (uint32_t i = ctid; i < nv; i += LZ_THREADS) dst[i] = __ldg(src + i);
		/* tail bytes, then zeros as far as a comparison can look past the data (320 guard) */
		const uint32_t fill_end = nv * 16 + 16 + 320 < DATA_BYTES ? nv * 16 + 16 + 320 : DATA_BYTES;
		for (uint32_t i = nv * 16 + ctid; i < fill_end; i += LZ_THREADS)
			W.data[i] = i < nbytes ? in[hist0 + i] : 0;
		/* links are staged as absolute shared-memory positions (0xffff = none), so a
		 * chain step is one load and one range check */
		const uint32_t nlinks = (uint32_t) (pair1 - hist0);
		const uint4* ps = (const uint4*) (prev + hist0);
		const uint32_t npv = nlinks / 8;
		for (uint32_t i = ctid; i < npv; i += LZ_THREADS) {
			const uint4 v = __ldg(ps + i);
			const uint32_t w[4] = { v.x, v.y, v.z, v.w };
			uint32_t o[4];
#pragma unroll
			for (int u = 0; u < 4; u++) {
				const uint32_t i0 = i * 8 + 2 * u, i1 = i0 + 1;
				const uint32_t d0 = w[u] & 0xffffu, d1 = w[u] >> 16;
				const uint32_t l0 = (d0 && d0 <= i0) ? i0 - d0 : 0xffffu;
				const uint32_t l1 = (d1 && d1 <= i1) ? i1 - d1 : 0xffffu;
				o[u] = l0 | (l1 << 16);
			}
			((uint4*) W.prev)[i] = make_uint4(o[0], o[1], o[2], o[3]);
		}
		for (uint32_t i = npv * 8 + ctid; i < nlinks; i += LZ_THREADS) {
			const uint32_t d = prev[hist0 + i];
			W.prev[i] = (uint16_t) ((d && d <= i) ? i - d : 0xffffu);
		}
		if (tid == 0) { S.nomatch = 0; S.nmatch1 = 0; }
		for (uint32_t i = tid; i < SEG / 32; i += LZ_GTHREADS) { S.done[i] = 0; S.cand[i] = 0; S.cls[0][i] = 0; S.cls[1][i] = 0; }
	}
	__syncthreads();
	LZ_PROF_MARK(LZP_STAGE);

	/* ---- from here on the two groups go their own ways ---- */
	const uint32_t seg = blockIdx.x * LZ_GROUPS + grp;
	if (seg < prm.skip_segs) return;             /* dictionary: nothing to parse, nothing to emit */
	const uint64_t seg0 = (uint64_t) seg * SEG;
	const uint64_t chunk0 = grp ? chunkB0 : chunkA0;                      /* the chunk of this group's segment */
	const uint64_t chunk1 = grp ? chunkB1 : chunkA1;
	if (seg0 >= chunk1) return;                  /* ragged chunk: empty segment slot */
	uint64_t seg1 = seg0 + SEG;
	if (seg1 > chunk1) seg1 = chunk1;
	const uint32_t seg_len = (uint32_t) (seg1 - seg0);
	const uint32_t nwords = (seg_len + 31) / 32;                          /* bitmap words in use */
	const uint32_t hoff = (uint32_t) (seg0 - hist0);                      /* segment start in smem coords */

	/* what lies in front of this group's chunk in the staged window is not its history: the
	 * padding in front of a preset dictionary, the other group's chunk */
	uint32_t first_valid = (chunk0 == 0 && prm.hist_min > hist0) ? (uint32_t) (prm.hist_min - hist0) : 0u;
	if (chunk0 > hist0 && (uint32_t) (chunk0 - hist0) > first_valid) first_valid = (uint32_t) (chunk0 - hist0);

	/* ---- probe: is there anything to find at all? ----------------------------------
	 * 512 positions spread evenly over a full segment look at their first chain candidate.
	 * When not one of them matches -- random or already compressed data: at one matching
	 * position in 64, the density below which the search is skipped anyway, the chance
	 * of missing all of them is 0.03 % -- the segment is emitted as literals right away
	 * (the block will be stored: huffman.cu prices it), without pass 1, parses or the
	 * per-token emission: such segments cost 61 K cycles against 230-250 K for text or
	 * binary data, nearly all of it in those three. */
	if (seg_len == SEG && prm.skip_div) {
		/* one position in every 16, at a scrambled offset: record-structured data must not alias with the sample */
		uint32_t p = tid * (SEG / LZ_GTHREADS) + ((tid * 0x9E3779B1u) >> 28);
		if (p > SEG - 4) p = SEG - 4;
		const uint32_t j = hoff + p;
		uint32_t jmin = j > WND - 1 ? j - (WND - 1) : 0;
		if (jmin < first_valid) jmin = first_valid;
		const uint32_t q = W.prev[j];
		bool hit = false;
		if (q - jmin < j - jmin) {
			const uint32_t x = jdb_ld32u(W.data, j) ^ jdb_ld32u(W.data, q);
			hit = (x & 0xffffffu) == 0;
		}
		if (__any_sync(JDB_FULL_MASK, hit) && lane == 0) S.nmatch1 = 1;
		lz_gsync(grp);
		const bool nothing = S.nmatch1 == 0;
		lz_gsync(grp);                               /* everybody has read the flag: pass 1 counts in the same word */
		if (nothing) {
			uint32_t* const hist = S.need;
			for (uint32_t i = tid; i < NSYM; i += LZ_GTHREADS) hist[i] = 0;
			lz_gsync(grp);
			uint32_t* const out = tok + seg0;
			for (uint32_t k = 0; k < PER_THREAD; k++) {
				const uint32_t pp = tid + k * LZ_GTHREADS;
				const uint32_t b = W.data[hoff + pp];
				out[pp] = b;
				atomicAdd(&hist[b], 1u);
			}
			if (tid == 0) { S.nmatch1 = 0; seg_ntok[seg] = SEG; }
			lz_gsync(grp);
			for (uint32_t i = tid; i < NSYM; i += LZ_GTHREADS) seg_hist[(uint64_t) seg * NSYM + i] = hist[i];
			LZ_PROF_MARK(LZP_EMIT);
#if defined(LZ_PROF) && !defined(JDB_SIMT_EMU)
			if (tid == 0) atomicAdd(&g_lz_prof[LZP_N], 1ull);
#endif
			return;
		}
		if (tid == 0) S.nmatch1 = 0;
		lz_gsync(grp);
	}

	/* ---- pass 1: the first chain candidate of every position -----------------
	 * One link, one comparison per position, converged.  It seeds the search
	 * proper (best so far, chain advanced by one), feeds the tentative parse that
	 * tells which positions a parser is likely to visit, and classifies: is there
	 * a chain left to walk at all (`cand`), and is it probably a long one (`lng`:
	 * the nearer the previous occurrence of a hash, the more of them the window
	 * holds) -- long searches are started first so that short ones fill the gaps
	 * instead of everybody waiting for a few long ones at the end. */
	{
		uint32_t nm = 0;
		for (uint32_t k = 0; k < PER_THREAD; k++) {
			if (k * LZ_GTHREADS >= seg_len) break;
			const uint32_t p = tid + k * LZ_GTHREADS;
			uint32_t result = 0;
			uint32_t cls = 0;
			bool cand = false;
			if (p < seg_len && seg_len - p >= MINLEN) {
				uint32_t maxlen = seg_len - p;
				if (maxlen > MAXLEN) maxlen = MAXLEN;
				const uint32_t j = hoff + p;
				uint32_t jmin = j > WND - 1 ? j - (WND - 1) : 0;
				if (jmin < first_valid) jmin = first_valid;
				const uint32_t dmax = j - jmin;
				const uint32_t q = W.prev[j];
				if (q - jmin < dmax) {
					uint32_t len = 0;
					while (len < maxlen) {
						const uint32_t x = jdb_ld32u(W.data, j + len) ^ jdb_ld32u(W.data, q + len);
						if (x) { len += (uint32_t) (__ffs((int) x) - 1) >> 3; break; }
						len += 4;
					}
					if (len > maxlen) len = maxlen;
					if (len >= MINLEN) result = (len << 16) | (j - q);
					/* anything left to search (a first match that does not end the search by
					 * itself)?  Then: how many candidates follow?  Up to `prewalk` more links
					 * are followed: a chain that ends before that has its exact length, a
					 * longer one is extrapolated from the span of the links seen. */
					if (len < prm.nice && len < maxlen && maxlen >= 8) {
						uint32_t qk = q, nk = 0;
#pragma unroll
						for (uint32_t u = 0; u < 8; u++) {
							const uint32_t qn = W.prev[qk];
							const bool ok = u < prm.prewalk && qn - jmin < dmax;
							qk = ok ? qn : qk;
							nk += ok ? 1u : 0u;
						}
						/* nk candidates after the first one (at least: the walk stopped at prewalk) */
						/* classes: 1-2, 3-6, 7-24, 25.. steps */
						cls = lz_class(nk);
						if (nk == prm.prewalk && nk) {
							/* nk + 1 occurrences within d bytes: about N / d - 1 steps in the whole window */
							const uint32_t d = j - qk, N = (nk + 1) * (WND - 1);
							const uint32_t ec = N < 4 * d ? 0u : N < 8 * d ? 1u : N < 26 * d ? 2u : 3u;
							if (ec > cls) cls = ec;
						}
						const uint32_t capc = lz_class(prm.chain - 1);
						if (cls > capc) cls = capc;
						cand = nk != 0;
					}
				}
			}
			if (p < SEG) S.m[p] = result;
			const unsigned bc = __ballot_sync(JDB_FULL_MASK, cand);
			const unsigned b0 = __ballot_sync(JDB_FULL_MASK, cls & 1u);
			const unsigned b1 = __ballot_sync(JDB_FULL_MASK, cls & 2u);
			nm += (uint32_t) __popc(__ballot_sync(JDB_FULL_MASK, result != 0));
			if (lane == 0) { S.cand[p >> 5] = bc; S.cls[0][p >> 5] = b0; S.cls[1][p >> 5] = b1; }
		}
		if (lane == 0 && nm) atomicAdd(&S.nmatch1, nm);
	}
	lz_gsync(grp);
	LZ_PROF_MARK(LZP_PASS1);

	/* ---- which positions get the full search ----------------------------------
	 * A parse over the matches known so far marks the positions a parser visits;
	 * those (and their successors, which the lazy rule reads) are searched with
	 * the full chain budget, everything else keeps its pass-1 match.  The
	 * reference searches only where its parser goes, too (skipbytes2,
	 * src/deflator.c:2729).  A segment in which next to nothing matched (random or
	 * already compressed data) is not searched at all. */
	uint32_t nrounds = prm.rounds ? prm.rounds : 1u;
	if (prm.skip_div && S.nmatch1 * prm.skip_div < seg_len) nrounds = 0;
	for (uint32_t round = 0; round < nrounds; round++) {
	{
		uint32_t nd = 0;
		if (prm.rounds) {
			lz_parse_phase(S, grp, prm, tid, seg_len, prm.lazy && (prm.tlazy || round > 0) LZ_PROF_ARG);
			LZ_PROF_MARK(LZP_PARSE);
			uint32_t so = 0;
			if (tid < nwords) {
				nd = lz_path_word(S, tid);
				if (prm.succ) {
					/* successors: shift the path bits up by one position (carry from the word below) */
					uint32_t up = nd << 1;
					if (tid) up |= lz_path_word(S, tid - 1) >> 31;
					so = up & ~nd;
					nd |= up;
				}
				/* only where a chain is left to walk, and what has not been searched in an earlier round */
				nd &= S.cand[tid] & ~S.done[tid];
			}
			lz_gsync(grp);                      /* the parse bitmaps are dead from here */
			/* positions that are searched only because the lazy rule looks at them: half the
			 * chain budget, as the reference searches with a match in hand (getmatch2, chain >> 1) */
			if (tid < SEG / 32) S.fix[tid] = prm.succ > 1 ? so & nd : 0u;
		} else {
			if (tid < nwords) nd = S.cand[tid];
			if (tid < SEG / 32) S.fix[tid] = 0;
		}
		if (tid < SEG / 32) {
			S.need[tid] = nd;
			S.done[tid] |= nd;
		}
		if (tid == 0) { S.ccount[0] = 0; S.ccount[1] = 0; S.next_batch = 0; S.improved = 0; }
	}
	lz_gsync(grp);

	/* ---- the positions to search, grouped by class ------------------------------
	 * What a lane of the search does is the same as what its neighbours do as long
	 * as their chains are equally long, so the positions are sorted by class into
	 * four lists (the segment's token slots in HBM serve as scratch: 32 Ki entries
	 * of 16 bits; classes 3 and 2 grow towards each other in the lower half, 1 and
	 * 0 in the upper half).  One packed shared-memory atomic per 32 positions. */
	uint16_t* const joblist = (uint16_t*) (tok + seg0);
	{
		/* a warp owns 16 bitmap words: count its positions per class, reserve its share of
		 * the four lists with two atomics, then hand out the slots */
		const uint32_t WPW = SEG / 32 / (LZ_GTHREADS / 32);
		const uint32_t w0 = (tid >> 5) * WPW;
		uint32_t lo = 0, hi = 0;                              /* counts: class 0 | class 1 << 16, class 2 | class 3 << 16 */
		if (w0 < nwords && lane < WPW) {
			const uint32_t wi = w0 + lane;
			const uint32_t nd = wi < nwords ? S.need[wi] : 0u;
			const uint32_t c0 = S.cls[0][wi], c1 = S.cls[1][wi];
			lo = (uint32_t) __popc(nd & ~c1 & ~c0) | (uint32_t) __popc(nd & ~c1 & c0) << 16;
			hi = (uint32_t) __popc(nd & c1 & ~c0) | (uint32_t) __popc(nd & c1 & c0) << 16;
		}
		if (w0 < nwords) {
			/* exclusive prefix over the warp's 16 words (fields cannot overflow: <= 512 per warp) */
			uint32_t plo = lo, phi = hi;
			for (int o = 1; o < (int) WPW; o <<= 1) {
				const uint32_t tl = __shfl_up_sync(JDB_FULL_MASK, plo, o), th = __shfl_up_sync(JDB_FULL_MASK, phi, o);
				if ((int) lane >= o) { plo += tl; phi += th; }
			}
			uint32_t blo = 0, bhi = 0;
			if (lane == WPW - 1) {
				blo = atomicAdd(&S.ccount[0], plo);
				bhi = atomicAdd(&S.ccount[1], phi);
			}
			blo = __shfl_sync(JDB_FULL_MASK, blo, WPW - 1) + plo - lo;
			bhi = __shfl_sync(JDB_FULL_MASK, bhi, WPW - 1) + phi - hi;
			for (uint32_t i = 0; i < WPW; i++) {
				const uint32_t wi = w0 + i;
				if (wi >= nwords) break;
				const uint32_t base_lo = __shfl_sync(JDB_FULL_MASK, blo, (int) i), base_hi = __shfl_sync(JDB_FULL_MASK, bhi, (int) i);
				const uint32_t nd = S.need[wi];
				if ((nd >> lane) & 1u) {
					const uint32_t c0 = S.cls[0][wi], c1 = S.cls[1][wi];
					const uint32_t b0 = (c0 >> lane) & 1u, b1 = (c1 >> lane) & 1u;
					const uint32_t same = nd & (b0 ? c0 : ~c0) & (b1 ? c1 : ~c1);
					const uint32_t base = ((b1 ? base_hi : base_lo) >> (16 * b0)) & 0xffffu;
					const uint32_t k = base + (uint32_t) __popc(same & ((1u << lane) - 1u));
					/* class 3: 0 up, class 2: 16383 down, class 1: 16384 up, class 0: 32767 down */
					const uint32_t slot = b1 ? (b0 ? k : SEG - 1 - k) : (b0 ? SEG + k : 2 * SEG - 1 - k);
					joblist[slot] = (uint16_t) (wi * 32 + lane);
				}
			}
		}
	}
	lz_gsync(grp);
	LZ_PROF_MARK(LZP_NEED);

	/* ---- match search -------------------------------------------------------
	 * Batches of 32 positions of one class, one per lane, searched to completion
	 * together; the batches are handed out from one shared counter, longest class
	 * first: the long walks start first and the short ones fill the end of the
	 * round (a few long stragglers would leave everybody else waiting).
	 *
	 * One chain step = next link, window / budget check, and the first 8 bytes of
	 * the candidate against the first 8 bytes of the position (kept in registers):
	 * the same instructions for every lane, no pre-filter branch.  Most candidates
	 * are decided by that; one that matches all 8 bytes (and, against a best of 8
	 * or more, the byte that would extend it) is compared to its end right away by
	 * the lanes concerned. */
	{
		const uint32_t nice = prm.nice;
		const uint32_t n3 = S.ccount[1] >> 16, n2 = S.ccount[1] & 0xffffu;
		const uint32_t n1 = S.ccount[0] >> 16, n0 = S.ccount[0] & 0xffffu;
		const uint32_t e3 = (n3 + 31) / 32, e2 = e3 + (n2 + 31) / 32, e1 = e2 + (n1 + 31) / 32, e0 = e1 + (n0 + 31) / 32;
		LZ_PROF_COUNTERS();
		/* the position of lane `lane` in batch b, NOPOS past the end of its class */
#define LZ_BATCH_ENTRY(b, out) \
		do { \
			uint32_t slot_ = NOPOS; \
			if ((b) < e3) { const uint32_t k_ = (b) * 32 + lane; if (k_ < n3) slot_ = k_; } \
			else if ((b) < e2) { const uint32_t k_ = ((b) - e3) * 32 + lane; if (k_ < n2) slot_ = SEG - 1 - k_; } \
			else if ((b) < e1) { const uint32_t k_ = ((b) - e2) * 32 + lane; if (k_ < n1) slot_ = SEG + k_; } \
			else if ((b) < e0) { const uint32_t k_ = ((b) - e1) * 32 + lane; if (k_ < n0) slot_ = 2 * SEG - 1 - k_; } \
			(out) = slot_ == NOPOS ? NOPOS : (uint32_t) joblist[slot_]; \
		} while (0)
		uint32_t bcur = 0, pnext = NOPOS;
		if (lane == 0) bcur = atomicAdd(&S.next_batch, 1u);
		bcur = __shfl_sync(JDB_FULL_MASK, bcur, 0);
		LZ_BATCH_ENTRY(bcur, pnext);
		while (bcur < e0) {
			const uint32_t pcur = pnext;
			/* the batch after this one: its entries are on their way while this one is searched */
			uint32_t bnext = 0;
			if (lane == 0) bnext = atomicAdd(&S.next_batch, 1u);
			bnext = __shfl_sync(JDB_FULL_MASK, bnext, 0);
			LZ_BATCH_ENTRY(bnext, pnext);
			bcur = bnext;
			const bool mine = pcur != NOPOS;
			LZ_PROF_COUNT(4, 1); LZ_PROF_COUNT(5, __popc(__ballot_sync(JDB_FULL_MASK, mine)));

			/* set the search up; pass 1 already looked at the first candidate (and there
			 * is a second one: class > 0) */
			uint32_t p = 0, j = 0, jmin = 0, maxlen = 0, best = 0, bestd = 0, cur = 0, steps = 0, cb = 0, first_len = 0;
			uint32_t dmax = 0, jw0 = 0, jw1 = 0;
			bool walking = false;
			if (mine) {
				p = pcur;
				maxlen = seg_len - p;
				if (maxlen > MAXLEN) maxlen = MAXLEN;
				j = hoff + p;
				jmin = j > WND - 1 ? j - (WND - 1) : 0;
				if (jmin < first_valid) jmin = first_valid;
				dmax = j - jmin;
				const uint32_t v1 = S.m[p];
				best = MINLEN - 1;
				if (v1) { best = v1 >> 16; bestd = v1 & 0xffffu; }
				first_len = best;
				steps = (((S.fix[p >> 5] >> (p & 31u)) & 1u) ? prm.chain >> 1 : prm.chain) - 1;
				cur = W.prev[j];
				cb = W.data[j + best];
				const uint32_t* jw = (const uint32_t*) (W.data + (j & ~3u));
				const uint32_t sh = (j & 3u) * 8u;
				const uint32_t w0 = jw[0], w1 = jw[1], w2 = jw[2];
				jw0 = __funnelshift_r(w0, w1, sh);
				jw1 = __funnelshift_r(w1, w2, sh);
				walking = true;
			}

			for (;;) {
				const unsigned wm = __ballot_sync(JDB_FULL_MASK, walking);
				if (!wm) break;
				LZ_PROF_COUNT(0, 1); LZ_PROF_COUNT(1, __popc(wm));
				uint32_t cq = NOPOS;             /* a candidate to compare beyond 8 bytes */
				/* straight-line on purpose (selects, no branches): the lanes of a batch walk
				 * chains of similar length, so nearly all of them are here together.  A
				 * lane whose chain has ended computes on garbage (any 16-bit q is a valid
				 * shared-memory offset) and keeps nothing. */
#pragma unroll
				for (int u = 0; u < LZ_WALK_STEPS; u++) {
					const uint32_t q = W.prev[cur];
					const bool live = walking && q - jmin < dmax && steps != 0 && cq == NOPOS;
					LZ_STAT(if (live && jdb_emu_lz_steps) jdb_emu_lz_steps[seg0 + p]++;)
					const uint32_t* cw = (const uint32_t*) (W.data + (q & ~3u));
					const uint32_t sh = (q & 3u) * 8u;
					const uint32_t c0 = cw[0], c1 = cw[1], c2 = cw[2];
					const uint32_t x0 = __funnelshift_r(c0, c1, sh) ^ jw0;
					const uint32_t x1 = __funnelshift_r(c1, c2, sh) ^ jw1;
					const uint32_t xs = x0 ? x0 : x1;
					const uint32_t len = (x0 ? 0u : 4u) + ((uint32_t) (__ffs((int) xs) - 1) >> 3);
					const bool all8 = (x0 | x1) == 0;
					const bool imp = live && !all8 && len > best;       /* 4 <= len < 8 <= maxlen (pass 1 sees to that) */
					if (live && all8) cq = q;
					best = imp ? len : best;
					bestd = imp ? j - q : bestd;
					if (u == 0) {
						/* a lane that stops in the first step of the pair must not come back in the second */
						walking = walking && (live || cq != NOPOS);
					}
					steps = live ? steps - 1 : steps;
					if (imp && steps > prm.patience) steps = prm.patience;
					cur = live ? q : cur;
					if (u == LZ_WALK_STEPS - 1) walking = walking && (live || cq != NOPOS);
				}
				if (__any_sync(JDB_FULL_MASK, cq != NOPOS)) {
					LZ_PROF_COUNT(2, 1);
					/* against a best of 8 or more only a candidate that matches the byte after it matters */
					if (cq != NOPOS && (best < 8 || W.data[cq + best] == cb)) {
						LZ_PROF_COUNT(3, 1);
						uint32_t len = 8;
						while (len < maxlen) {
							const uint32_t x = jdb_ld32u(W.data, j + len) ^ jdb_ld32u(W.data, cq + len);
							if (x) { len += (uint32_t) (__ffs((int) x) - 1) >> 3; break; }
							len += 4;
						}
						if (len > maxlen) len = maxlen;
						if (len > best) {
							best = len;
							bestd = j - cq;
							cb = W.data[j + best];
							if (len >= nice || len == maxlen) walking = false;
							/* a match in hand: the rest of the chain gets the reduced budget */
							if (steps > prm.patience) steps = prm.patience;
						}
					}
				}
			}
			if (mine) S.m[p] = best >= MINLEN ? (best << 16) | bestd : 0;
			if (prm.r2min) {
				const unsigned im = __ballot_sync(JDB_FULL_MASK, mine && best >= MINLEN && best > first_len);
				if (lane == 0 && im) atomicAdd(&S.improved, (uint32_t) __popc(im));
			}
		}
#undef LZ_BATCH_ENTRY
		LZ_PROF_FLUSH();
	}
	lz_gsync(grp);
	LZ_PROF_MARK(LZP_SEARCH);
	/* a round that improved next to nothing leaves the next parse where it was: no second round */
	if (prm.r2min && S.improved * prm.r2min < seg_len) break;
	}       /* rounds */

	/* ---- 3-byte matches ------------------------------------------------------
	 * The hash-4 chains cannot see them.  The reference finds them with a second,
	 * 3-byte hash table when literals dominate (getmatch2, src/deflator.c:2676-2711:
	 * at most two probes, only offsets <= 8192).  Here positions that found
	 * nothing probe a fixed set of short distances instead -- the strides of
	 * record-structured binary data, which is where such matches occur. */
	if (prm.short3) {
		/* only where literals dominate (the reference's doshortmatches switch,
		 * src/deflator.c:2928-2933): more than 40 % of the positions found nothing */
		uint32_t nomatch = 0;
		for (uint32_t k = 0; k < PER_THREAD; k++) {
			const uint32_t p = tid + k * LZ_GTHREADS;
			nomatch += (p < seg_len && S.m[p] == 0) ? 1u : 0u;
		}
		for (int o = 16; o; o >>= 1) nomatch += __shfl_xor_sync(JDB_FULL_MASK, nomatch, o);
		if ((tid & 31u) == 0) atomicAdd(&S.nomatch, nomatch);
		lz_gsync(grp);
	}
	if (prm.short3 && (prm.short3 > 1 || (S.nomatch * 5u > seg_len * 2u && S.nomatch * 20u < seg_len * 19u))) {
		const uint32_t sv = first_valid;             /* 40 % .. 95 %: not random data either */
		for (uint32_t k = 0; k < PER_THREAD; k++) {
			const uint32_t p = tid + k * LZ_GTHREADS;
			if (p + 3 > seg_len || S.m[p] != 0) continue;
			const uint32_t j = hoff + p;
			const uint32_t w = jdb_ld32u(W.data, j) & 0xffffffu;
			uint32_t found = 0;
#pragma unroll
			for (int u = 0; u < 12; u++) {
				const uint32_t d = c_short_dist[u];
				if (!found && d <= j && j - d >= sv && (jdb_ld32u(W.data, j - d) & 0xffffffu) == w) found = d;
			}
			if (found) S.m[p] = (3u << 16) | found;
		}
		lz_gsync(grp);
	}

	LZ_PROF_MARK(LZP_SHORT3);
	lz_parse_phase(S, grp, prm, tid, seg_len, prm.lazy != 0 LZ_PROF_ARG);
	LZ_PROF_MARK(LZP_FINALPARSE);

	/* ---- emit: 16 consecutive positions per thread ---- */
	{
		const uint32_t p0 = tid * PER_THREAD;
		uint32_t bits = 0, takes = 0;
		if (p0 < seg_len) {
			bits = (lz_path_word(S, p0 >> 5) >> (p0 & 31)) & 0xffffu;
			takes = S.take[p0 >> 5] >> (p0 & 31);
		}

		const uint32_t cnt = (uint32_t) __popc(bits);
		uint32_t incl = cnt;
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(JDB_FULL_MASK, incl, o);
			if ((int) (tid & 31) >= o) incl += t;
		}
		if ((tid & 31) == 31) S.warp_sum[tid >> 5] = incl;
		lz_gsync(grp);
		uint32_t* const hist = S.need;                   /* need[] and done[] are dead: the histogram takes their place */
		for (uint32_t i = tid; i < NSYM; i += LZ_GTHREADS) hist[i] = 0;
		if (tid < 32) {
			uint32_t v = tid < LZ_GTHREADS / 32 ? S.warp_sum[tid] : 0u;
			uint32_t iv = v;
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t t = __shfl_up_sync(JDB_FULL_MASK, iv, o);
				if ((int) tid >= o) iv += t;
			}
			if (tid < LZ_GTHREADS / 32) S.warp_sum[tid] = iv - v;        /* exclusive */
			if (tid == LZ_GTHREADS / 32 - 1) seg_ntok[seg] = iv;
		}
		lz_gsync(grp);
		uint32_t o = S.warp_sum[tid >> 5] + incl - cnt;
		uint32_t* out = tok + seg0;
		while (bits) {
			const uint32_t b = (uint32_t) (__ffs((int) bits) - 1);
			const uint32_t p = p0 + b;
			bits &= bits - 1;
			uint32_t token;
			if ((takes >> b) & 1u) {
				const uint32_t v = S.m[p];
				const uint32_t len = v >> 16, dist = v & 0xffffu;
				token = TOK_MATCH | ((len - 3) << 16) | (dist - 1);
				atomicAdd(&hist[257 + len_symbol(len)], 1u);
				atomicAdd(&hist[DSYM0 + dist_symbol(dist)], 1u);
			} else {
				token = W.data[hoff + p];
				atomicAdd(&hist[token], 1u);
			}
			out[o++] = token;
		}
	}
	lz_gsync(grp);
	for (uint32_t i = tid; i < NSYM; i += LZ_GTHREADS) seg_hist[(uint64_t) seg * NSYM + i] = S.need[i];
	LZ_PROF_MARK(LZP_EMIT);
#if defined(LZ_PROF) && !defined(JDB_SIMT_EMU)
	if (tid == 0) atomicAdd(&g_lz_prof[LZP_N], 1ull);
#endif
}

/* ---- launchers ------------------------------------------------------------- */

extern "C" int jdb_lz_chain(const uint8_t* in, uint64_t n, uint32_t chunk_bytes, uint32_t range,
                            const uint32_t* chunk_len, uint16_t* prev, uint16_t* heads, jdb_stream s)
{
	if (n == 0) return JDB_OK;
	const size_t smem = sizeof(ChainSmem);
	JDB_CONFIGURE_SMEM(chain_kernel, smem);
	const uint64_t items = (n + range - 1) / range;
	if (range < WND || range >= chunk_bytes) heads = NULL;      /* one range per chunk: nothing to link across */
	JDB_LAUNCH(chain_kernel, dim3((unsigned) items), dim3(CH_THREADS), smem, s, in, n, chunk_bytes, range, chunk_len, prev, heads);
	int r = jdb_rt_check_launch("chain_kernel");
	if (r != JDB_OK || heads == NULL) return r;
	JDB_LAUNCH(chain_fix_kernel, dim3((unsigned) items), dim3(256), 0, s, in, n, chunk_bytes, range, chunk_len, prev, (const uint16_t*) heads);
	return jdb_rt_check_launch("chain_fix_kernel");
}

extern "C" size_t jdb_lz_chain_heads_bytes(uint64_t n, uint32_t chunk_bytes)
{
	/* one table per range of at least 64 KiB, and at least one range per chunk */
	const uint64_t ranges = n / 65536u + (n + chunk_bytes - 1) / chunk_bytes + 2;
	return (size_t) ranges * (1u << HASH_BITS) * 2u;
}

/* experiment switches, read once per process (JDB_LZ_*; -1 = the level's default) */
struct LzEnv { int rounds, short3, patience, tlazy, succ, prewalk, skip_div, r2min; };
static int lz_env_int(const char* name) { const char* v = getenv(name); return v ? atoi(v) : -1; }
static const LzEnv& lz_env()
{
	static const LzEnv e = { lz_env_int("JDB_LZ_ROUNDS"), lz_env_int("JDB_LZ_SHORT3"), lz_env_int("JDB_LZ_PATIENCE"),
	                         lz_env_int("JDB_LZ_TLAZY"), lz_env_int("JDB_LZ_SUCC"), lz_env_int("JDB_LZ_PREWALK"), lz_env_int("JDB_LZ_SKIP_DIV"),
	                         lz_env_int("JDB_LZ_R2MIN") };
	return e;
}

extern "C" int jdb_lz_parse(const uint8_t* in, uint64_t n, uint32_t chunk_bytes,
                            const uint32_t* chunk_len, const uint16_t* prev,
                            uint32_t good, uint32_t nice, uint32_t chain, uint32_t lazy,
                            uint32_t skip_segs, uint32_t hist_min,
                            uint32_t* tok, uint32_t* seg_ntok, uint32_t* seg_hist, jdb_stream s)
{
	if (n == 0) return JDB_OK;
	const size_t smem = sizeof(LzSmem);
	JDB_CONFIGURE_SMEM(lz_kernel, smem);
#if defined(LZ_PROF) && !defined(JDB_SIMT_EMU)
	{ static int once_; if (!once_) { once_ = 1; atexit(lz_prof_dump); } }
#endif
	const LzEnv& env = lz_env();
	LzParams prm;
	prm.good = good; prm.nice = nice; prm.chain = chain; prm.lazy = lazy;
	prm.skip_segs = skip_segs; prm.hist_min = hist_min;
	prm.rounds = env.rounds >= 0 ? (uint32_t) env.rounds : 2u;
	prm.short3 = env.short3 >= 0 ? (uint32_t) env.short3 : 1u;
	prm.patience = env.patience >= 0 ? (uint32_t) env.patience : chain;
	prm.tlazy = env.tlazy >= 0 ? (uint32_t) env.tlazy : 0u;
	prm.succ = lazy ? (env.succ >= 0 ? (uint32_t) env.succ : 2u) : 0u;
	prm.prewalk = env.prewalk >= 0 ? (uint32_t) env.prewalk : 8u;
	if (prm.prewalk > 8) prm.prewalk = 8;
	if (prm.prewalk + 1 > chain) prm.prewalk = chain ? chain - 1 : 0;
	prm.skip_div = env.skip_div >= 0 ? (uint32_t) env.skip_div : 64u;
	prm.r2min = env.r2min >= 0 ? (uint32_t) env.r2min : 0u;
	const uint64_t npair = (n + LZ_GROUPS * SEG - 1) / (LZ_GROUPS * SEG);
	JDB_LAUNCH(lz_kernel, dim3((unsigned) npair), dim3(LZ_THREADS), smem, s,
	           in, n, chunk_bytes, chunk_len, prev, prm, tok, seg_ntok, seg_hist);
	return jdb_rt_check_launch("lz_kernel");
}
