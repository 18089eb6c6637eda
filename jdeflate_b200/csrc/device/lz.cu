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
 *                 the segment start -- is done by 64 speculative walkers
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

#define HASH_BITS      14
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
			if (d >= WND) S.head[i] = (uint16_t) stale;
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
             const uint32_t* __restrict__ chunk_len, uint16_t* __restrict__ prev)
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
	const uint64_t start = r0 >= chunk0 + WND ? r0 - WND : chunk0;   /* warm-up start, multiple of SEG */

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
	 * rotating w0 = w1 ... would wait for the newest load every round) */
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
}

/* ---------------------------------------------------------------------------
 * lz_kernel
 * ------------------------------------------------------------------------- */

#define LZ_THREADS   1024
#ifndef LZ_WALK_STEPS
#define LZ_WALK_STEPS 4
#endif
#ifndef LZ_GRAB
#define LZ_GRAB 128u        /* positions a warp takes from the segment at a time */
#endif
#ifndef LZ_CMP_WORDS
#define LZ_CMP_WORDS 2   /* 32-bit words compared per COMPARE phase */
#endif
#define PER_THREAD   (SEG / LZ_THREADS)            /* 16 */
#define WALK_BLOCK   256u
#define WALKERS      (SEG / WALK_BLOCK)            /* 64 */
#define DATA_BYTES   (WND + SEG + 320)             /* history + segment + look-ahead/guard */

#ifdef JDB_SIMT_EMU
/* emulator-only instrumentation (tools/emu_lz_stats.py): chain steps per position */
extern "C" { uint32_t* jdb_emu_lz_iters = 0; uint8_t* jdb_emu_lz_steps = 0; }
#define LZ_STAT(x) x
#else
#define LZ_STAT(x)
#endif

struct LzParams {
	uint32_t good, nice, chain, lazy;
	uint32_t short3;                 /* probe short distances for 3-byte matches */
	uint32_t skip_segs;              /* leading segments that are preset dictionary: history only */
	uint32_t twophase;               /* full search only where a tentative parse goes */
	uint32_t hist_min;               /* first byte of the batch a match of chunk 0 may reach (dictionary start) */
	uint32_t patience;               /* chain steps still allowed once a match has been found */
};

struct LzSmem {
	uint32_t m[SEG];                 /* per position result, later flags      */
	uint16_t prev[WND + SEG];
	uint8_t  data[DATA_BYTES];
	uint32_t spec[SEG / 32];         /* positions on the speculative paths    */
	uint32_t fix[SEG / 32];          /* positions added by the stitching pass */
	uint32_t take[SEG / 32];         /* positions whose match a parser arriving there takes */
	uint32_t hist[NSYM];
	uint32_t land[WALKERS];          /* where each walker left its block      */
	uint32_t merge[WALKERS];
	uint32_t warp_sum[LZ_THREADS / 32];
	uint32_t next_pos;               /* work distribution of the match search */
	uint32_t nomatch;                /* positions without a match (3-byte probe switch) */
	uint32_t need[SEG / 32];         /* positions that get the full search */
	uint32_t done[SEG / 32];         /* ... that already got it in an earlier round */
};

__constant__ uint8_t c_short_dist[12] = { 1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64 };

static __device__ __forceinline__ uint32_t ilog2_u32(uint32_t v) { return 31 - __clz(v); }

/* the token decision for a fresh position p (segment coordinates): returns
 * the next fresh position */
static __device__ __forceinline__ uint32_t
next_pos(const uint32_t* m, uint32_t p)
{
	uint32_t v = m[p];
	return (v & M_TAKE) ? p + ((v >> 16) & 0x1ffu) : p + 1;
}

/*
 * The parse over the per-position matches in S.m: (1) per position, would a parser
 * arriving here take the match (lazy rule)?  (2) 64 speculative walkers follow the
 * decisions through their 256-position blocks; (3) thread 0 stitches the true path
 * across the blocks.  Afterwards position p is on the path iff
 * (spec bit && p >= merge[block]) || fix bit.  All threads call it.
 */
static __device__ __forceinline__ void
lz_parse_phase(LzSmem& S, const LzParams& prm, const uint32_t tid, const uint32_t seg_len)
{
	for (uint32_t i = tid; i < SEG / 32; i += LZ_THREADS) { S.spec[i] = 0; S.fix[i] = 0; }
	__syncthreads();
	/* ---- per position: would a parser arriving here take the match? ----
	 * The answers also go into a bitmap (one ballot per 32 positions) that lets
	 * the walkers below cross literal runs in one step. */
	for (uint32_t k = 0; k < PER_THREAD; k++) {
		const uint32_t p = tid + k * LZ_THREADS;
		const uint32_t v = p < seg_len ? (S.m[p] & ~M_TAKE) : 0u;
		const uint32_t len = v >> 16;
		bool take = len != 0;
		if (take && prm.lazy && len < prm.good && p + 1 < seg_len) {
			const uint32_t w = S.m[p + 1] & ~M_TAKE;
			const uint32_t nlen = w >> 16;
			if (nlen >= len) {
				/* the reference's accept rule, src/deflator.c:2865-2879 */
				const int32_t delta = (int32_t) nlen - (int32_t) len;
				if (delta > 4) take = false;
				else {
					const int32_t l1 = (int32_t) ilog2_u32(v & 0xffffu), l2 = (int32_t) ilog2_u32(w & 0xffffu);
					if ((delta << 2) + (l1 - l2) >= 2) take = false;
				}
			}
		}
		const unsigned tb = __ballot_sync(JDB_FULL_MASK, take);
		if ((tid & 31u) == 0) S.take[p >> 5] = tb;
		/* m[p + 1] is read by the neighbouring lane in this same iteration: the flag
		 * goes in only after every lane of the warp has read (the ballot above) --
		 * and the mask keeps a flag set by another warp's earlier iteration harmless */
		if (p < seg_len) S.m[p] = take ? (v | M_TAKE) : v;      /* (a flag of an earlier parse must not survive) */
	}
	__syncthreads();

	/* ---- speculative walkers: one per 256 positions ---- */
	if (tid < WALKERS) {
		const uint32_t b0 = tid * WALK_BLOCK, b1 = b0 + WALK_BLOCK;
		uint32_t p = b0;
		if (b0 < seg_len) {
			const uint32_t lim = b1 < seg_len ? b1 : seg_len;
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
					p += (S.m[p] >> 16) & 0x1ffu;
				} else {
					uint32_t nlit = bits ? (uint32_t) (__ffs((int) bits) - 1) : 32u - sh;
					if (nlit > lim - p) nlit = lim - p;
					sw |= (nlit >= 32u ? 0xffffffffu : ((1u << nlit) - 1u)) << sh;
					p += nlit;
				}
			}
			S.spec[curw] = sw;
		}
		S.land[tid] = p;
		S.merge[tid] = b0;
	}
	__syncthreads();

	/* ---- stitch: thread 0 follows the true path across the walker blocks ---- */
	if (tid == 0) {
		uint32_t t = 0;                                   /* true entry position */
		for (uint32_t w = 0; w < WALKERS; w++) {
			const uint32_t b0 = w * WALK_BLOCK, b1 = b0 + WALK_BLOCK;
			if (b0 >= seg_len) break;
			const uint32_t lim = b1 < seg_len ? b1 : seg_len;
			if (t >= lim) { S.merge[w] = lim; continue; }          /* block jumped over */
			if (t == b0) { S.merge[w] = b0; t = S.land[w]; continue; }
			uint32_t p = t;
			while (p < lim && !((S.spec[p >> 5] >> (p & 31)) & 1u)) {
				S.fix[p >> 5] |= 1u << (p & 31);
				p = next_pos(S.m, p);
			}
			if (p < lim) { S.merge[w] = p; t = S.land[w]; }
			else { S.merge[w] = lim; t = p; }
		}
	}
	__syncthreads();
}

template <bool ROUNDS>
__global__ void __launch_bounds__(LZ_THREADS, 1)
lz_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes,
          const uint32_t* __restrict__ chunk_len, const uint16_t* __restrict__ prev, LzParams prm,
          uint32_t* __restrict__ tok, uint32_t* __restrict__ seg_ntok, uint32_t* __restrict__ seg_hist)
{
	JDB_DYN_SMEM(smem_raw);
	LzSmem& S = *(LzSmem*) smem_raw;
	const uint32_t tid = threadIdx.x;
	const uint32_t seg = blockIdx.x;
	if (seg < prm.skip_segs) return;             /* dictionary: nothing to parse, nothing to emit */

	const uint64_t seg0 = (uint64_t) seg * SEG;
	const uint64_t chunk0 = seg0 / chunk_bytes * chunk_bytes;
	const uint64_t chunk1 = chunk_end(chunk_len, chunk0, chunk_bytes, n);
	if (seg0 >= chunk1) return;                  /* ragged chunk: empty segment slot */
	uint64_t seg1 = seg0 + SEG;
	if (seg1 > chunk1) seg1 = chunk1;
	const uint32_t seg_len = (uint32_t) (seg1 - seg0);
	const uint64_t hist0 = seg0 >= chunk0 + WND ? seg0 - WND : chunk0;    /* first staged byte */
	const uint32_t hoff = (uint32_t) (seg0 - hist0);                      /* segment start in smem coords */

	/* ---- stage bytes and links (16-byte vectors; `in` and hist0 are 16-aligned) ---- */
	{
		const uint32_t nbytes = (uint32_t) ((chunk1 - hist0) < (uint64_t) DATA_BYTES ? (chunk1 - hist0) : DATA_BYTES);
		const uint4* src = (const uint4*) (in + hist0);
		uint4* dst = (uint4*) S.data;
		const uint32_t nv = nbytes / 16;
		for (uint32_t i = tid; i < nv; i += LZ_THREADS) dst[i] = __ldg(src + i);
		/* tail bytes, then zeros as far as a comparison can look past the data (320 guard) */
		const uint32_t fill_end = nv * 16 + 16 + 320 < DATA_BYTES ? nv * 16 + 16 + 320 : DATA_BYTES;
		for (uint32_t i = nv * 16 + tid; i < fill_end; i += LZ_THREADS)
			S.data[i] = i < nbytes ? in[hist0 + i] : 0;
		/* links are staged as absolute shared-memory positions (0xffff = none), so a
		 * chain step is one load and one range check */
		const uint32_t nlinks = hoff + seg_len;
		const uint4* ps = (const uint4*) (prev + hist0);
		const uint32_t npv = nlinks / 8;
		for (uint32_t i = tid; i < npv; i += LZ_THREADS) {
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
			((uint4*) S.prev)[i] = make_uint4(o[0], o[1], o[2], o[3]);
		}
		for (uint32_t i = npv * 8 + tid; i < nlinks; i += LZ_THREADS) {
			const uint32_t d = prev[hist0 + i];
			S.prev[i] = (uint16_t) ((d && d <= i) ? i - d : 0xffffu);
		}
		if (tid == 0) { S.next_pos = 0; S.nomatch = 0; }
		for (uint32_t i = tid; i < SEG / 32; i += LZ_THREADS) { S.spec[i] = 0; S.fix[i] = 0; }
		for (uint32_t i = tid; i < NSYM; i += LZ_THREADS) S.hist[i] = 0;
	}
	__syncthreads();

	/* with a preset dictionary the padding in front of it is not history */
	const uint32_t first_valid = (chunk0 == 0 && prm.hist_min > hist0) ? (uint32_t) (prm.hist_min - hist0) : 0u;

	/* ---- pass 1: the first chain candidate of every position -----------------
	 * One link, one comparison per position, converged.  It seeds the search
	 * proper (best so far, chain advanced by one) and, in two-phase mode, feeds a
	 * tentative parse that tells which positions a parser is likely to visit. */
	if (ROUNDS)
	for (uint32_t k = 0; k < PER_THREAD; k++) {
		const uint32_t p = tid + k * LZ_THREADS;
		uint32_t result = 0;
		if (p < seg_len && seg_len - p >= MINLEN) {
			uint32_t maxlen = seg_len - p;
			if (maxlen > MAXLEN) maxlen = MAXLEN;
			const uint32_t j = hoff + p;
			uint32_t jmin = j > WND - 1 ? j - (WND - 1) : 0;
			if (jmin < first_valid) jmin = first_valid;
			const uint32_t q = S.prev[j];
			if (q - jmin < j - jmin) {
				uint32_t len = 0;
				while (len < maxlen) {
					const uint32_t x = jdb_ld32u(S.data, j + len) ^ jdb_ld32u(S.data, q + len);
					if (x) { len += (uint32_t) (__ffs((int) x) - 1) >> 3; break; }
					len += 4;
				}
				if (len > maxlen) len = maxlen;
				if (len >= MINLEN) result = (len << 16) | (j - q);
			}
		}
		if (p < SEG) S.m[p] = result;
	}
	__syncthreads();

	/* ---- which positions get the full search ----------------------------------
	 * Two-phase mode: a parse over the pass-1 matches marks the positions a parser
	 * visits; those and their successors (the lazy rule looks one ahead) are
	 * searched with the full chain budget, everything else keeps its pass-1
	 * match.  The reference searches only where its parser goes, too
	 * (skipbytes2, src/deflator.c:2729). */
	for (uint32_t w = tid; w < SEG / 32; w += LZ_THREADS) S.done[w] = 0;
	for (uint32_t round = 0; round < (ROUNDS ? prm.twophase : 1u); round++) {
	if (ROUNDS) {
		lz_parse_phase(S, prm, tid, seg_len);
		for (uint32_t w = tid; w < SEG / 32; w += LZ_THREADS) {
			const uint32_t mg = S.merge[w >> 3];
			uint32_t pb = S.spec[w];
			if (mg >= (w + 1) * 32) pb = 0;
			else if (mg > w * 32) pb &= 0xffffffffu << (mg - w * 32);
			pb |= S.fix[w];
			S.need[w] = pb;
		}
		__syncthreads();
		for (uint32_t w = tid; w < SEG / 32; w += LZ_THREADS) {
			/* successors: shift the path bits up by one position (carry from the word below) */
			uint32_t nb = S.need[w];
			uint32_t up = nb << 1;
			if (w) up |= S.need[w - 1] >> 31;
			S.fix[w] = nb | up;                   /* staged in fix[]: need[] is still being read by neighbours */
		}
		__syncthreads();
		for (uint32_t w = tid; w < SEG / 32; w += LZ_THREADS) {
			/* only what has not been searched in an earlier round */
			const uint32_t nd = S.fix[w] & ~S.done[w];
			S.need[w] = nd;
			S.done[w] |= nd;
		}
	}
	if (tid == 0) S.next_pos = 0;
	__syncthreads();

	/* ---- match search -------------------------------------------------------
	 * Positions are handed out dynamically (one shared counter, warp-aggregated).
	 * A lane is in one of three modes and the warp runs three phases per round:
	 *   WALK     one chain step: next link, window / budget check, pre-filter on
	 *            the byte that would extend the best match so far
	 *   COMPARE  up to 16 more bytes of the candidate that passed the pre-filter
	 *   FETCH    store the finished position, take the next one
	 * WALK runs every round; COMPARE and FETCH only when enough lanes wait for
	 * them (or nobody can walk), so the two rare, long phases execute with many
	 * lanes active instead of diverging on every step. */
	{
		enum { M_FETCH = 0, M_WALK = 1, M_COMPARE = 2, M_DONE = 3 };
		const uint32_t nice = prm.nice;
		const uint32_t lane = tid & 31u;
		uint32_t mode = M_FETCH;
		uint32_t p = 0xffffffffu, j = 0, jmin = 0, maxlen = 0, best = 0, bestd = 0, cur = 0, steps = 0, cb = 0;
		uint32_t cq = 0, clen = 0, dmax = 0;
		/* positions are handed out in blocks of LZ_GRAB per warp; a short segment (a small
		 * record of a batch) in blocks of 32 so that all warps get some, and only up to
		 * its end */
		const uint32_t grab = seg_len >= SEG / 2 ? LZ_GRAB : 32u;
		const uint32_t slim = (seg_len + LZ_GRAB - 1) & ~(LZ_GRAB - 1);
		uint32_t wnext = 0, wend = 0;
		for (;;) {
			/* a few chain steps per round amortise the phase bookkeeping below; a lane
			 * that leaves WALK mode sits out the remaining steps */
#pragma unroll
			for (int u = 0; u < LZ_WALK_STEPS; u++) {
				if (mode == M_WALK) {
					const uint32_t q = S.prev[cur];
					LZ_STAT(if (jdb_emu_lz_steps) jdb_emu_lz_steps[seg0 + p]++;)
					if (q - jmin >= dmax || steps == 0) mode = M_FETCH;          /* not in [jmin, j) */
					else {
						steps--;
						cur = q;
						if (S.data[q + best] == cb) { cq = q; clen = 0; mode = M_COMPARE; }
					}
				}
			}
			/* mode bits: FETCH 00, WALK 01, COMPARE 10, DONE 11 */
			const unsigned bit0 = __ballot_sync(JDB_FULL_MASK, mode & 1u);
			const unsigned bit1 = __ballot_sync(JDB_FULL_MASK, mode & 2u);
			if ((bit0 & bit1) == JDB_FULL_MASK) break;
			if (bit1 & ~bit0) {
				/* COMPARE: 8 more bytes of the candidate (most comparisons end here) */
				if (mode == M_COMPARE) {
					uint32_t len = clen;
					bool done = false;
#pragma unroll
					for (int u = 0; u < LZ_CMP_WORDS; u++) {
						if (!done) {
							if (len >= maxlen) done = true;
							else {
								const uint32_t x = jdb_ld32u(S.data, j + len) ^ jdb_ld32u(S.data, cq + len);
								if (x) { len += (uint32_t) (__ffs((int) x) - 1) >> 3; done = true; }
								else len += 4;
							}
						}
					}
					if (!done && len >= maxlen) done = true;
					if (done) {
						if (len > maxlen) len = maxlen;
						mode = M_WALK;
						if (len > best) {
							best = len;
							bestd = j - cq;
							cb = S.data[j + best];
							if (len >= nice || len == maxlen) mode = M_FETCH;
							/* a match in hand: the rest of the chain gets the reduced budget */
							if (steps > prm.patience) steps = prm.patience;
						}
					} else {
						clen = len;
					}
				}
			}
			const unsigned fetchers = ~(bit0 | bit1) | __ballot_sync(JDB_FULL_MASK, mode == M_FETCH);
			if (fetchers) {
				/* FETCH: positions come from the warp's own range; the counter lives in a
				 * register (the phase is warp-synchronous), ranks from the ballot */
				if (wnext == wend && wend < slim) {
					/* the warp's block is used up: take the next LZ_GRAB positions of the
					 * segment (one shared-memory atomic per block keeps the warps level) */
					uint32_t g = 0;
					if (lane == 0) g = atomicAdd(&S.next_pos, grab);
					g = __shfl_sync(JDB_FULL_MASK, g, 0);
					wnext = g < slim ? g : slim;
					wend = g < slim ? g + grab : slim;
				}
				if (ROUNDS) {
				/* the next positions that need a search: set bits of the need bitmap from
				 * wnext on, within one bitmap word per round */
				uint32_t nbits = 0;
				while (wnext < wend) {
					nbits = S.need[wnext >> 5] & (0xffffffffu << (wnext & 31u));
					if (nbits) break;
					wnext = (wnext | 31u) + 1u;
				}
				const uint32_t rank = (uint32_t) __popc(fetchers & ((1u << lane) - 1u));
				const uint32_t have = (uint32_t) __popc(nbits);
				const uint32_t want = (uint32_t) __popc(fetchers);
				const uint32_t wbase = wnext & ~31u;
				if (have) {
					if (want >= have) wnext = wbase + 32u;
					else wnext = wbase + (uint32_t) __fns(nbits, 0, (int) want + 1);      /* first one not taken */
				}
				if ((fetchers >> lane) & 1u) {
					if (p != 0xffffffffu) S.m[p] = best >= MINLEN ? (best << 16) | bestd : 0;
					p = rank < have ? wbase + (uint32_t) __fns(nbits, 0, (int) rank + 1) : 0xffffffffu;
					if (rank >= have) {
						/* nothing (more) in this word / block: next round, or done when the segment is */
						if (wnext >= wend && wend >= slim) mode = M_DONE;
					}
					else if (p >= seg_len || seg_len - p < MINLEN) {
						/* nothing to find here; stay in FETCH (pass 1 left 0 there) */
						p = 0xffffffffu;
					} else {
						maxlen = seg_len - p;                     /* never past the segment */
						if (maxlen > MAXLEN) maxlen = MAXLEN;
						j = hoff + p;
						jmin = j > WND - 1 ? j - (WND - 1) : 0;
						if (jmin < first_valid) jmin = first_valid;
						dmax = j - jmin;
						{
							/* pass 1 already looked at the first candidate */
							const uint32_t v1 = S.m[p] & ~M_TAKE;
							best = MINLEN - 1; bestd = 0; cur = j; steps = prm.chain;
							if (S.prev[j] - jmin < dmax) {            /* there was a first candidate */
								cur = S.prev[j];
								steps--;
								if (v1) {
									best = v1 >> 16; bestd = v1 & 0xffffu;
									if (best >= nice || best == maxlen) steps = 0;
								}
							}
						}
						cb = S.data[j + best];
						mode = M_WALK;
					}
				}
				} else {
				const uint32_t base = wnext;
				const uint32_t rank = (uint32_t) __popc(fetchers & ((1u << lane) - 1u));
				const uint32_t have = wend - wnext;
				const uint32_t want = (uint32_t) __popc(fetchers);
				wnext += want < have ? want : have;
				if ((fetchers >> lane) & 1u) {
					if (p != 0xffffffffu) S.m[p] = best >= MINLEN ? (best << 16) | bestd : 0;
					p = base + rank;
					if (rank >= have) {
						/* block exhausted: next round (or done when the segment is) */
						if (wend >= slim) mode = M_DONE;
						p = 0xffffffffu;
					}
					else if (p >= seg_len || seg_len - p < MINLEN) {
						/* nothing to find here; stay in FETCH */
						S.m[p] = 0;
						p = 0xffffffffu;
					} else {
						maxlen = seg_len - p;                     /* never past the segment */
						if (maxlen > MAXLEN) maxlen = MAXLEN;
						j = hoff + p;
						jmin = j > WND - 1 ? j - (WND - 1) : 0;
						if (jmin < first_valid) jmin = first_valid;
						dmax = j - jmin;
						best = MINLEN - 1; bestd = 0; cur = j; steps = prm.chain;
						cb = S.data[j + best];
						mode = M_WALK;
					}
				}
				}
			}
		}
	}
	__syncthreads();
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
			const uint32_t p = tid + k * LZ_THREADS;
			nomatch += (p < seg_len && S.m[p] == 0) ? 1u : 0u;
		}
		for (int o = 16; o; o >>= 1) nomatch += __shfl_xor_sync(JDB_FULL_MASK, nomatch, o);
		if ((tid & 31u) == 0) atomicAdd(&S.nomatch, nomatch);
		__syncthreads();
	}
	if (prm.short3 && (prm.short3 > 1 || (S.nomatch * 5u > seg_len * 2u && S.nomatch * 20u < seg_len * 19u))) {
		const uint32_t sv = (chunk0 == 0 && prm.hist_min > hist0) ? (uint32_t) (prm.hist_min - hist0) : 0u;   /* 40 % .. 95 %: not random data either */
		for (uint32_t k = 0; k < PER_THREAD; k++) {
			const uint32_t p = tid + k * LZ_THREADS;
			if (p + 3 > seg_len || S.m[p] != 0) continue;
			const uint32_t j = hoff + p;
			const uint32_t w = jdb_ld32u(S.data, j) & 0xffffffu;
			uint32_t found = 0;
#pragma unroll
			for (int u = 0; u < 12; u++) {
				const uint32_t d = c_short_dist[u];
				if (!found && d <= j && j - d >= sv && (jdb_ld32u(S.data, j - d) & 0xffffffu) == w) found = d;
			}
			if (found) S.m[p] = (3u << 16) | found;
		}
		__syncthreads();
	}

	lz_parse_phase(S, prm, tid, seg_len);

	/* ---- emit: 16 consecutive positions per thread ---- */
	{
		const uint32_t p0 = tid * PER_THREAD;
		const uint32_t w = p0 / WALK_BLOCK;
		const uint32_t mg = S.merge[w];
		uint32_t bits = (S.spec[p0 >> 5] >> (p0 & 31)) & 0xffffu;
		/* speculative positions count only from the merge point on */
		if (mg >= p0 + PER_THREAD) bits = 0;
		else if (mg > p0) bits &= ~((1u << (mg - p0)) - 1u);
		bits |= (S.fix[p0 >> 5] >> (p0 & 31)) & 0xffffu;
		if (p0 >= seg_len) bits = 0;

		const uint32_t cnt = (uint32_t) __popc(bits);
		uint32_t incl = cnt;
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(JDB_FULL_MASK, incl, o);
			if ((int) (tid & 31) >= o) incl += t;
		}
		if ((tid & 31) == 31) S.warp_sum[tid >> 5] = incl;
		__syncthreads();
		if (tid < 32) {
			uint32_t v = S.warp_sum[tid];
			uint32_t iv = v;
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t t = __shfl_up_sync(JDB_FULL_MASK, iv, o);
				if ((int) tid >= o) iv += t;
			}
			S.warp_sum[tid] = iv - v;                    /* exclusive */
			if (tid == 31) seg_ntok[seg] = iv;
		}
		__syncthreads();
		uint32_t o = S.warp_sum[tid >> 5] + incl - cnt;
		uint32_t* out = tok + seg0;
		while (bits) {
			const uint32_t p = p0 + (uint32_t) (__ffs((int) bits) - 1);
			bits &= bits - 1;
			const uint32_t v = S.m[p];
			uint32_t token;
			if (v & M_TAKE) {
				const uint32_t len = (v >> 16) & 0x1ffu, dist = v & 0xffffu;
				token = TOK_MATCH | ((len - 3) << 16) | (dist - 1);
				atomicAdd(&S.hist[257 + len_symbol(len)], 1u);
				atomicAdd(&S.hist[DSYM0 + dist_symbol(dist)], 1u);
			} else {
				token = S.data[hoff + p];
				atomicAdd(&S.hist[token], 1u);
			}
			out[o++] = token;
		}
	}
	__syncthreads();
	for (uint32_t i = tid; i < NSYM; i += LZ_THREADS) seg_hist[(uint64_t) seg * NSYM + i] = S.hist[i];
}

/* ---- launchers ------------------------------------------------------------- */

extern "C" int jdb_lz_chain(const uint8_t* in, uint64_t n, uint32_t chunk_bytes, uint32_t range,
                            const uint32_t* chunk_len, uint16_t* prev, jdb_stream s)
{
	if (n == 0) return JDB_OK;
	const size_t smem = sizeof(ChainSmem);
#ifndef JDB_SIMT_EMU
	static int configured[64];
	int dev = jdb_rt_get_device();
	if (dev >= 0 && dev < 64 && !configured[dev]) {
		cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		configured[dev] = 1;
	}
#endif
	const uint64_t items = (n + range - 1) / range;
	JDB_LAUNCH(chain_kernel, dim3((unsigned) items), dim3(CH_THREADS), smem, s, in, n, chunk_bytes, range, chunk_len, prev);
	return jdb_rt_check_launch("chain_kernel");
}

extern "C" int jdb_lz_parse(const uint8_t* in, uint64_t n, uint32_t chunk_bytes,
                            const uint32_t* chunk_len, const uint16_t* prev,
                            uint32_t good, uint32_t nice, uint32_t chain, uint32_t lazy,
                            uint32_t skip_segs, uint32_t hist_min,
                            uint32_t* tok, uint32_t* seg_ntok, uint32_t* seg_hist, jdb_stream s)
{
	if (n == 0) return JDB_OK;
	const size_t smem = sizeof(LzSmem);
#ifndef JDB_SIMT_EMU
	static int configured[64];
	int dev = jdb_rt_get_device();
	if (dev >= 0 && dev < 64 && !configured[dev]) {
		cudaFuncSetAttribute(lz_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		cudaFuncSetAttribute(lz_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		configured[dev] = 1;
	}
#endif
	LzParams prm;
	prm.good = good; prm.nice = nice; prm.chain = chain; prm.lazy = lazy;
	prm.skip_segs = skip_segs; prm.hist_min = hist_min;
	/* Search only where a tentative parse goes (two refinement rounds) when chains are long:
	 * measured on B200, level 9: LOGS 2.36 -> 5.09 GB/s, TEXT 3.82 -> 4.38 GB/s for +0.3-0.65 %
	 * size; at level 6 (chain 48) the extra parse rounds cost what the search saves on the
	 * mixed corpus, so it stays off there. */
	prm.twophase = getenv("JDB_LZ_TWOPHASE") ? (uint32_t) atoi(getenv("JDB_LZ_TWOPHASE")) : (chain >= 128 ? 2u : 0u);
	prm.short3 = getenv("JDB_LZ_SHORT3") ? (uint32_t) atoi(getenv("JDB_LZ_SHORT3")) : 1u;
	prm.patience = getenv("JDB_LZ_PATIENCE") ? (uint32_t) atoi(getenv("JDB_LZ_PATIENCE")) : chain;
	const uint64_t nseg = (n + SEG - 1) / SEG;
	if (prm.twophase)
		JDB_LAUNCH((lz_kernel<true>), dim3((unsigned) nseg), dim3(LZ_THREADS), smem, s,
		           in, n, chunk_bytes, chunk_len, prev, prm, tok, seg_ntok, seg_hist);
	else
		JDB_LAUNCH((lz_kernel<false>), dim3((unsigned) nseg), dim3(LZ_THREADS), smem, s,
		           in, n, chunk_bytes, chunk_len, prev, prm, tok, seg_ntok, seg_hist);
	return jdb_rt_check_launch("lz_kernel");
}
