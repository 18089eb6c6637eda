/*
 * inflate.cu -- table driven DEFLATE decoder for sm_100a: one warp per stream of a batch,
 * one thread block for a stream on its own.
 *
 * Replaces the reference's serial decoder: buildtable (src/inflator.c:380-568),
 * decodednmc/readlengths (:1029-1190), decodestrd (:930-1019), the hot loop
 * decodefast (:1529-1823) and the resumable decodeblock/copybytes
 * (:1213-1518), plus updatewindow (:616-675) for streaming use.
 *
 * Work decomposition
 *   A DEFLATE stream is bit-serial, so parallelism comes from many streams
 *   (BASELINE config 3: 1 M independent records; our own output: independent
 *   chunks).  inflate_batch_kernel: each warp owns one stream at a time and pulls
 *   the next stream index from a global counter (dynamic load balance for 4-64 KiB
 *   records).  inflate_wide_kernel: ONE stream that cannot be split (anybody's
 *   zlib / gzip output) gets a CTA of 16 warps (see "one stream on a whole CTA").
 *
 *   Per warp, in shared memory: the two-level lookup tables (lit/len root 9 bits,
 *   distance root 8 bits; the reference uses 10 / 8, src/inflator.c:30-32), built
 *   warp-cooperatively per block, a 1 KiB ring of staged input and a 4 KiB ring of
 *   the newest output.  The symbol loop is lane-parallel: every lane decodes its
 *   own subsequence of the bit window, speculatively, and the chain of lanes whose
 *   start is their predecessor's end is the decoded sequence (par_round); anything
 *   a round does not judge goes through the step-by-step decoder, which owns every
 *   status and error decision.  Tokens become bytes 32 at a time (emit_queue).
 *
 *   Algorithmic traffic: C compressed bytes read + N bytes written per stream
 *   (match sources older than the ring are re-read from L2).
 *
 * Error model: the INFLT_* codes of jdeflate/inflator.h with the acceptance
 * rules of the reference (see oracle/jd_oracle.c for the restatement); the two
 * places where the reference accepts an invalid stream (lit/len symbols
 * 286/287 and distance symbols 30/31 of the fixed code, SURVEY / DESIGN.md
 * "deviations") are rejected with INFLT_EBADCODE like zlib does.
 */
#include "common.cuh"
#ifdef WIDE_PROF
#include <stdio.h>
#endif

#ifndef INF_WARPS
#define INF_WARPS        16
#endif
#define INF_THREADS      (INF_WARPS * 32)
#ifndef LIT_ROOT
#define LIT_ROOT         JDB_INF_LIT_ROOT
#endif
#define DIST_ROOT        8
#ifndef LIT_TABLE
#define LIT_TABLE        JDB_INF_LIT_TABLE
#endif
#define DIST_TABLE       JDB_INF_DIST_TABLE
#define QUEUE            32
#ifndef RING
#define RING             4096u     /* per-warp window of the newest output bytes */
#endif
#ifndef MAXBATCH
#define MAXBATCH         1024u     /* a batch stops growing beyond this many bytes  */
#endif
#define RING_KEEP        (RING - MAXBATCH - 258u)
#define INW              256u      /* words of staged input per warp (1 KiB) */

/* lane-parallel symbol decode (see par_round): every lane decodes its own PAR_S-bit
 * subsequence of the input window into PQ_K token slots */
#ifndef PQ_K
#define PQ_K             32u
#endif
#ifndef PAR_S0
#define PAR_S0           192u      /* subsequence length in bits: first guess, then adaptive */
#endif
#define PAR_SMIN         64u
#define PAR_SMAX         192u      /* 32 * PAR_SMAX + 64 bits must fit the input ring with one fill granule to spare */
#ifndef PAR_MAXPASS
#define PAR_MAXPASS      8
#endif
#ifdef LANE_NOINLINE
#define LANE_INLINE __noinline__
#else
#define LANE_INLINE __forceinline__
#endif
#define PF_OK            0u        /* ran to its boundary */
#define PF_EOB           1u        /* decoded the end-of-block symbol */
#define PF_ANOM          2u        /* met something the step-by-step decoder has to judge */
#define PF_FULL          3u        /* token slots used up before the boundary */
#define PF_STOP          4u        /* reached the end of the safely decodable input */

/* table entry: value<<16 | type<<8 | extra<<4 | nbits   (nbits==0: invalid) */
#define T_LIT   0u
#define T_BASE  1u
#define T_EOB   2u
#define T_SUB   3u
#define ENTRY(value, type, extra, nbits) \
	(((uint32_t) (value) << 16) | ((uint32_t) (type) << 8) | ((uint32_t) (extra) << 4) | (uint32_t) (nbits))

/* INFLT_* codes (jdeflate/inflator.h:48-66) */
#define ST_OK        0u
#define ST_SRCEXH    1u
#define ST_TGTEXH    2u
#define ST_ERROR     3u
#define ST_MARKER    4u            /* internal (count mode): stopped after an empty stored block */
#define ST_REDO      0xffffffffu   /* internal: fast path hands the stream to the general decoder */
#define E_BADCODE    2u
#define E_BADTREE    3u
#define E_FAROFFSET  4u
#define E_BADBLOCK   5u
#define E_INPUTEND   6u

__constant__ uint16_t c_len_base[32] = {
	3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
	67, 83, 99, 115, 131, 163, 195, 227, 258, 0, 0, 0
};
__constant__ uint8_t c_len_extra[32] = {
	0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 0, 0, 0
};
__constant__ uint16_t c_dist_base[32] = {
	1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
	1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577, 0, 0
};
__constant__ uint8_t c_dist_extra[32] = {
	0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 0, 0
};
__constant__ uint8_t c_precode_order[19] = {
	16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15
};

/* per-warp shared memory */
/* scratch of the header parser / table builder (one per warp) */
#define HDR_STAGE 640u     /* >= the longest block header: 3 + 14 + 19 * 3 + 316 * 14 bits = 563 bytes */
struct BuildMem {
	uint16_t code[320];
	uint8_t  len[320];
	uint32_t count[16];
	uint32_t next[16];
	uint32_t scratch[8];
	uint8_t  hdr[HDR_STAGE];       /* the bytes of the header, staged by all lanes for lane 0's bit reader */
};

struct WarpMem {
	uint32_t lit[LIT_TABLE];
	uint32_t dist[DIST_TABLE];
	uint32_t inbuf[INW];           /* ring of staged input words (symbol loop) */
	uint32_t pend[2];              /* a match cut by the end of the target: rest length, distance */
	union {
		BuildMem bm;                    /* block header */
		uint32_t cq[PQ_K * 32u];        /* symbol loop: token slot i of lane l at cq[i * 32 + l], then compacted */
	};
	uint8_t  ring[RING];
};

enum { KIND_LIT = 0, KIND_DIST = 1, KIND_PRE = 2 };

/*
 * Build one two-level table from code lengths in m->len[0..n).
 * Returns 0, or 1 when the length set is not acceptable (rules of
 * src/inflator.c:428-474).  All lanes call it; all lanes get the result.
 */
static __device__ __noinline__ int
build_table(BuildMem* m, uint32_t* table, int n, int kind, int lenoff)
{
	const unsigned lane = jdb_lane();
	const int root = kind == KIND_LIT ? LIT_ROOT : kind == KIND_DIST ? DIST_ROOT : 7;
	const int limit = kind == KIND_LIT ? LIT_TABLE : kind == KIND_DIST ? DIST_TABLE : 128;
	const uint8_t* len = m->len + lenoff;

	for (int i = lane; i < limit; i += 32) table[i] = 0;
	if (lane < 16) m->count[lane] = 0;
	__syncwarp();
	for (int i = lane; i < n; i += 32) atomicAdd(&m->count[len[i]], 1u);
	__syncwarp();

	int rc = 0;
	if (lane == 0) {
		if (m->count[0] == (uint32_t) n) {
			rc = kind == KIND_DIST ? 2 : 1;         /* 2: empty distance code is legal */
		} else {
			m->count[0] = 0;
			int mlen = 15;
			while (m->count[mlen] == 0) mlen--;
			int left = 1;
			for (int l = 1; l <= 15; l++) {
				left = (left << 1) - (int) m->count[l];
				if (left < 0) { rc = 1; break; }
			}
			if (rc == 0 && left && !(mlen == 1 && kind == KIND_DIST)) rc = 1;
			if (rc == 0) {
				uint32_t c = 0;
				m->next[0] = 0;
				for (int l = 1; l <= 15; l++) {
					c = (c + m->count[l - 1]) << 1;
					m->next[l] = c;
				}
			}
		}
	}
	rc = __shfl_sync(JDB_FULL_MASK, rc, 0);
	if (rc) return rc == 2 ? 0 : 1;
	/* canonical code of every symbol, bit reversed (LSB-first stream): symbols of one length
	 * take consecutive codes in symbol order -- 32 symbols at a time, a symbol's rank among
	 * the lanes that hold the same length on top of the running count of that length */
	for (int base = 0; base < n; base += 32) {
		const int i = base + (int) lane;
		const int l = i < n ? len[i] : 0;
		const unsigned peers = __match_any_sync(JDB_FULL_MASK, l);
		const uint32_t rank = (uint32_t) __popc(peers & ((1u << lane) - 1u));
		if (l) m->code[i] = (uint16_t) (__brev(m->next[l] + rank) >> (32 - l));
		__syncwarp();
		if (l && rank == 0) m->next[l] += (uint32_t) __popc(peers);
		__syncwarp();
	}

	const uint32_t rootmask = (1u << root) - 1;

	/* pass A: longest code below every root slot that needs a sub-table */
	for (int i = lane; i < n; i += 32) {
		int l = len[i];
		if (l > root) atomicMax(&table[m->code[i] & rootmask], (uint32_t) l);
	}
	__syncwarp();
	/* pass B: lay the sub-tables out in root order (warp scan of their sizes) */
	{
		const int per = (1 << root) / 32;
		uint32_t mine = 0;
		for (int k = 0; k < per; k++) {
			uint32_t v = table[lane * per + k];
			if (v) mine += 1u << (v - root);
		}
		uint32_t incl = mine;
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(JDB_FULL_MASK, incl, o);
			if ((int) lane >= o) incl += t;
		}
		uint32_t off = (1u << root) + incl - mine;
		uint32_t total = __shfl_sync(JDB_FULL_MASK, incl, 31);
		if ((1u << root) + total > (uint32_t) limit) return 1;
		for (int k = 0; k < per; k++) {
			uint32_t v = table[lane * per + k];
			if (v) {
				table[lane * per + k] = ENTRY(off, T_SUB, v - root, root);
				off += 1u << (v - root);
			}
		}
	}
	__syncwarp();
	/* pass C: replicate every symbol over the slots whose low bits match */
	for (int i = lane; i < n; i += 32) {
		int l = len[i];
		if (l == 0) continue;
		uint32_t e;
		if (kind == KIND_LIT) {
			if (i < 256) e = ENTRY(i, T_LIT, 0, 0);
			else if (i == 256) e = ENTRY(0, T_EOB, 0, 0);
			else if (i <= 285) e = ENTRY(c_len_base[i - 257], T_BASE, c_len_extra[i - 257], 0);
			else e = ENTRY(0, T_BASE, 0, 0);   /* 286/287: reserved, base 0 marks them invalid */
		} else if (kind == KIND_DIST) {
			/* 30/31: reserved, base 0 marks them invalid */
			e = i > 29 ? ENTRY(0, T_BASE, 0, 0) : ENTRY(c_dist_base[i], T_BASE, c_dist_extra[i], 0);
		} else {
			e = ENTRY(i, T_LIT, 0, 0);
		}
		uint32_t code = m->code[i];
		if (l <= root) {
			e |= (uint32_t) l;
			for (uint32_t s = code; s < (1u << root); s += 1u << l) table[s] = e;
		} else {
			uint32_t link = table[code & rootmask];
			uint32_t base = link >> 16, sbits = (link >> 4) & 15u;
			e |= (uint32_t) l;
			for (uint32_t s = code >> root; s < (1u << sbits); s += 1u << (l - root)) table[base + s] = e;
		}
	}
	__syncwarp();
	return 0;
}

/* ---- bit reader (meaningful on lane 0 only) -------------------------------- */

struct Bits {
	uint64_t bb;
	uint32_t bc;
	const uint8_t* p;       /* next unread byte */
	const uint8_t* end;
};

/* byte-wise refill up to `need` bits; 0 when the input ends first */
static __device__ __forceinline__ int
bits_need(Bits& b, uint32_t need)
{
	while (b.bc < need) {
		if (b.p >= b.end) return 0;
		b.bb |= (uint64_t) (*b.p++) << b.bc;
		b.bc += 8;
	}
	return 1;
}

/* fast refill: p is 4-byte aligned and at least 4 bytes remain */
static __device__ __forceinline__ void
bits_refill32(Bits& b)
{
	if (b.bc <= 32) {
		b.bb |= (uint64_t) (*(const uint32_t*) b.p) << b.bc;
		b.p += 4;
		b.bc += 32;
	}
}

static __device__ __forceinline__ uint32_t
bits_take(Bits& b, uint32_t k)
{
	uint32_t v = (uint32_t) b.bb & ((1u << k) - 1u);
	b.bb >>= k;
	b.bc -= k;
	return v;
}

static __device__ __forceinline__ uint32_t
lookup(const uint32_t* table, uint64_t bb, int root)
{
	uint32_t e = table[(uint32_t) bb & ((1u << root) - 1u)];
	if (((e >> 8) & 3u) == T_SUB) {
		uint32_t sbits = (e >> 4) & 15u;
		e = table[(e >> 16) + (((uint32_t) (bb >> root)) & ((1u << sbits) - 1u))];
	}
	return e;
}

/* 32 bits starting at bit offset `o` of the input ring */
static __device__ __forceinline__ uint32_t
peek32(const uint32_t* ring, uint32_t o)
{
	const uint32_t i = o >> 5;
	return __funnelshift_r(ring[i & (INW - 1u)], ring[(i + 1u) & (INW - 1u)], o & 31u);
}

static __device__ __forceinline__ uint32_t
lookup32(const uint32_t* table, uint32_t bits, int root)
{
	uint32_t e = table[bits & ((1u << root) - 1u)];
	if (((e >> 8) & 3u) == T_SUB) {
		const uint32_t sbits = (e >> 4) & 15u;
		e = table[(e >> 16) + ((bits >> root) & ((1u << sbits) - 1u))];
	}
	return e;
}

/*
 * Parse one block header at the bit position of `b` (meaningful on lane 0) and,
 * for Huffman coded blocks, build the two decoding tables into `lit` / `dist`.
 * All lanes call it.  Returns 0, 1 when the input ends inside the header (the
 * caller rewinds), or an INFLT error code + 1.  `type` and `lb` (last block)
 * come back on all lanes; a stored block's length is left in bm->scratch[0].
 */
static __device__ uint32_t
parse_block_header(BuildMem* bm, uint32_t* lit, uint32_t* dist, Bits& gb_, uint32_t& type_out, uint32_t& lb_out)
{
	const unsigned lane = jdb_lane();
	uint32_t r = 0;        /* 0 ok, 1 starved, 2+ : INFLT error code + 1 */
	uint32_t type = 0, lb = 0, hlit = 0, hdist = 0;
	/* lane 0 reads the header bit by bit: from shared memory, where all lanes put the next
	 * HDR_STAGE bytes of the input first (no header is longer), not byte by byte from L2 */
	Bits& gb = gb_;
	const uint8_t* const gp = gb.p;
	{
		const uint64_t left = (uint64_t) (gb.end - gb.p);
		const uint32_t gn = left < HDR_STAGE ? (uint32_t) left : HDR_STAGE;
		for (uint32_t i = lane; i < gn; i += 32) bm->hdr[i] = gp[i];
		__syncwarp();
	}
	Bits b = gb;
	b.p = bm->hdr;
	b.end = bm->hdr + ((uint64_t) (gb.end - gb.p) < HDR_STAGE ? (uint64_t) (gb.end - gb.p) : (uint64_t) HDR_STAGE);
	if (lane == 0) {
		if (!bits_need(b, 3)) r = 1;
		else {
			lb = bits_take(b, 1);
			type = bits_take(b, 2);
			if (type == 0) {
				/* stored: src/inflator.c:930-1019 */
				bits_take(b, b.bc & 7u);
				if (!bits_need(b, 32)) r = 1;
				else {
					uint32_t l = bits_take(b, 16), nl = bits_take(b, 16);
					if ((l ^ nl) != 0xffffu) r = 1 + E_BADBLOCK;
					else bm->scratch[0] = l;
				}
			} else if (type == 3) {
				r = 1 + E_BADBLOCK;                 /* src/inflator.c:888 */
			} else if (type == 2) {
				/* dynamic header: src/inflator.c:1103-1190 */
				if (!bits_need(b, 14)) r = 1;
				else {
					hlit = bits_take(b, 5) + 257;
					hdist = bits_take(b, 5) + 1;
					uint32_t hclen = bits_take(b, 4) + 4;
					if (hlit > 286 || hdist > 30) r = 1 + E_BADTREE;
					else {
						for (int i = 0; i < 19; i++) bm->len[i] = 0;
						for (uint32_t i = 0; i < hclen; i++) {
							if (!bits_need(b, 3)) { r = 1; break; }
							bm->len[c_precode_order[i]] = (uint8_t) bits_take(b, 3);
						}
					}
				}
			}
		}
	}
	r = __shfl_sync(JDB_FULL_MASK, r, 0);
	type = __shfl_sync(JDB_FULL_MASK, type, 0);
	lb = __shfl_sync(JDB_FULL_MASK, lb, 0);
	hlit = __shfl_sync(JDB_FULL_MASK, hlit, 0);
	hdist = __shfl_sync(JDB_FULL_MASK, hdist, 0);
	__syncwarp();

	if (r == 0 && type == 1) {
		/* fixed code, RFC 1951 3.2.6 (reference tables src/inflator.c:1840-2164) */
		for (int i = lane; i < 288; i += 32)
			bm->len[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
		bm->len[288 + lane] = 5;
		__syncwarp();
		build_table(bm, lit, 288, KIND_LIT, 0);
		build_table(bm, dist, 32, KIND_DIST, 288);
	}
	if (r == 0 && type == 2) {
		if (build_table(bm, lit, 19, KIND_PRE, 0)) r = 1 + E_BADTREE;
		if (r == 0) {
			if (lane == 0) {
				/* readlengths: src/inflator.c:1029-1101; the precode table
				 * occupies lit[0..128) */
				uint32_t i = 0;
				const uint32_t n = hlit + hdist;
				while (i < n) {
					/* (four bytes of the staged header at a time; a refill per byte is a dependent
					 * load in front of every code length) */
					if (b.bc <= 32 && b.p + 4 <= b.end) {
						const uint32_t v = (uint32_t) b.p[0] | ((uint32_t) b.p[1] << 8) | ((uint32_t) b.p[2] << 16) | ((uint32_t) b.p[3] << 24);
						b.bb |= (uint64_t) v << b.bc;
						b.p += 4;
						b.bc += 32;
					}
					bits_need(b, 7);
					uint32_t e = lit[(uint32_t) b.bb & 127u];
					uint32_t nb = e & 15u;
					if (nb == 0) { r = b.bc >= 7 ? 1 + E_BADCODE : 1; break; }
					if (nb > b.bc) { r = 1; break; }
					uint32_t sym = e >> 16;
					if (sym < 16) {
						bits_take(b, nb);
						bm->len[i++] = (uint8_t) sym;
						continue;
					}
					uint32_t xb = sym == 16 ? 2u : sym == 17 ? 3u : 7u;
					if (!bits_need(b, nb + xb)) { r = 1; break; }
					bits_take(b, nb);
					uint32_t rep = (sym == 18 ? 11u : 3u) + bits_take(b, xb);
					uint32_t val = 0;
					if (sym == 16) {
						if (i == 0) { r = 1 + E_BADTREE; break; }
						val = bm->len[i - 1];
					}
					/* the reference bounds runs by its array size, not by
					 * hlit + hdist: src/inflator.c:1090-1093 */
					if (i + rep > 320) { r = 1 + E_BADTREE; break; }
					while (rep--) bm->len[i++] = (uint8_t) val;
				}
				if (r == 0 && bm->len[256] == 0) r = 1 + E_BADTREE;   /* :1171-1174 */
			}
			r = __shfl_sync(JDB_FULL_MASK, r, 0);
			__syncwarp();
		}
		if (r == 0) {
			if (build_table(bm, lit, (int) hlit, KIND_LIT, 0) ||
			    build_table(bm, dist, (int) hdist, KIND_DIST, (int) hlit))
				r = 1 + E_BADTREE;
		}
	}
	gb.bb = b.bb;
	gb.bc = b.bc;
	gb.p = gp + (b.p - bm->hdr);
	type_out = type;
	lb_out = lb;
	return r;
}

/* ---- the per-stream decoder -------------------------------------------------- */

struct Stream {
	/* inputs */
	const uint8_t* src;
	uint64_t src_len;
	uint8_t* dst;
	uint64_t dst_cap;
	int final;
	int count_only;             /* measure a chunk: no output, stop after an empty stored block */
	int stop_marker;            /* decode ONE chunk: stop after an empty stored block */
	jdb_inflate_state* st;      /* NULL in batch mode */
	/* running */
	uint64_t out;               /* bytes written in this call */
	uint64_t consumed;          /* source bytes used by this call */
	uint64_t hist_avail;        /* bytes available before dst[0] (history / dictionary) */
	uint64_t total_before;      /* absolute output position of dst[0] (ring index base) */
	uint8_t* ring;              /* the newest output bytes, indexed by the low bits of their target address */
	int64_t  ring_lo;           /* positions >= ring_lo are served from the ring */
	uint32_t rbase;             /* low bits of dst */
	uint64_t flushed;           /* bytes [0, flushed) have reached the target */
	uint32_t status, error;
};

/*
 * Output path.  Decoded bytes go to a per-warp ring in shared memory first and from
 * there to the target in aligned 32-bit words (flush_ring): the symbol loop produces bytes
 * one or a few at a time per lane, which as global stores are one memory instruction per
 * byte; the ring takes them at shared-memory cost, serves most match sources (the newest
 * bytes) and leaves the target with coalesced word stores.  The ring is indexed by the low
 * bits of the target ADDRESS, so its words line up with the target's words whatever the
 * alignment of dst.  Positions are relative to this call's dst; negative ones are the
 * history of earlier calls.
 */
static __device__ __forceinline__ uint32_t
ring_index(const Stream& s, uint32_t pos)
{
	return (s.rbase + pos) & (RING - 1u);
}

/* byte at output position `pos`: the ring has the newest bytes, L2 the older ones (they were
 * flushed by other lanes of this warp: ld.global.cg), earlier calls' bytes sit in the history ring */
static __device__ __forceinline__ uint8_t
out_byte(const Stream& s, int64_t pos)
{
	if (pos >= s.ring_lo) return s.ring[ring_index(s, (uint32_t) pos)];
	if (pos >= 0) return __ldcg(s.dst + pos);
	uint64_t abs = s.total_before + (uint64_t) pos;
	return __ldcg(s.st->history + (abs & (JDB_INFLATE_HISTORY - 1)));
}

/* all lanes: bytes [s.flushed, upto) of the ring -> target; without `all` only up to the last
 * word boundary (the rest follows with the next flush) */
static __device__ __forceinline__ void
flush_ring(Stream& s, uint64_t upto, bool all)
{
	const unsigned lane = jdb_lane();
	uint8_t* const a = s.dst + s.flushed;
	uint8_t* b = s.dst + upto;
	if (!all) b = (uint8_t*) ((uintptr_t) b & ~(uintptr_t) 3);
	if (b <= a) return;
	uint8_t* w0 = (uint8_t*) (((uintptr_t) a + 3) & ~(uintptr_t) 3);
	if (w0 > b) w0 = b;
	uint8_t* const w1 = w0 + ((size_t) (b - w0) & ~(size_t) 3);
	if (a + lane < w0) a[lane] = s.ring[(uint32_t) (uintptr_t) (a + lane) & (RING - 1u)];
	for (uint8_t* g = w0 + 4u * lane; g < w1; g += 128)
		*(uint32_t*) g = *(const uint32_t*) (s.ring + ((uint32_t) (uintptr_t) g & (RING - 1u)));
	if (w1 + lane < b) w1[lane] = s.ring[(uint32_t) (uintptr_t) (w1 + lane) & (RING - 1u)];
	s.flushed = (uint64_t) (b - s.dst);
	__syncwarp();
}

/*
 * All lanes: turn up to `nq` (<= 32) queued symbols (literal: byte value; match: len << 16 |
 * dist) of stream `s` into output bytes starting at s.out; returns how many it took -- a
 * group stops growing once it holds MAXBATCH bytes (the ring holds one group plus at least
 * RING_KEEP older bytes).  A warp prefix sum gives every symbol its offset; literals and
 * short matches whose source is older than the group are written by their own lane, long
 * matches and matches that read bytes produced by the same group are copied by all lanes,
 * in order.  A match cut by the end of the target is reported through pend_len / pend_dist.
 */
static __device__ __forceinline__ uint32_t
emit_queue(Stream& s, uint32_t* pend, const uint32_t* queue, uint32_t nq, uint32_t& pend_len, uint32_t& pend_dist)
{
	const unsigned lane = jdb_lane();
	uint32_t q = lane < nq ? queue[lane] : 0;
	uint32_t len = lane < nq ? ((q >> 16) ? (q >> 16) : 1u) : 0u;
	uint32_t incl = len;
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(JDB_FULL_MASK, incl, o);
		if ((int) lane >= o) incl += t;
	}
	/* symbols that would start at or beyond MAXBATCH bytes wait for the next group */
	{
		const unsigned over = __ballot_sync(JDB_FULL_MASK, lane < nq && incl - len >= MAXBATCH);
		if (over) {
			nq = (uint32_t) __ffs(over) - 1u;
			if (lane >= nq) { q = 0; len = 0; }
		}
	}
	const bool is_match = lane < nq && (q >> 16) != 0;
	const uint32_t dist = q & 0xffffu;
	const uint64_t base = s.out;
	const uint32_t total = __shfl_sync(JDB_FULL_MASK, incl, (int) nq - 1);
	const uint32_t off = incl - len;                     /* of this symbol, from `base` */
	/* clip the last symbol to the target capacity */
	const uint64_t room = s.dst_cap - base;
	uint32_t emit = len;
	if ((uint64_t) off + len > room) emit = (uint64_t) off < room ? (uint32_t) (room - off) : 0u;
	if (lane < nq && emit < len) {
		pend[0] = len - emit;
		pend[1] = dist;
	}
	const uint64_t done = (uint64_t) total > room ? s.dst_cap : base + total;
	/* the ring will hold [done - RING, done): what is older comes from L2 */
	{
		const int64_t lo = (int64_t) done - (int64_t) RING;
		if (lo > s.ring_lo) s.ring_lo = lo;
	}
	const int64_t src = (int64_t) base + (int64_t) off - (int64_t) dist;
	const bool dependent = is_match && (src + (int64_t) len > (int64_t) base || dist < len);
	const bool longm = is_match && !dependent && emit > 16;
	const uint32_t wi = ring_index(s, (uint32_t) base + off);
	if (lane < nq && !is_match && emit) s.ring[wi] = (uint8_t) q;
	if (is_match && !dependent && !longm) {
		/* short match from older bytes */
		if (src >= s.ring_lo) {
			const uint32_t ri = ring_index(s, (uint32_t) src);
			if (wi + emit <= RING && ri + emit <= RING) {
				/* neither range wraps around the ring (nearly always): plain pointers */
				const uint8_t* const rp = s.ring + ri;
				uint8_t* const wp = s.ring + wi;
				for (uint32_t j = 0; j < emit; j++) wp[j] = rp[j];
			} else {
				for (uint32_t j = 0; j < emit; j++) s.ring[(wi + j) & (RING - 1u)] = s.ring[(ri + j) & (RING - 1u)];
			}
		} else if (src >= 0 && src + (int64_t) emit <= s.ring_lo) {
			/* all of it in L2: the (up to five) aligned words that hold the source bytes are
			 * loaded together -- one round trip instead of one per byte */
			const uint8_t* const g = s.dst + src;
			const uint32_t* const gw = (const uint32_t*) ((uintptr_t) g & ~(uintptr_t) 3);
			const uint32_t sh = (uint32_t) ((uintptr_t) g & 3u);
			const uint32_t nw = (sh + emit + 3u) >> 2;
			uint32_t w[5];
#pragma unroll
			for (uint32_t k = 0; k < 5; k++) w[k] = k < nw ? __ldcg(gw + k) : 0u;
			uint32_t x[4];
#pragma unroll
			for (uint32_t k = 0; k < 4; k++) x[k] = __funnelshift_r(w[k], w[k + 1], 8u * sh);
#pragma unroll
			for (uint32_t j = 0; j < 16; j++)
				if (j < emit) s.ring[(wi + j) & (RING - 1u)] = (uint8_t) (x[j >> 2] >> (8u * (j & 3u)));
		} else {
			for (uint32_t j = 0; j < emit; j++) s.ring[(wi + j) & (RING - 1u)] = out_byte(s, src + j);
		}
	}
	/* cooperative copies: the long ones, then -- in order -- the ones that depend on this group */
	unsigned lm = __ballot_sync(JDB_FULL_MASK, longm);
	unsigned dm = __ballot_sync(JDB_FULL_MASK, dependent && emit);
	__syncwarp();
	while (lm | dm) {
		int from;
		if (lm) { from = __ffs(lm) - 1; lm &= lm - 1; }
		else { from = __ffs(dm) - 1; dm &= dm - 1; }
		const uint32_t w2 = __shfl_sync(JDB_FULL_MASK, wi, from);
		const uint32_t l2 = __shfl_sync(JDB_FULL_MASK, emit, from);
		const uint32_t d2 = __shfl_sync(JDB_FULL_MASK, dist, from);
		const int64_t s2 = __shfl_sync(JDB_FULL_MASK, (long long) src, from);
		/* every source byte lies before the match (k < dist) */
		for (uint32_t j = lane; j < l2; j += 32) {
			const uint32_t k = d2 >= l2 ? j : j % d2;
			s.ring[(w2 + j) & (RING - 1u)] = out_byte(s, s2 + k);
		}
		__syncwarp();
	}
	if ((uint64_t) total > room) {
		pend_len = pend[0];
		pend_dist = pend[1];
	}
	s.out = done;
	flush_ring(s, done, false);
	return nq;
}

/*
 * One lane's share of a lane-parallel round: decode symbols from bit `start` until the
 * position reaches `limit` (symbols START before the limit; the last one may end behind it),
 * writing tokens to this lane's slots.  Purely speculative: it judges nothing.  Whatever
 * is not a plain literal / length+distance / end-of-block stops the lane with PF_ANOM at
 * the position of that symbol, and the step-by-step decoder decides what it means.
 * The caller guarantees 64 readable bits behind `limit`.
 */
struct LaneRun {
	uint32_t start, end, n, bytes, flag;
	int32_t  need;          /* max over matches of (distance - bytes this lane produced before it) */
};

static __device__ LANE_INLINE void
lane_decode(const uint32_t* inbuf, const uint32_t* lit, const uint32_t* dtab, uint32_t* slots, LaneRun& r,
            uint32_t start, uint32_t limit, uint32_t safe_end)
{
	uint32_t pos = start, n = 0, bytes = 0, flag = PF_OK;
	int32_t need = 0;
	while (pos < limit) {
		if (n >= PQ_K) { flag = PF_FULL; break; }
		const uint32_t bits = peek32(inbuf, pos);
		const uint32_t e = lookup32(lit, bits, LIT_ROOT);
		const uint32_t nb = e & 15u;
		const uint32_t type = (e >> 8) & 3u;
		uint32_t tok;
		if (nb == 0) { flag = PF_ANOM; break; }
		if (type == T_LIT) {
			tok = e >> 16;
			pos += nb;
			bytes += 1;
		} else if (type == T_BASE) {
			if ((e >> 16) == 0) { flag = PF_ANOM; break; }
			const uint32_t lxb = (e >> 4) & 15u;
			const uint32_t len = (e >> 16) + ((bits >> nb) & ((1u << lxb) - 1u));
			const uint32_t o2 = pos + nb + lxb;
			const uint32_t bits2 = peek32(inbuf, o2);
			const uint32_t d = lookup32(dtab, bits2, DIST_ROOT);
			const uint32_t dnb = d & 15u, dxb = (d >> 4) & 15u;
			if (dnb == 0 || (d >> 16) == 0) { flag = PF_ANOM; break; }
			const uint32_t dist = (d >> 16) + ((bits2 >> dnb) & ((1u << dxb) - 1u));
			const int32_t nd = (int32_t) dist - (int32_t) bytes;
			if (nd > need) need = nd;
			tok = (len << 16) | dist;
			pos = o2 + dnb + dxb;
			bytes += len;
		} else {
			/* T_EOB (T_SUB never comes out of a two-level look-up) */
			pos += nb;
			flag = PF_EOB;
			break;
		}
		slots[n * 32u] = tok;
		n++;
	}
	if (flag == PF_OK && limit == safe_end) flag = PF_STOP;
	r.start = start; r.end = pos; r.n = n; r.bytes = bytes; r.flag = flag; r.need = need;
}

/* ---- one stream on a whole CTA (inflate_wide_kernel) ---------------------------
 *
 * One warp per stream is the right shape for a batch of records; a caller with ONE stream
 * that carries no chunk markers (anybody's zlib / gzip output through inflator_inflate)
 * leaves a lone warp to the dependent latency of every instruction it issues.  The wide
 * kernel gives such a stream a CTA of WIDE_WARPS warps.  Warp 0 (the master) runs
 * inflate_stream as it is -- block headers, tables, stored blocks, the step-by-step decoder
 * and with it every status and error decision -- and the other warps wait for its commands.
 * What changes is the lane-parallel round: it runs on WIDE_LANES lanes instead of 32 (same
 * speculative subsequences, same chain of agreeing lanes, now linked across warps through
 * shared memory), and the tokens of the chain are turned into bytes by all lanes at once:
 *
 *   expand   every lane of the chain writes one 16-bit entry per output byte of its tokens
 *            into E[]: the byte itself (literals; match bytes whose source lies before the
 *            round, read from H, a shared-memory mirror of the last 32 KiB of output), or
 *            the E-index of its source byte (match bytes whose source is produced by the
 *            same round -- by any lane)
 *   resolve  pointer jumping, E[x] = E[E[x]] until every entry is a byte: the dependencies
 *            between the matches of a round are followed by all threads at once in
 *            O(log depth) sweeps instead of in token order (in place: an entry only ever
 *            moves towards its source, so a racing reader sees an older but valid link)
 *   write    four entries -> one aligned 32-bit word to the target and to H
 *
 * (reference: decodefast, src/inflator.c:1529-1823 -- one symbol after the other on one core)
 */
#ifndef WIDE_WARPS
#define WIDE_WARPS       16
#endif
#ifndef WIDE_LEAD
#define WIDE_LEAD        0u         /* bits a lane decodes in front of its subsequence before its first guess; 0 = off:
                                    * measured, a lead-in is pass 1 under another name; experiment switch */
#endif
#ifndef WIDE_MAXPASS
#define WIDE_MAXPASS     12
#endif
#define WIDE_LANES       (WIDE_WARPS * 32)
#define WIDE_E           32768u     /* entries: indices need 15 bits */
#define WIDE_CAP         (WIDE_E - 8u)
#define WIDE_H           32768u
#ifndef WIDE_S0
#define WIDE_S0          160u
#endif
#define WIDE_SMIN         PQ_K       /* bits: a subsequence this short never runs out of token slots (a code has >= 1 bit) */
#define WIDE_COPY_MIN     4096u      /* stored bytes from which the whole CTA copies */
#define WIDE_MIN_ROOM    2048u      /* target room below which a round of the whole CTA is not worth its barriers */
#define E_BYTE           0x8000u
#ifndef WIDE_LONG
#define WIDE_LONG        24u        /* longer matches are expanded by all lanes of a warp */
#endif

/*
 * Flattened tables of the wide rounds (built by all threads of the CTA whenever the master has
 * new block tables, wide_flatten): one look-up, WIDE_LROOT / WIDE_DROOT bits wide, answers every
 * code of up to that many bits -- a literal or a length base in the entry format of the two-level
 * tables -- and everything else (longer codes, end of block, codes that do not exist, reserved
 * symbols) is W_SPECIAL: the two-level tables and the code that judges them take over.
 */
#define WIDE_LROOT       11
#define WIDE_DROOT       10
#define W_SPECIAL        0x8000u
#define W_MATCH          0x0100u    /* = T_BASE << 8 */

/*
 * lane_decode for the wide rounds, where the latency of ONE warp's pass is what a round waits
 * for.  Called by all lanes of a warp (`run`: this lane takes part).  Differences to lane_decode:
 *  - the three input words around the position stay in registers (a peek is one funnel shift,
 *    the ring is read when the position enters the next word);
 *  - a symbol is straight-line code for literals and matches alike: one look-up in the flattened
 *    lit/len table, the distance look-up done whether or not it is needed, selects instead of
 *    branches -- a warp whose lanes hold literals and matches runs ONE path per trip, not both one
 *    after the other (measured: the dependent instructions of a trip, not the memory, are what a
 *    pass costs);
 *  - the loop is closed by a vote, one symbol per trip for every lane still at work: a loop
 *    that lanes leave one by one through breaks was compiled without a reconvergence point by some
 *    builds, after which the lanes of a warp ran its trips in ever smaller groups (3x the time).
 */
static __device__ __forceinline__ void
lane_decode_w(const uint32_t* inbuf, const uint32_t* lit, const uint32_t* dtab, const uint32_t* wlit, const uint32_t* wdist,
              uint32_t* slots, LaneRun& r, bool run, uint32_t start, uint32_t limit, uint32_t safe_end)
{
	uint32_t pos = start, n = 0, bytes = 0, flag = PF_OK;
	int32_t need = 0;
	uint32_t* slot = slots;
	uint32_t cur = pos >> 5;
	uint32_t w0 = inbuf[cur & (INW - 1u)], w1 = inbuf[(cur + 1u) & (INW - 1u)], w2 = inbuf[(cur + 2u) & (INW - 1u)];
	bool act = run && pos < limit;
	while (__any_sync(JDB_FULL_MASK, act)) {
		if (act) {
			if (n >= PQ_K) {
				flag = PF_FULL;
				act = false;
			} else {
				const uint32_t i = pos >> 5;
				if (i != cur) {
					if (i == cur + 1u) { w0 = w1; w1 = w2; }
					else { w0 = inbuf[i & (INW - 1u)]; w1 = inbuf[(i + 1u) & (INW - 1u)]; }
					w2 = inbuf[(i + 2u) & (INW - 1u)];
					cur = i;
				}
				const uint32_t bits = __funnelshift_r(w0, w1, pos);
				const uint32_t e = wlit[bits & ((1u << WIDE_LROOT) - 1u)];
				const uint32_t nb = e & 15u, lxb = (e >> 4) & 15u;
				const uint32_t val = (e >> 16) + ((bits >> nb) & ((1u << lxb) - 1u));      /* the literal (lxb = 0), or the length */
				const uint32_t o2 = pos + nb + lxb;                                        /* in this word or the next */
				const uint32_t bits2 = (o2 >> 5) == cur ? __funnelshift_r(w0, w1, o2) : __funnelshift_r(w1, w2, o2);
				const uint32_t d = wdist[bits2 & ((1u << WIDE_DROOT) - 1u)];
				const bool m = (e & W_MATCH) != 0;
				uint32_t tok;
				if ((e | (m ? d : 0u)) & W_SPECIAL) {
					/* ---- the two-level tables decide (lane_decode's body) ---- */
					const uint32_t e1 = lookup32(lit, bits, LIT_ROOT);
					const uint32_t nb1 = e1 & 15u;
					const uint32_t type = (e1 >> 8) & 3u;
					tok = 0;
					if (nb1 == 0) {
						flag = PF_ANOM;
						act = false;
					} else if (type == T_LIT) {
						tok = e1 >> 16;
						pos += nb1;
						bytes += 1;
					} else if (type == T_BASE) {
						const uint32_t lxb1 = (e1 >> 4) & 15u;
						const uint32_t len = (e1 >> 16) + ((bits >> nb1) & ((1u << lxb1) - 1u));
						const uint32_t o3 = pos + nb1 + lxb1;
						const uint32_t bits3 = (o3 >> 5) == cur ? __funnelshift_r(w0, w1, o3) : __funnelshift_r(w1, w2, o3);
						const uint32_t d1 = lookup32(dtab, bits3, DIST_ROOT);
						const uint32_t dnb = d1 & 15u, dxb = (d1 >> 4) & 15u;
						if ((e1 >> 16) == 0 || dnb == 0 || (d1 >> 16) == 0) {
							flag = PF_ANOM;
							act = false;
						} else {
							const uint32_t dist = (d1 >> 16) + ((bits3 >> dnb) & ((1u << dxb) - 1u));
							const int32_t nd = (int32_t) dist - (int32_t) bytes;
							if (nd > need) need = nd;
							tok = (len << 16) | dist;
							pos = o3 + dnb + dxb;
							bytes += len;
						}
					} else {
						pos += nb1;
						flag = PF_EOB;
						act = false;
					}
				} else {
					const uint32_t dnb = d & 15u, dxb = (d >> 4) & 15u;
					const uint32_t dist = (d >> 16) + ((bits2 >> dnb) & ((1u << dxb) - 1u));
					const int32_t nd = m ? (int32_t) dist - (int32_t) bytes : 0;
					if (nd > need) need = nd;
					tok = m ? (val << 16) | dist : val;
					pos = m ? o2 + dnb + dxb : o2;
					bytes += m ? val : 1u;
				}
				if (act) {
					*slot = tok;
					slot += 32;
					n++;
					act = pos < limit;
				}
			}
		}
	}
	if (run) {
		if (flag == PF_OK && limit == safe_end) flag = PF_STOP;
		r.start = start; r.end = pos; r.n = n; r.bytes = bytes; r.flag = flag; r.need = need;
	}
}

/*
 * Lead-in of a speculative lane: walk symbols from `start` without recording anything and return
 * the first symbol boundary at or behind `boundary`.  A lane that has already decoded a stretch in
 * front of its subsequence has most likely fallen into step with the true symbol sequence by the
 * time it reaches it, so that its first guess links with its predecessor.  Anything unusual on
 * the way: give up and guess the boundary itself.
 */
static __device__ __forceinline__ uint32_t
lane_lead_in(const uint32_t* inbuf, const uint32_t* lit, const uint32_t* dtab, uint32_t start, uint32_t boundary)
{
	uint32_t pos = start;
	while (pos < boundary) {
		const uint32_t bits = peek32(inbuf, pos);
		const uint32_t e = lookup32(lit, bits, LIT_ROOT);
		const uint32_t nb = e & 15u;
		const uint32_t type = (e >> 8) & 3u;
		if (nb == 0) return boundary;
		if (type == T_LIT) { pos += nb; continue; }
		if (type != T_BASE || (e >> 16) == 0) return boundary;
		const uint32_t o2 = pos + nb + ((e >> 4) & 15u);
		const uint32_t d = lookup32(dtab, peek32(inbuf, o2), DIST_ROOT);
		if ((d & 15u) == 0 || (d >> 16) == 0) return boundary;
		pos = o2 + (d & 15u) + ((d >> 4) & 15u);
	}
	return pos;
}

struct WideMem {
	/* command block: written by the master before the command barrier */
	uint32_t cmd;                   /* 0: leave, 1: round, 2: copy (stored block), 3: Adler-32 of the output */
	uint64_t ad_n;                  /* Adler-32: bytes at cp_src; partial sums per warp, result */
	unsigned long long ad_a[WIDE_WARPS], ad_b[WIDE_WARPS];
	uint32_t ad_result;
	const uint8_t* cp_src;          /* copy: source, target, bytes */
	uint8_t* cp_dst;
	uint32_t cp_n;
	uint32_t o, S, endbit, nwords, pre, c, prefix_word;
	int32_t  delta;
	uint32_t room, reach;
	const uint32_t* wbase;
	uint8_t* dst;                   /* target of this call */
	uint64_t out;                   /* bytes of this call in front of the round */
	const uint8_t* history;         /* ring of earlier calls' output, or NULL */
	uint64_t total_before;
	uint64_t hist_avail;
	/* H holds the output positions [h_hi - WIDE_H, h_hi) of this call once h_valid */
	uint64_t h_hi;
	uint32_t h_valid;
	/* results of a round */
	uint32_t newo, nbytes, lastflag, stepwise, newS;
	/* scratch */
	uint32_t wend[WIDE_WARPS], wflag[WIDE_WARPS], wbrk[WIDE_WARPS], wsb[WIDE_WARPS], wsn[WIDE_WARPS];
	uint32_t whard[WIDE_WARPS], wcap[WIDE_WARPS], wfull[WIDE_WARPS], wmaxn[WIDE_WARPS];
	uint32_t tables_new;            /* the master has new block tables: flatten them first */
	uint32_t wlit[1u << WIDE_LROOT];
	uint32_t wdist[1u << WIDE_DROOT];
	uint32_t inbuf[WIDE_WARPS][INW];
	uint32_t slots[WIDE_WARPS][PQ_K * 32u];
	__align__(8) uint16_t E[WIDE_E];
	__align__(8) uint8_t  H[WIDE_H];
#ifdef WIDE_PROF
	long long pt[16];               /* cycles per phase (thread 0), rounds, passes, sweeps, bytes */
	long long runs[16];             /* lanes that decoded in pass k */
	long long pcyc[16];             /* cycles of pass k */
	long long t_pass;
	long long t_last;
#endif
};

#ifdef WIDE_PROF
#define WPROF(k) do { if (threadIdx.x == 0) { const long long t_ = clock64(); w->pt[k] += t_ - w->t_last; w->t_last = t_; } } while (0)
#define WCOUNT(k, n) do { if (threadIdx.x == 0) w->pt[k] += (n); } while (0)
#else
#define WPROF(k) do { } while (0)
#define WCOUNT(k, n) do { } while (0)
#endif

/* all threads of the CTA; the command block is set and a barrier lies behind us */
static __device__ __noinline__ void
wide_round(const uint32_t* lit_, const uint32_t* dtab_, WideMem* w_)
{
	const uint32_t* const lit = jdb_pin_shared(lit_);
	const uint32_t* const dtab = jdb_pin_shared(dtab_);
	WideMem* const w = jdb_pin_shared(w_);
	const unsigned lane = jdb_lane(), wp = jdb_warp(), gl = threadIdx.x;
	const uint32_t o = w->o, S = w->S, endbit = w->endbit;
	const uint32_t safe_end = endbit - 64u;
	uint32_t* const inbuf = jdb_pin_shared(w->inbuf[wp]);        /* (in registers, not rebuilt from the warp number in the decode loop) */
	uint32_t* const slots = jdb_pin_shared(w->slots[wp] + lane);
	uint8_t* const dst = w->dst;
	const uint64_t out = w->out;
	const uint32_t rbase = (uint32_t) (uintptr_t) dst;
	const uint32_t a = (uint32_t) ((uintptr_t) (dst + out) & 3u);       /* E-index of the round's first byte */
	WPROF(0);           /* master between rounds */

	/* ---- new block tables: flatten them (every thread a few entries; the two-level look-up with the
	 * unknown upper bits zero is the answer for all of them when the code it finds is short enough) ---- */
	if (w->tables_new) {
		for (uint32_t i = gl; i < (1u << WIDE_LROOT); i += WIDE_LANES) {
			const uint32_t e = lookup32(lit, i, LIT_ROOT);
			const uint32_t nb = e & 15u, type = (e >> 8) & 3u;
			const bool plain = nb != 0 && nb <= WIDE_LROOT && (type == T_LIT || (type == T_BASE && (e >> 16) != 0));
			w->wlit[i] = plain ? e : W_SPECIAL;
		}
		for (uint32_t i = gl; i < (1u << WIDE_DROOT); i += WIDE_LANES) {
			const uint32_t d = lookup32(dtab, i, DIST_ROOT);
			const uint32_t nb = d & 15u;
			const bool plain = nb != 0 && nb <= WIDE_DROOT && ((d >> 8) & 3u) == T_BASE && (d >> 16) != 0;
			w->wdist[i] = plain ? d : W_SPECIAL;
		}
		__syncthreads();
		if (gl == 0) w->tables_new = 0;
	}
	const uint32_t* const wlit = w->wlit;
	const uint32_t* const wdist = w->wdist;

	/* ---- the history mirror: whatever was produced since the last round (step-by-step
	 * batches, stored blocks; everything, on the first round of a call) comes from L2 ---- */
	{
		int64_t lo = (int64_t) out - (int64_t) WIDE_H;
		if (w->h_valid && (int64_t) w->h_hi > lo) lo = (int64_t) w->h_hi;
		for (int64_t p = lo + (int64_t) gl; p < (int64_t) out; p += WIDE_LANES) {
			uint8_t v = 0;
			if (p >= 0) v = __ldcg(dst + p);
			else if (w->history && (uint64_t) (-p) <= w->hist_avail)
				v = __ldcg(w->history + ((w->total_before + (uint64_t) p) & (JDB_INFLATE_HISTORY - 1)));
			w->H[(rbase + (uint32_t) p) & (WIDE_H - 1u)] = v;
		}
	}
	/* ---- this warp's part of the input window ---- */
	{
		uint32_t b0 = o + 32u * wp * S, b1 = o + 32u * (wp + 1u) * S;
		b0 = b0 >= o + WIDE_LEAD ? b0 - WIDE_LEAD : o;
		if (b0 > safe_end) b0 = safe_end;
		if (b1 > safe_end) b1 = safe_end;
		const uint32_t last = ((b1 + 64u) >> 5) + 1u;
		for (uint32_t k = (b0 >> 5) + lane; k <= last; k += 32u) {
			uint32_t v = 0;
			if (k < w->nwords) {
				const int64_t gi = (int64_t) k + w->delta;
				if (gi >= 0) v = __ldg(w->wbase + gi);
				if (k == 0 && w->pre) v = (v & ~(0xffu << (8u * w->c))) | w->prefix_word;
			}
			inbuf[k & (INW - 1u)] = v;
		}
	}
	__syncwarp();
	WPROF(1);           /* mirror + staging */

	/* ---- speculative decode until the links between neighbours hold (par_round of the
	 * one-warp decoder, with lane 31 of a warp handing over to lane 0 of the next) ---- */
	LaneRun r;
	uint32_t pe = 0, pf = PF_OK;
	{
		uint32_t lim = o + (gl + 1u) * S;
		if (lim > safe_end) lim = safe_end;
		uint32_t st = o + gl * S;
		if (st > safe_end) st = safe_end;
		else if (gl) st = lane_lead_in(inbuf, lit, dtab, st >= o + WIDE_LEAD ? st - WIDE_LEAD : o, st);
		if (st > safe_end) st = safe_end;
		bool run = true;
		for (int pass = 0; ; pass++) {
#ifdef WIDE_PROF
			if (gl == 0) { const long long t_ = clock64(); if (pass) w->pcyc[pass < 16 ? pass - 1 : 15] += t_ - w->t_pass; w->t_pass = t_; }
#endif
			lane_decode_w(inbuf, lit, dtab, wlit, wdist, slots, r, run, st, lim, safe_end);
			WCOUNT(9, 1);
#ifdef WIDE_PROF
			{
				const unsigned rm = __ballot_sync(JDB_FULL_MASK, run);      /* (one atomic per warp: 512 contended ones cost more than the pass) */
				if (lane == 0 && rm) atomicAdd((unsigned long long*) &w->runs[pass < 15 ? pass : 15], (unsigned long long) __popc(rm));
			}
#endif
			if (lane == 31) { w->wend[wp] = r.end; w->wflag[wp] = r.flag; }
			__syncthreads();
			pe = __shfl_up_sync(JDB_FULL_MASK, r.end, 1);
			pf = __shfl_up_sync(JDB_FULL_MASK, r.flag, 1);
			if (lane == 0 && wp > 0) { pe = w->wend[wp - 1]; pf = w->wflag[wp - 1]; }
			run = gl > 0 && pf == PF_OK && pe != r.start;
			if (pass + 1 >= WIDE_MAXPASS) break;
			if (!__syncthreads_or(run)) break;
			st = pe;
		}
	}
	WPROF(2);           /* decode passes */
	/* ---- the chain: lanes 0..v-1 ---- */
	{
		const bool linked = gl == 0 || (pf == PF_OK && pe == r.start);
		const unsigned broken = __ballot_sync(JDB_FULL_MASK, !linked);
		if (lane == 0) w->wbrk[wp] = broken ? (uint32_t) __ffs(broken) - 1u : 32u;
	}
	__syncthreads();
	uint32_t v = WIDE_LANES;
	for (uint32_t k = 0; k < WIDE_WARPS; k++)
		if (w->wbrk[k] < 32u) { v = 32u * k + w->wbrk[k]; break; }
	/* bytes and tokens in front of every lane */
	uint32_t bincl = gl < v ? r.bytes : 0u;
	uint32_t nincl = gl < v ? r.n : 0u;
	for (int k = 1; k < 32; k <<= 1) {
		const uint32_t tb = __shfl_up_sync(JDB_FULL_MASK, bincl, k);
		const uint32_t tn = __shfl_up_sync(JDB_FULL_MASK, nincl, k);
		if ((int) lane >= k) { bincl += tb; nincl += tn; }
	}
	if (lane == 31) { w->wsb[wp] = bincl; w->wsn[wp] = nincl; }
	__syncthreads();
	for (uint32_t k = 0; k < wp; k++) { bincl += w->wsb[k]; nincl += w->wsn[k]; }
	/* the target room and the reach of the distances end the chain for the step-by-step
	 * decoder, the size of E[] ends it for the next round */
	{
		const bool hard = gl < v && (bincl > w->room || (int64_t) r.need > (int64_t) w->reach + (int64_t) (bincl - r.bytes));
		const bool cap = gl < v && a + bincl > WIDE_CAP;
		const unsigned hm = __ballot_sync(JDB_FULL_MASK, hard), cm = __ballot_sync(JDB_FULL_MASK, cap);
		if (lane == 0) {
			w->whard[wp] = hm ? (uint32_t) __ffs(hm) - 1u : 32u;
			w->wcap[wp] = cm ? (uint32_t) __ffs(cm) - 1u : 32u;
		}
	}
	__syncthreads();
	uint32_t stepwise = 0, capped = 0;
	{
		uint32_t vh = WIDE_LANES, vc = WIDE_LANES;
		for (uint32_t k = 0; k < WIDE_WARPS; k++)
			if (w->whard[k] < 32u) { vh = 32u * k + w->whard[k]; break; }
		for (uint32_t k = 0; k < WIDE_WARPS; k++)
			if (w->wcap[k] < 32u) { vc = 32u * k + w->wcap[k]; break; }
		if (vh < v && vh <= vc) { v = vh; stepwise = 1; }
		else if (vc < v) { v = vc; capped = 1; }
	}
	{
		const unsigned full = __ballot_sync(JDB_FULL_MASK, gl < v && r.flag == PF_FULL);
		const uint32_t maxn = __reduce_max_sync(JDB_FULL_MASK, gl < v ? r.n : 0u);
		if (lane == 0) { w->wfull[wp] = full != 0; w->wmaxn[wp] = maxn; }
	}
	if (v && gl == v - 1u) {
		w->newo = r.end;
		w->nbytes = bincl;
		w->lastflag = r.flag;
	}
	if (v == 0 && gl == 0) {
		w->newo = o;
		w->nbytes = 0;
		w->lastflag = PF_OK;
	}
	__syncthreads();
	const uint32_t nbytes = w->nbytes;
	if (gl == 0) {
		/* next round: shorter subsequences when token slots or E[] ran out, longer ones
		 * again while both stay half empty */
		uint32_t full = 0, maxn = 0, ns = S;
		for (uint32_t k = 0; k < WIDE_WARPS; k++) { full |= w->wfull[k]; if (w->wmaxn[k] > maxn) maxn = w->wmaxn[k]; }
		if (full || capped) ns = S >= WIDE_SMIN + 32u ? S - 32u : WIDE_SMIN;
		else if (maxn <= 3u * PQ_K / 4u && S < PAR_SMAX && v >= 32u &&
		         (uint64_t) nbytes * WIDE_LANES * (S + 32u) < (uint64_t) (WIDE_CAP - WIDE_CAP / 8u) * v * S) ns = S + 32u;
		w->newS = ns;
		w->stepwise = stepwise;
		/* entries outside the round in its first and last quad */
		for (uint32_t x = 0; x < a; x++) w->E[x] = E_BYTE;
		for (uint32_t x = a + nbytes; x < ((a + nbytes + 3u) & ~3u); x++) w->E[x] = E_BYTE;
	}

	WPROF(3);           /* chain, sums, cuts */
	/* ---- expand: one entry per trip of ONE loop, so that the lanes of a warp, whose tokens
	 * differ in kind and length, stay together (a loop per token would run as long as the
	 * longest token of every step); matches longer than WIDE_LONG bytes are left out here and
	 * written by all lanes of the warp afterwards ---- */
	{
		const uint32_t hb = rbase + (uint32_t) out - a;        /* H index of E-index 0 */
		const uint32_t mybytes = gl < v ? r.bytes : 0u;
		const uint32_t myn = gl < v ? r.n : 0u;
		const uint32_t x0 = a + bincl - mybytes;
		{
			uint32_t x = x0, i = 0, left = 0;
			int32_t back = 0;             /* distance of the match in hand */
			while (__any_sync(JDB_FULL_MASK, i < myn || left)) {
				if (!(i < myn || left)) continue;
				if (left == 0) {
					const uint32_t tok = slots[i * 32u];
					i++;
					const uint32_t len = tok >> 16;
					if (len == 0) {
						w->E[x++] = (uint16_t) (E_BYTE | tok);
						continue;
					}
					if (len > WIDE_LONG) { x += len; continue; }
					left = len;
					back = (int32_t) (tok & 0xffffu);
				}
				/* up to four entries of the match in hand per trip */
				const uint32_t k = left < 4u ? left : 4u;
				const int32_t sp = (int32_t) x - back;
#pragma unroll
				for (uint32_t j = 0; j < 4u; j++)
					if (j < k)
						w->E[x + j] = sp + (int32_t) j >= (int32_t) a
						            ? (uint16_t) (sp + (int32_t) j)
						            : (uint16_t) (E_BYTE | w->H[(hb + (uint32_t) sp + j) & (WIDE_H - 1u)]);
				x += k;
				left -= k;
			}
		}
		WPROF(12);          /* warp 0: flat expand loop */
		const uint32_t maxn = __reduce_max_sync(JDB_FULL_MASK, myn);
		uint32_t x = x0;
		for (uint32_t i = 0; i < maxn; i++) {
			const uint32_t tok = i < myn ? slots[i * 32u] : 0u;
			const uint32_t len = tok >> 16;
			unsigned lm = __ballot_sync(JDB_FULL_MASK, len > WIDE_LONG);
			while (lm) {
				const int from = __ffs(lm) - 1;
				lm &= lm - 1;
				const uint32_t x2 = __shfl_sync(JDB_FULL_MASK, x, from);
				const uint32_t t2 = __shfl_sync(JDB_FULL_MASK, tok, from);
				const uint32_t l2 = t2 >> 16;
				const int32_t d2 = (int32_t) (t2 & 0xffffu);
				for (uint32_t j = lane; j < l2; j += 32u) {
					const int32_t sp = (int32_t) (x2 + j) - d2;
					w->E[x2 + j] = sp >= (int32_t) a ? (uint16_t) sp : (uint16_t) (E_BYTE | w->H[(hb + (uint32_t) sp) & (WIDE_H - 1u)]);
				}
			}
			x += len ? len : (i < myn ? 1u : 0u);
		}
	}
	WPROF(13);          /* warp 0: long matches */
	__syncthreads();
	WPROF(4);           /* expand: waiting for the other warps */

	/* ---- resolve ---- */
	const uint32_t nquad = (a + nbytes + 3u) >> 2;
	uint2* const E4 = (uint2*) w->E;
	while (nbytes) {
		bool more = false;
		WCOUNT(10, 1);
		for (uint32_t q0 = 0; q0 < nquad; q0 += WIDE_LANES, __syncwarp()) {
			const uint32_t q = q0 + gl;
			if (q >= nquad) continue;
			uint2 e4 = E4[q];
			if ((e4.x & e4.y & 0x80008000u) == 0x80008000u) continue;
			/* (the four sources are loaded before any of them is looked at) */
			const uint32_t e0 = e4.x & 0xffffu, e1 = e4.x >> 16, e2 = e4.y & 0xffffu, e3 = e4.y >> 16;
			const uint32_t f0 = (e0 & E_BYTE) ? e0 : w->E[e0];
			const uint32_t f1 = (e1 & E_BYTE) ? e1 : w->E[e1];
			const uint32_t f2 = (e2 & E_BYTE) ? e2 : w->E[e2];
			const uint32_t f3 = (e3 & E_BYTE) ? e3 : w->E[e3];
			e4.x = f0 | (f1 << 16);
			e4.y = f2 | (f3 << 16);
			if ((e4.x & e4.y & 0x80008000u) != 0x80008000u) more = true;
			E4[q] = e4;
		}
		if (!__syncthreads_or(more)) break;
	}
	WPROF(5);           /* resolve */
	/* ---- write: target and mirror ---- */
	{
		uint8_t* const dal = dst + out - a;                       /* 4-byte aligned */
		const uint32_t hal = rbase + (uint32_t) out - a;
		for (uint32_t q0 = 0; q0 < nquad; q0 += WIDE_LANES, __syncwarp()) {
			const uint32_t q = q0 + gl;
			if (q >= nquad) continue;
			const uint2 e4 = E4[q];
			const uint32_t word = __byte_perm(e4.x, e4.y, 0x6420);
			const uint32_t x0 = 4u * q;
			if (x0 >= a && x0 + 4u <= a + nbytes) {
				*(uint32_t*) (dal + x0) = word;
				*(uint32_t*) (w->H + ((hal + x0) & (WIDE_H - 1u))) = word;
			} else {
				for (uint32_t k = 0; k < 4; k++) {
					const uint32_t x = x0 + k;
					if (x >= a && x < a + nbytes) {
						const uint8_t b = (uint8_t) (word >> (8u * k));
						dal[x] = b;
						w->H[(hal + x) & (WIDE_H - 1u)] = b;
					}
				}
			}
		}
	}
	if (gl == 0) {
		w->h_hi = out + nbytes;
		w->h_valid = 1;
	}
	__syncthreads();
	WPROF(6);           /* write */
	WCOUNT(8, 1);
	WCOUNT(11, nbytes);
	WCOUNT(7, v);
}

/* all threads of the CTA: the bytes of a stored block, source to target (any alignments): target-aligned
 * 32-bit words, each from the two aligned source words that hold its bytes */
static __device__ __noinline__ void
wide_copy(WideMem* w_)
{
	WideMem* const w = jdb_pin_shared(w_);
	const unsigned gl = threadIdx.x;
	uint8_t* const d = w->cp_dst;
	const uint8_t* const p = w->cp_src;
	const uint32_t n = w->cp_n;
	uint32_t head = (4u - (uint32_t) ((uintptr_t) d & 3u)) & 3u;
	if (head > n) head = n;
	const uint32_t nw = (n - head) >> 2;
	const uint32_t tail0 = head + 4u * nw;
	if (gl < head) d[gl] = __ldg(p + gl);
	if (tail0 + gl < n) d[tail0 + gl] = __ldg(p + tail0 + gl);
	const uint8_t* const ps = p + head;
	const uint32_t* const pw = (const uint32_t*) ((uintptr_t) ps & ~(uintptr_t) 3);
	const uint32_t sh = 8u * (uint32_t) ((uintptr_t) ps & 3u);
	uint32_t* const dw = (uint32_t*) (d + head);
#pragma unroll 4
	for (uint32_t i = gl; i < nw; i += WIDE_LANES) {
		const uint32_t lo = __ldg(pw + i);
		const uint32_t hi = sh ? __ldg(pw + i + 1) : 0u;      /* never past the word of the last byte */
		dw[i] = __funnelshift_r(lo, hi, sh);
	}
	__syncthreads();
}

/* all threads of the CTA: Adler-32 of the n bytes at cp_src (the decoded stream, read back through L2);
 * warp_adler32 on 512 threads -- a lone warp took as long over the checksum of a long stream as the CTA
 * over decoding it */
static __device__ __noinline__ void
wide_adler(WideMem* w_)
{
	WideMem* const w = jdb_pin_shared(w_);
	const unsigned gl = threadIdx.x, lane = jdb_lane(), wp = jdb_warp();
	const uint8_t* const p = w->cp_src;
	const uint64_t n = w->ad_n;
	unsigned long long a = 0, b = 0;
	uint32_t k = 0;
#pragma unroll 8
	for (uint64_t i = gl; i < n; i += WIDE_LANES) {
		const uint32_t v = __ldcg(p + i);
		a += v;
		b += (unsigned long long) (n - i) * v;
		if (++k == 65536u) { b %= 65521u; k = 0; }      /* (n - i) * v < 2^40 for the sizes a batch item can have */
	}
	a %= 65521u;
	b %= 65521u;
	for (int o = 16; o; o >>= 1) {
		a += __shfl_xor_sync(JDB_FULL_MASK, a, o);
		b += __shfl_xor_sync(JDB_FULL_MASK, b, o);
	}
	if (lane == 0) { w->ad_a[wp] = a; w->ad_b[wp] = b; }
	__syncthreads();
	if (gl == 0) {
		unsigned long long sa = 0, sb = 0;
		for (uint32_t q = 0; q < WIDE_WARPS; q++) { sa += w->ad_a[q]; sb += w->ad_b[q]; }
		sa = (1 + sa) % 65521u;                        /* start value 1: a = 1 + sum, b = n * 1 + weighted sum */
		sb = (n % 65521u + sb % 65521u) % 65521u;
		w->ad_result = (uint32_t) ((sb << 16) | sa);
	}
	__syncthreads();
}

/*
 * Decode one stream (or one call's worth of a streaming decode).
 * Called by all 32 lanes of a warp with identical arguments.
 */
template <bool WIDE>
static __device__ void
inflate_stream(WarpMem* m_, Stream& s, WideMem* w_)
{
	/* (the master of a wide CTA is a lone warp on the critical path of its CTA: see jdb_pin_shared) */
	WarpMem* const m = WIDE ? jdb_pin_shared(m_) : m_;
	WideMem* const w = WIDE ? jdb_pin_shared(w_) : w_;
	const unsigned lane = jdb_lane();
	Bits b;
	uint32_t phase = JDB_INF_HEADER;    /* where we are in the block structure */
	uint32_t lastblock = 0;
	uint32_t stored_left = 0;
	uint32_t pend_len = 0, pend_dist = 0;

	b.bb = 0; b.bc = 0; b.p = s.src; b.end = s.src + s.src_len;
	s.out = 0;
	s.status = ST_OK;
	s.error = 0;
	s.hist_avail = 0;
	s.total_before = 0;
	s.ring = m->ring;
	s.ring_lo = 0;
	s.rbase = (uint32_t) (uintptr_t) s.dst;
	s.flushed = 0;

	if (s.st) {
		jdb_inflate_state* st = s.st;
		b.bb = st->bitbuf;
		b.bc = st->bitcnt;
		phase = st->phase;
		lastblock = st->lastblock;
		stored_left = st->stored_left;
		pend_len = st->pend_len;
		pend_dist = st->pend_dist;
		s.hist_avail = st->hist_avail;
		s.total_before = st->total_out;
		if (phase == JDB_INF_SYMBOLS) {
			/* tables of the block in progress */
			for (int i = lane; i < LIT_TABLE; i += 32) m->lit[i] = st->lit[i];
			for (int i = lane; i < DIST_TABLE; i += 32) m->dist[i] = st->dist[i];
			__syncwarp();
		}
	}

	/* (a match that did not fit into the previous target window is finished by the first
	 * round of the symbol loop: pend_len != 0 implies phase == JDB_INF_SYMBOLS) */

	for (;;) {
		/* ---------------- block header ---------------- */
		if (phase == JDB_INF_HEADER) {
			/* The whole header (3 bits, stored LEN/NLEN or the dynamic code
			 * lengths) is parsed speculatively: when the input runs out inside
			 * it lane 0 rewinds to `hdr` and the header is replayed by the next
			 * call, so no partial-header state has to be kept. */
			uint32_t type = 0, lb = 0;
			Bits hdr = b;
			if (lastblock) { s.status = ST_OK; break; }
			const uint32_t r = parse_block_header(&m->bm, m->lit, m->dist, b, type, lb);
			/* the header is parsed by lane 0; everywhere else the bit reader is kept
			 * identical on all lanes (uniform control flow in the symbol loop) */
			b.bb = __shfl_sync(JDB_FULL_MASK, (unsigned long long) b.bb, 0);
			b.bc = __shfl_sync(JDB_FULL_MASK, b.bc, 0);
			b.p = (const uint8_t*) __shfl_sync(JDB_FULL_MASK, (unsigned long long) b.p, 0);
			if (r == 1) {
				b = hdr;
				s.status = ST_SRCEXH;
				break;
			}
			if (r > 1) { s.status = ST_ERROR; s.error = r - 1; break; }
			lastblock = lb;
			if (type == 0) {
				stored_left = m->bm.scratch[0];
				phase = JDB_INF_STORED;
			} else {
				phase = JDB_INF_SYMBOLS;
			}
		}

		/* ---------------- stored block ---------------- */
		if (phase == JDB_INF_STORED) {
			/* whole bytes in the bit buffer go back to the input */
			b.p -= b.bc >> 3;
			b.bb = 0;
			b.bc = 0;
			const uint64_t pos = (uint64_t) (b.p - s.src);
			uint64_t n = stored_left;
			uint64_t srcleft = s.src_len - pos, dstleft = s.dst_cap - s.out;
			if (n > srcleft) n = srcleft;
			if (n > dstleft) n = dstleft;
			if (WIDE && n >= WIDE_COPY_MIN) {
				/* a block worth the barriers: copied by all warps; like the bytes of a wide round,
				 * these reach later matches through L2, not through the master's ring */
				flush_ring(s, s.out, true);
				if (lane == 0) {
					w->cmd = 2;
					w->cp_src = s.src + pos;
					w->cp_dst = s.dst + s.out;
					w->cp_n = (uint32_t) n;
				}
				__syncthreads();
				wide_copy(w);
				s.flushed = s.out + n;
				s.ring_lo = (int64_t) (s.out + n);
			} else
			if (!s.count_only) {
				/* straight to the target, and into the ring for the matches of later blocks */
				flush_ring(s, s.out, true);
				{
					/* target-aligned 32-bit words, each from the two aligned source words that
					 * hold its bytes; the head and tail bytes on their own.  Only the newest
					 * RING bytes are worth a place in the ring. */
					uint8_t* const d = s.dst + s.out;
					const uint8_t* const p = s.src + pos;
					uint64_t head = (uint64_t) ((4u - ((uintptr_t) d & 3u)) & 3u);
					if (head > n) head = n;
					const uint64_t nw = (n - head) >> 2;
					const uint64_t tail0 = head + 4u * nw;
					const uint64_t keep0 = n > RING ? n - RING : 0;      /* bytes from here on go to the ring */
					if (lane < head) {
						const uint8_t v = p[lane];
						d[lane] = v;
						if (lane >= keep0) s.ring[ring_index(s, (uint32_t) (s.out + lane))] = v;
					}
					const uint8_t* const ps = p + head;
					const uint32_t* const pw = (const uint32_t*) ((uintptr_t) ps & ~(uintptr_t) 3);
					const uint32_t sh = 8u * (uint32_t) ((uintptr_t) ps & 3u);
					uint32_t* const dw = (uint32_t*) (d + head);
					for (uint64_t i = lane; i < nw; i += 32) {
						const uint32_t lo = __ldg(pw + i);
						const uint32_t hi = sh ? __ldg(pw + i + 1) : 0u;      /* never past the word of the last byte */
						const uint32_t v = __funnelshift_r(lo, hi, sh);
						dw[i] = v;
						if (head + 4u * i + 4u > keep0)      /* (a word that straddles keep0 brings up to 3 older bytes:
						                                      * their slots are rewritten by this lane, or by the tail below) */
							*(uint32_t*) (s.ring + ((uint32_t) (uintptr_t) (dw + i) & (RING - 1u))) = v;
					}
					__syncwarp();
					if (tail0 + lane < n) {
						const uint8_t v = p[tail0 + lane];
						d[tail0 + lane] = v;
						s.ring[ring_index(s, (uint32_t) (s.out + tail0 + lane))] = v;
					}
				}
				s.flushed = s.out + n;
				const int64_t lo = (int64_t) (s.out + n) - (int64_t) RING;
				if (lo > s.ring_lo) s.ring_lo = lo;
			}
			__syncwarp();
			s.out += n;
			const bool empty_block = stored_left == 0;
			stored_left -= (uint32_t) n;
			b.p += n;
			if (stored_left) {
				s.status = (s.dst_cap - s.out) == 0 ? ST_TGTEXH : ST_SRCEXH;
				break;
			}
			phase = JDB_INF_HEADER;
			if ((s.count_only || s.stop_marker) && empty_block) {
				/* the sync / end marker that closes a chunk of our own encoder
				 * (and of any deflate sync flush) */
				s.status = ST_MARKER;
				break;
			}
			continue;
		}

		/* ---------------- Huffman coded symbols ---------------- */
		if (phase == JDB_INF_SYMBOLS) {
			uint32_t ev = 0;      /* 0 continue, 1 end of block, 2 starved, 3 target full, 4+ error+4 */

			/* ---- enter: a bit window over the unread input --------------------
			 * The symbol loop does not use the byte reader: the input is staged in
			 * shared memory by all lanes (coalesced 128-byte lines into a 1 KiB
			 * ring) and the position is ONE 32-bit bit offset `o`; a symbol is a
			 * two-word peek + funnel shift, a table look-up and an add -- no refill
			 * logic, no 64-bit arithmetic, identical on all lanes (the lit/len and
			 * distance look-ups are shared-memory broadcasts).  Up to 7 bits left in
			 * the byte reader become a virtual byte in front of the window. */
			if (WIDE && lane == 0) w->tables_new = 1;       /* (a new block, or the block in progress of an earlier call) */
			b.p -= b.bc >> 3;
			b.bc &= 7u;
			b.bb &= (1ull << b.bc) - 1ull;
			const uint8_t* const p0 = b.p;
			const uint32_t pre = b.bc ? 1u : 0u;
			const uint32_t al = (uint32_t) ((uintptr_t) p0 & 3u);
			const uint32_t c = (al - pre) & 3u;                       /* ring byte of the virtual byte */
			const int32_t delta = ((int32_t) al - (int32_t) pre - (int32_t) c) / 4;   /* ring word k = global word k + delta */
			const uint32_t* const wbase = (const uint32_t*) (p0 - al);
			uint64_t left = (uint64_t) (b.end - p0);
			const bool clamped = left > (1ull << 28);                 /* 32-bit bit offsets: windows of 256 MiB */
			if (clamped) left = 1ull << 28;
			const uint32_t endbit = 8u * (pre + c + (uint32_t) left);
			const uint32_t nwords = (endbit + 31u) >> 5;
			const uint32_t fill_limit = (endbit + 1023u) & ~1023u;
			const uint32_t prefix_word = pre ? (uint32_t) ((b.bb << (8u - b.bc)) & 0xffull) << (8u * c) : 0u;
			uint32_t o = pre ? 8u * c + 8u - b.bc : 8u * al;
			uint32_t filled = 0;

			uint32_t S = WIDE ? WIDE_S0 : PAR_S0;   /* subsequence length of the lane-parallel rounds */
			bool step_by_step = false;        /* the next batch goes through the step-by-step decoder */
			uint32_t wide_ev = 0;             /* how the last wide round ended (acted on once the master's own input ring follows) */
			for (;;) {
				/* ---- all lanes: keep the input of one lane-parallel round (or >= 3 KiBit) ahead
				 * of `o` in the ring ---- */
				/* (the master of a wide CTA needs its own ring for the step-by-step decoder and for the
				 * way out only: wide rounds stage their input themselves) */
				const bool go_wide = WIDE && !pend_len && !step_by_step && !wide_ev && endbit - o >= 2u * S + 64u &&
				                     s.dst_cap - s.out >= WIDE_MIN_ROOM;      /* the tail of a target window: step by step */
				const uint32_t ahead = WIDE ? 3072u : 32u * S + 64u > 3072u ? 32u * S + 64u : 3072u;
				if (o >= filled) filled = o & ~1023u;
				while (!go_wide && filled < o + ahead && filled < fill_limit) {
					const uint32_t k = (filled >> 5) + lane;
					uint32_t v = 0;
					if (k < nwords) {
						const int64_t gi = (int64_t) k + delta;
						if (gi >= 0) v = __ldg(wbase + gi);
						if (k == 0 && pre) v = (v & ~(0xffu << (8u * c))) | prefix_word;
					}
					m->inbuf[k & (INW - 1u)] = v;
					filled += 1024u;
				}
				__syncwarp();
				if (WIDE && wide_ev) { ev = wide_ev; break; }

				/* ---- lane-parallel round --------------------------------------------------
				 * Lane l decodes the symbols that start in bits [o + l*S, o + (l+1)*S) of the
				 * window.  Only lane 0 knows where its first symbol starts; the others guess
				 * (their subsequence boundary), and a Huffman decoder that starts at a wrong
				 * bit falls into step with the true symbol sequence after a few symbols.  So:
				 * all lanes decode; then every lane whose predecessor ended somewhere else than
				 * where it started decodes again from there, until nothing changes.  Lane 0 is
				 * right by construction, a lane that started where a right lane ended is right:
				 * the chain of agreeing lanes from lane 0 on is the decoded symbol sequence,
				 * whatever the others did.  The round judges nothing: anything but a plain
				 * literal / match / end of block inside the chain ends the chain there and the
				 * step-by-step decoder below takes the next batch from that bit, so every status
				 * and error decision stays in one place. */
				uint32_t gtok = 0, gbytes = 0;       /* tokens in m->cq for the emission below, and their bytes */
				ev = 0;
				if (pend_len) {
					/* the rest of a match the previous target window cut */
					if (lane == 0) m->cq[0] = (pend_len << 16) | pend_dist;
					gtok = 1;
					gbytes = pend_len;
					pend_len = 0;
				} else
#ifndef PAR_OFF
				if (WIDE && go_wide) {
					/* ---- the round on all warps of the CTA (wide_round): everything produced so
					 * far goes to the target first, the round writes behind it ---- */
					flush_ring(s, s.out, true);
					if (lane == 0) {
						const uint64_t room64 = s.dst_cap - s.out;
						const uint64_t reach64 = s.out + s.hist_avail;
						w->cmd = 1;
						w->o = o; w->S = S; w->endbit = endbit; w->nwords = nwords;
						w->pre = pre; w->c = c; w->prefix_word = prefix_word; w->delta = delta;
						w->room = room64 > 0xfffff000ull ? 0xfffff000u : (uint32_t) room64;
						w->reach = reach64 > 0x10000ull ? 0x10000u : (uint32_t) reach64;
						w->wbase = wbase;
						w->dst = s.dst; w->out = s.out;
						w->history = s.st ? s.st->history : (const uint8_t*) 0;
						w->total_before = s.total_before;
						w->hist_avail = s.hist_avail;
					}
					__syncthreads();
					wide_round(m->lit, m->dist, w);
					const uint32_t newo = w->newo, nb = w->nbytes, lastflag = w->lastflag;
					S = w->newS;
					if (nb) {
						s.out += nb;
						s.flushed = s.out;
						s.ring_lo = (int64_t) s.out;       /* the master's ring holds none of these bytes */
					}
					if (lastflag == PF_ANOM || newo == o || w->stepwise) step_by_step = true;
					o = newo;
					wide_ev = lastflag == PF_EOB ? 1u : 0u;
					continue;
				} else
				if (!WIDE && !step_by_step && endbit - o >= 2u * S + 64u) {
					const uint32_t safe_end = endbit - 64u;
					const uint64_t room64 = s.dst_cap - s.out;
					const uint32_t room = room64 > 0xfffff000ull ? 0xfffff000u : (uint32_t) room64;
					const uint64_t reach64 = s.out + s.hist_avail;
					const uint32_t reach = reach64 > 0x10000ull ? 0x10000u : (uint32_t) reach64;
					uint32_t* const slots = m->cq + lane;
					uint32_t lim = o + (lane + 1u) * S;
					if (lim > safe_end) lim = safe_end;
					LaneRun r;
					{
						uint32_t st = o + lane * S;
						if (st > safe_end) st = safe_end;
						bool run = true;
						for (int pass = 0; ; pass++) {
							if (run) lane_decode(m->inbuf, m->lit, m->dist, slots, r, st, lim, safe_end);
							if (pass + 1 >= PAR_MAXPASS) break;
							const uint32_t pe = __shfl_up_sync(JDB_FULL_MASK, r.end, 1);
							const uint32_t pf = __shfl_up_sync(JDB_FULL_MASK, r.flag, 1);
							run = lane > 0 && pf == PF_OK && pe != r.start;
							if (!__any_sync(JDB_FULL_MASK, run)) break;
							st = pe;
						}
					}
					/* the chain: lanes 0..v */
					uint32_t v;
					{
						const uint32_t pe = __shfl_up_sync(JDB_FULL_MASK, r.end, 1);
						const uint32_t pf = __shfl_up_sync(JDB_FULL_MASK, r.flag, 1);
						const bool linked = lane == 0 || (pf == PF_OK && pe == r.start);
						const unsigned broken = __ballot_sync(JDB_FULL_MASK, !linked);
						v = broken ? (uint32_t) __ffs(broken) - 1u : 32u;      /* lanes 0..v-1 are in the chain */
					}
					/* bytes in front of every lane; the target room and the reach of the distances
					 * may cut the chain short (the step-by-step decoder finds out exactly where) */
					uint32_t bincl = lane < v ? r.bytes : 0u;
					uint32_t nincl = lane < v ? r.n : 0u;
					for (int k = 1; k < 32; k <<= 1) {
						const uint32_t tb = __shfl_up_sync(JDB_FULL_MASK, bincl, k);
						const uint32_t tn = __shfl_up_sync(JDB_FULL_MASK, nincl, k);
						if ((int) lane >= k) { bincl += tb; nincl += tn; }
					}
					{
						const bool bad = lane < v && (bincl > room || (int64_t) r.need > (int64_t) reach + (int64_t) (bincl - r.bytes));
						const unsigned cut = __ballot_sync(JDB_FULL_MASK, bad);
						if (cut) {
							v = (uint32_t) __ffs(cut) - 1u;
							step_by_step = true;
						}
					}
					uint32_t ntok = 0, nbytes = 0, lastflag = PF_OK, newo = o;
					if (v) {
						ntok = __shfl_sync(JDB_FULL_MASK, nincl, (int) v - 1);
						nbytes = __shfl_sync(JDB_FULL_MASK, bincl, (int) v - 1);
						lastflag = __shfl_sync(JDB_FULL_MASK, r.flag, (int) v - 1);
						newo = __shfl_sync(JDB_FULL_MASK, r.end, (int) v - 1);
					}
					/* next round: shorter subsequences when the token slots ran out, longer ones
					 * again when they stay half empty */
					{
						const unsigned full = __ballot_sync(JDB_FULL_MASK, lane < v && r.flag == PF_FULL);
						const uint32_t maxn = __reduce_max_sync(JDB_FULL_MASK, lane < v ? r.n : 0u);
						if (full) S = S >= PAR_SMIN + 32u ? S - 32u : PAR_SMIN;
						else if (maxn <= PQ_K / 2u && S < PAR_SMAX) S += 32u;
					}
					if (ntok && !s.count_only) {
						/* compact the tokens of the chain: through registers, in place */
						uint32_t t[PQ_K];
						const uint32_t mine = lane < v ? r.n : 0u;
						const uint32_t at = nincl - mine;
#pragma unroll
						for (uint32_t i = 0; i < PQ_K; i++)
							if (i < mine) t[i] = slots[i * 32u];
						__syncwarp();
#pragma unroll
						for (uint32_t i = 0; i < PQ_K; i++)
							if (i < mine) m->cq[at + i] = t[i];
					}
					ev = lastflag == PF_EOB ? 1u : 0u;
					if (lastflag == PF_ANOM || newo == o) step_by_step = true;
					o = newo;
					gtok = ntok;
					gbytes = nbytes;
				} else
#endif
				{
				step_by_step = false;

				/* ---- decode up to QUEUE symbols (uniform; lane 0 writes the queue) ---- */
				uint32_t nq = 0;
				uint32_t qbytes = 0;
				{
					/* everything in 32 bits: a queue never holds more than MAXBATCH + 258 bytes */
					const uint64_t room64 = s.dst_cap - s.out;
					const uint32_t room = room64 > 0xfffff000ull ? 0xfffff000u : (uint32_t) room64;
					const uint64_t reach64 = s.out + s.hist_avail;
					const uint32_t reach = reach64 > 0x10000ull ? 0x10000u : (uint32_t) reach64;   /* distances are <= 32768 */
					while (nq < QUEUE && qbytes < MAXBATCH) {
						const uint32_t avail = endbit - o;
						const uint32_t bits = peek32(m->inbuf, o);
						const uint32_t e = lookup32(m->lit, bits, LIT_ROOT);
						uint32_t nb = e & 15u;
						/* ---- the two common cases with one combined test each; anything
						 * unusual falls through to the step-by-step code below, which owns
						 * the exact order of the status / error decisions ---- */
						if (((e >> 8) & 3u) == T_LIT) {
							if (nb - 1u < avail && qbytes < room) {
								o += nb;
								if (lane == 0) m->cq[nq] = e >> 16;
								nq++;
								qbytes++;
								continue;
							}
						} else if (((e >> 8) & 3u) == T_BASE && (e >> 16) != 0) {
							const uint32_t lxb = (e >> 4) & 15u;
							const uint32_t flen = (e >> 16) + ((bits >> nb) & ((1u << lxb) - 1u));
							const uint32_t fo2 = o + nb + lxb;
							const uint32_t fbits2 = peek32(m->inbuf, fo2);
							const uint32_t fd = lookup32(m->dist, fbits2, DIST_ROOT);
							const uint32_t dnb = fd & 15u, dxb = (fd >> 4) & 15u;
							const uint32_t fdist = (fd >> 16) + ((fbits2 >> dnb) & ((1u << dxb) - 1u));
							const uint32_t fo3 = fo2 + dnb + dxb;
							if (dnb != 0 && (fd >> 16) != 0 && fo3 <= endbit && fdist <= reach + qbytes && qbytes + flen <= room) {
								o = fo3;
								if (lane == 0) m->cq[nq] = (flen << 16) | fdist;
								nq++;
								qbytes += flen;
								continue;
							}
						}
						if (nb == 0) {
							/* no code for these bits; with a short tail it may also be starvation */
							ev = avail < 15u ? 2u : 4u + E_BADCODE;
							break;
						}
						if (nb > avail) { ev = 2; break; }
						const uint32_t type = (e >> 8) & 3u;
						if (type == T_LIT) {
							if (qbytes >= room) { ev = 3; break; }
							o += nb;
							if (lane == 0) m->cq[nq] = e >> 16;    /* len field 0: literal */
							nq++;
							qbytes++;
							continue;
						}
						if (type == T_EOB) { o += nb; ev = 1; break; }
						if ((e >> 16) == 0) { o += nb; ev = 4 + E_BADCODE; break; }     /* reserved symbol 286/287 */
						/* length + distance */
						uint32_t xb = (e >> 4) & 15u;
						if (nb + xb > avail) { ev = 2; break; }
						const uint32_t len = (e >> 16) + ((bits >> nb) & ((1u << xb) - 1u));
						uint32_t o2 = o + nb + xb;
						const uint32_t avail2 = endbit - o2;
						const uint32_t bits2 = peek32(m->inbuf, o2);
						const uint32_t d = lookup32(m->dist, bits2, DIST_ROOT);
						nb = d & 15u;
						if (nb == 0) {
							if (avail2 < 15u) ev = 2;
							else { o = o2; ev = 4 + E_BADCODE; }
							break;
						}
						if (nb > avail2) { ev = 2; break; }
						if ((d >> 16) == 0) { o = o2; ev = 4 + E_BADCODE; break; }      /* reserved symbol 30/31 */
						xb = (d >> 4) & 15u;
						if (nb + xb > avail2) { ev = 2; break; }
						const uint32_t dist = (d >> 16) + ((bits2 >> nb) & ((1u << xb) - 1u));
						o2 += nb + xb;
						if (dist > reach + qbytes) { o = o2; ev = 4 + E_FAROFFSET; break; }
						if (qbytes >= room) { ev = 3; break; }
						o = o2;
						if (lane == 0) m->cq[nq] = (len << 16) | dist;       /* len <= 258, dist <= 32768 */
						nq++;
						qbytes += len;
						if (qbytes > room) { ev = 3; break; }        /* partially fits: split below */
					}
				}
				gtok = nq;
				gbytes = qbytes;
				}
				__syncwarp();

				/* ---- all lanes: turn the tokens into bytes ---- */
				__syncwarp();
				if (gtok) {
					if (s.count_only) {
						const uint64_t room64 = s.dst_cap - s.out;
						s.out += gbytes < room64 ? gbytes : room64;
					} else {
						for (uint32_t done = 0; done < gtok; ) {
							const uint32_t take = gtok - done < 32u ? gtok - done : 32u;
							done += emit_queue(s, m->pend, m->cq + done, take, pend_len, pend_dist);
						}
					}
					if (pend_len) ev = 3;
				}
				if (ev) break;
			}

			/* ---- leave: back to the byte reader (a partly used byte counts as
			 * consumed, its unused bits stay in the bit buffer) ---- */
			{
				const uint32_t r = o >> 3, k = o & 7u;
				const uint32_t byte = (m->inbuf[(r >> 2) & (INW - 1u)] >> (8u * (r & 3u))) & 0xffu;
				b.p = p0 + ((int64_t) r - (int64_t) pre - (int64_t) c) + (k ? 1 : 0);
				b.bc = k ? 8u - k : 0u;
				b.bb = k ? (uint64_t) (byte >> k) : 0ull;
			}
			if (ev == 2 && clamped) continue;         /* only the 256 MiB window ended, not the input */
			if (ev == 1) { phase = JDB_INF_HEADER; continue; }
			if (ev == 2) { s.status = ST_SRCEXH; break; }
			if (ev == 3) { s.status = ST_TGTEXH; break; }
			s.status = ST_ERROR;
			s.error = ev - 4;
			break;
		}
	}

	if (!s.count_only) flush_ring(s, s.out, true);
	/* a starved final input is an error (src/inflator.c:810-816, 838-842) */
	if (s.status == ST_SRCEXH && s.final) {
		s.status = ST_ERROR;
		s.error = E_INPUTEND;
	}
	{
		/* consumed bytes; at the end of the stream whole unread bytes go back */
		if (s.status == ST_OK || s.status == ST_MARKER) {
			b.p -= b.bc >> 3;
			b.bc &= 7u;
		}
		s.consumed = (uint64_t) (b.p - s.src);
		if (s.status == ST_MARKER) s.error = lastblock;       /* 1: the marker carried BFINAL */

		if (s.st && s.status != ST_ERROR) {
			jdb_inflate_state* st = s.st;
			/* keep the last 32 KiB of output for later calls (updatewindow,
			 * src/inflator.c:616-675) */
			uint64_t keep = s.out < JDB_INFLATE_HISTORY ? s.out : JDB_INFLATE_HISTORY;
			uint64_t abs0 = s.total_before + s.out - keep;
			for (uint64_t j = lane; j < keep; j += 32)
				st->history[(abs0 + j) & (JDB_INFLATE_HISTORY - 1)] = __ldcg(s.dst + (s.out - keep) + j);
			if (phase == JDB_INF_SYMBOLS) {
				for (int i = lane; i < LIT_TABLE; i += 32) st->lit[i] = m->lit[i];
				for (int i = lane; i < DIST_TABLE; i += 32) st->dist[i] = m->dist[i];
			}
			if (lane == 0) {
				st->bitbuf = b.bb;
				st->bitcnt = b.bc;
				st->phase = phase;
				st->lastblock = lastblock;
				st->stored_left = stored_left;
				st->pend_len = pend_len;
				st->pend_dist = pend_dist;
				st->total_out = s.total_before + s.out;
				uint64_t h = s.hist_avail + s.out;
				st->hist_avail = h > JDB_INFLATE_HISTORY ? JDB_INFLATE_HISTORY : h;
			}
		}
	}
	__syncwarp();
}

/* ---- container framing around one stream (zlib / gzip records) --------------- */

static __device__ uint32_t
warp_adler32(const uint8_t* p, uint64_t n)
{
	/* whole-record Adler-32 by one warp: lane-strided bytes with position weights */
	const unsigned lane = jdb_lane();
	unsigned long long a = 0, bsum = 0;
	for (uint64_t i = lane; i < n; i += 32) {
		uint32_t v = __ldcg(p + i);
		a += v;
		bsum += (unsigned long long) (n - i) * v;
	}
	for (int o = 16; o; o >>= 1) {
		a += __shfl_xor_sync(JDB_FULL_MASK, a, o);
		bsum += __shfl_xor_sync(JDB_FULL_MASK, bsum, o);
	}
	/* start value 1: a = 1 + sum, b = n*1 + weighted sum */
	a = (1 + a) % 65521u;
	bsum = (n % 65521u + bsum % 65521u) % 65521u;
	return (uint32_t) ((bsum << 16) | a);
}

__global__ void __launch_bounds__(INF_THREADS)
inflate_batch_kernel(const uint8_t* __restrict__ src_base, uint8_t* __restrict__ dst_base,
                     const jdb_inflate_item* __restrict__ items, jdb_inflate_result* __restrict__ results,
                     jdb_inflate_state* states, uint32_t count, uint32_t format, uint32_t final,
                     uint32_t* __restrict__ counter, uint32_t redo_only, uint32_t count_only)
{
	JDB_DYN_SMEM(smem_raw);
	WarpMem* m = (WarpMem*) smem_raw + jdb_warp();
	const unsigned lane = jdb_lane();

	for (;;) {
		uint32_t idx = 0;
		if (lane == 0) idx = atomicAdd(counter, 1u);
		idx = __shfl_sync(JDB_FULL_MASK, idx, 0);
		if (idx >= count) break;
		/* (a second pass over what an earlier kernel handed over; unused today) */
		if (redo_only && results[idx].status != ST_REDO) continue;

		const jdb_inflate_item it = items[idx];
		Stream s;
		s.src = src_base + it.src_off;
		s.src_len = it.src_len;
		s.dst = dst_base + it.dst_off;
		s.dst_cap = it.dst_cap;
		s.final = (int) final;
		s.count_only = (int) (count_only & 1u);
		s.stop_marker = (int) ((count_only >> 1) & 1u);
		s.st = states ? states + idx : NULL;

		uint32_t zerr = 0;
		uint32_t head = 0;
		if (format == JDB_FMT_ZLIB) {
			/* RFC 1950 header as the reference parses it: CM 8, CINFO <= 7,
			 * FCHECK ignored (src/zstrm.c:510-565); preset dictionaries are not
			 * available in batch mode */
			if (s.src_len < 2) zerr = JDB_ZERR_BADDATA;
			else {
				uint32_t cmf = s.src[0], flg = s.src[1];
				if ((cmf & 15u) != 8 || (cmf >> 4) > 7) zerr = JDB_ZERR_BADDATA;
				else if (flg & 0x20u) zerr = JDB_ZERR_MISSINGDICT;
				head = 2;
			}
		}
		jdb_inflate_result r;
		r.status = ST_ERROR; r.error = 0; r.zerror = zerr; r.checksum = 0;
		r.consumed = 0; r.produced = 0;
		if (!zerr) {
			s.src += head;
			s.src_len -= head;
			inflate_stream<false>(m, s, (WideMem*) 0);
			r.status = s.status;
			r.error = s.error;
			r.consumed = s.consumed + head;
			r.produced = s.out;
			if (format == JDB_FMT_ZLIB && s.status == ST_OK) {
				uint32_t ad = warp_adler32(s.dst, s.out);
				r.checksum = ad;
				if (r.consumed + 4 > it.src_len) {
					r.zerror = JDB_ZERR_BADDATA;
				} else {
					const uint8_t* t = src_base + it.src_off + r.consumed;
					uint32_t want = ((uint32_t) t[0] << 24) | ((uint32_t) t[1] << 16) | ((uint32_t) t[2] << 8) | t[3];
					if (want != ad) r.zerror = JDB_ZERR_CHECKSUM;
					r.consumed += 4;
				}
			}
		}
		if (lane == 0) results[idx] = r;
		__syncwarp();
	}
}

/*
 * One CTA per stream: warp 0 decodes (inflate_stream), the other warps serve its wide rounds.
 * Raw DEFLATE only -- this is the kernel behind inflator_inflate for a stream without chunk
 * markers (container framing is zstrm's business on the host).
 */
__global__ void __launch_bounds__(WIDE_LANES, 1)
inflate_wide_kernel(const uint8_t* __restrict__ src_base, uint8_t* __restrict__ dst_base,
                    const jdb_inflate_item* __restrict__ items, jdb_inflate_result* __restrict__ results,
                    jdb_inflate_state* states, uint32_t format, uint32_t final)
{
	JDB_DYN_SMEM(smem_raw);
	WarpMem* m = (WarpMem*) smem_raw;
	WideMem* w = (WideMem*) (smem_raw + ((sizeof(WarpMem) + 15u) & ~(size_t) 15u));
	const unsigned lane = jdb_lane();
	const uint32_t idx = blockIdx.x;

	if (jdb_warp() != 0) {
		for (;;) {
			__syncthreads();
			const uint32_t cmd = w->cmd;
			if (cmd == 0) break;
			if (cmd == 2) wide_copy(w);
			else if (cmd == 3) wide_adler(w);
			else wide_round(m->lit, m->dist, w);
		}
		return;
	}
	if (lane == 0) { w->h_valid = 0; w->h_hi = 0; }
#ifdef WIDE_PROF
	if (lane == 0) { for (int k = 0; k < 16; k++) w->pt[k] = 0; for (int k = 0; k < 16; k++) { w->runs[k] = 0; w->pcyc[k] = 0; } w->t_last = clock64(); }
#endif
	__syncwarp();
	const jdb_inflate_item it = items[idx];
	Stream s;
	s.src = src_base + it.src_off;
	s.src_len = it.src_len;
	s.dst = dst_base + it.dst_off;
	s.dst_cap = it.dst_cap;
	s.final = (int) final;
	s.count_only = 0;
	s.stop_marker = 0;
	s.st = states ? states + idx : NULL;
	/* container framing as in inflate_batch_kernel (src/zstrm.c:510-565) */
	uint32_t zerr = 0, head = 0;
	if (format == JDB_FMT_ZLIB) {
		if (s.src_len < 2) zerr = JDB_ZERR_BADDATA;
		else {
			const uint32_t cmf = s.src[0], flg = s.src[1];
			if ((cmf & 15u) != 8 || (cmf >> 4) > 7) zerr = JDB_ZERR_BADDATA;
			else if (flg & 0x20u) zerr = JDB_ZERR_MISSINGDICT;
			head = 2;
		}
	}
	jdb_inflate_result r;
	r.status = ST_ERROR; r.error = 0; r.zerror = zerr; r.checksum = 0;
	r.consumed = 0; r.produced = 0;
	if (!zerr) {
		s.src += head;
		s.src_len -= head;
		inflate_stream<true>(m, s, w);
		r.status = s.status;
		r.error = s.error;
		r.consumed = s.consumed + head;
		r.produced = s.out;
		if (format == JDB_FMT_ZLIB && s.status == ST_OK) {
			if (lane == 0) {
				w->cmd = 3;
				w->cp_src = s.dst;
				w->ad_n = s.out;
			}
			__syncthreads();
			wide_adler(w);
			const uint32_t ad = w->ad_result;
			r.checksum = ad;
			if (r.consumed + 4 > it.src_len) {
				r.zerror = JDB_ZERR_BADDATA;
			} else {
				const uint8_t* t = src_base + it.src_off + r.consumed;
				const uint32_t want = ((uint32_t) t[0] << 24) | ((uint32_t) t[1] << 16) | ((uint32_t) t[2] << 8) | t[3];
				if (want != ad) r.zerror = JDB_ZERR_CHECKSUM;
				r.consumed += 4;
			}
		}
	}
	if (lane == 0) {
		results[idx] = r;
		w->cmd = 0;
#ifdef WIDE_PROF
		WPROF(0);
		printf("wide: lanes decoding in pass k, per round:");
		for (int k = 0; k < 13; k++) printf(" %.1f", (double) w->runs[k] / w->pt[8]);
		printf("\n");
		printf("wide: kcycles of pass k, per round:");
		for (int k = 0; k < 12; k++) printf(" %.1f", (double) w->pcyc[k] / 1e3 / w->pt[8]);
		printf("\n");
		printf("wide: expand of warp 0, kcycles/round: flat %.1f long %.1f\n", w->pt[12] / 1e3 / w->pt[8], w->pt[13] / 1e3 / w->pt[8]);
		printf("wide: rounds %lld lanes/round %.1f passes/round %.2f sweeps/round %.2f bytes/round %.0f | kcycles/round: master %.1f stage %.1f decode %.1f chain %.1f expand %.1f resolve %.1f write %.1f\n",
		       w->pt[8], (double) w->pt[7] / w->pt[8], (double) w->pt[9] / w->pt[8], (double) w->pt[10] / w->pt[8], (double) w->pt[11] / w->pt[8],
		       w->pt[0] / 1e3 / w->pt[8], w->pt[1] / 1e3 / w->pt[8], w->pt[2] / 1e3 / w->pt[8], w->pt[3] / 1e3 / w->pt[8],
		       w->pt[4] / 1e3 / w->pt[8], w->pt[5] / 1e3 / w->pt[8], w->pt[6] / 1e3 / w->pt[8]);
#endif
	}
	__syncthreads();
}

extern "C" int jdb_inflate_wide(const uint8_t* src_base, uint8_t* dst_base,
                                const jdb_inflate_item* items, jdb_inflate_result* results,
                                jdb_inflate_state* states, uint32_t count, uint32_t format, uint32_t final, jdb_stream s)
{
	if (count == 0) return JDB_OK;
	const size_t smem = ((sizeof(WarpMem) + 15u) & ~(size_t) 15u) + sizeof(WideMem);
	JDB_CONFIGURE_SMEM(inflate_wide_kernel, smem);
	JDB_LAUNCH(inflate_wide_kernel, dim3(count), dim3(WIDE_LANES), smem, s,
	           src_base, dst_base, items, results, states, format, final);
	return jdb_rt_check_launch("inflate_wide_kernel");
}

extern "C" size_t jdb_inflate_state_bytes(void) { return sizeof(jdb_inflate_state); }

extern "C" int jdb_inflate_batch(const uint8_t* src_base, uint8_t* dst_base,
                                 const jdb_inflate_item* items, jdb_inflate_result* results,
                                 jdb_inflate_state* states, uint32_t count, uint32_t format,
                                 uint32_t final, uint32_t* counter, jdb_stream s)
{
	if (count == 0) return JDB_OK;
	int r = jdb_memset_async(counter, 0, sizeof(uint32_t), s);
	if (r != JDB_OK) return r;
	const size_t smem = sizeof(WarpMem) * INF_WARPS;
	JDB_CONFIGURE_SMEM(inflate_batch_kernel, smem);
	const uint32_t redo_only = 0;
	uint32_t ctas = (count + INF_WARPS - 1) / INF_WARPS;
	uint32_t cap = (uint32_t) jdb_rt_sm_count();
	if (ctas > cap) ctas = cap;
	JDB_LAUNCH(inflate_batch_kernel, dim3(ctas), dim3(INF_THREADS), smem, s,
	           src_base, dst_base, items, results, states, count, format, final, counter, redo_only, 0u);
	return jdb_rt_check_launch("inflate_batch_kernel");
}


/* ---------------------------------------------------------------------------
 * chunk discovery for the parallel decode of one large stream
 * ------------------------------------------------------------------------- */

/*
 * Every chunk our encoder writes (and every deflate sync flush) ends in the byte
 * aligned empty stored block .. 00 00 FF FF.  marker_scan_kernel lists the end
 * offsets of every occurrence of those four bytes -- candidates only: the same
 * bytes can occur inside compressed data.  The caller decodes from every
 * candidate (jdb_inflate_chunks) and keeps the chain that starts
 * at the true stream position and hops from marker to marker.
 */
__global__ void __launch_bounds__(256)
marker_scan_kernel(const uint8_t* __restrict__ src, uint64_t n, uint32_t* __restrict__ ends,
                   uint32_t max_ends, uint32_t* __restrict__ count)
{
	const uint64_t stride = (uint64_t) gridDim.x * blockDim.x * 4u;
	for (uint64_t base = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) * 4u; base + 4u <= n; base += stride) {
		/* four candidate offsets per thread from two aligned words (src is 4-byte aligned) */
		const uint32_t w0 = *(const uint32_t*) (src + base);
		uint32_t w1 = 0x01010101u;
		if (base + 8u <= n) w1 = *(const uint32_t*) (src + base + 4u);
		else for (uint32_t k = 0; k < 4; k++)
			if (base + 4u + k < n) w1 = (w1 & ~(0xffu << (8u * k))) | ((uint32_t) src[base + 4u + k] << (8u * k));
#pragma unroll
		for (uint32_t k = 0; k < 4; k++) {
			if (base + k + 4u > n) break;
			if (__funnelshift_r(w0, w1, 8u * k) == 0xffff0000u) {
				const uint32_t i = atomicAdd(count, 1u);
				if (i < max_ends) ends[i] = (uint32_t) (base + k + 4u);
			}
		}
	}
}

extern "C" int jdb_marker_scan(const uint8_t* src, uint64_t n, uint32_t* ends, uint32_t max_ends,
                               uint32_t* count, jdb_stream s)
{
	int r = jdb_memset_async(count, 0, sizeof(uint32_t), s);
	if (r != JDB_OK) return r;
	if (n < 4 || n > 0xfffffff0ull || ((uintptr_t) src & 3u)) return n < 4 ? JDB_OK : JDB_EARG;
	uint64_t threads = (n + 3) / 4;
	uint32_t ctas = (uint32_t) ((threads + 255) / 256);
	const uint32_t cap = (uint32_t) jdb_rt_sm_count() * 8;
	if (ctas > cap) ctas = cap;
	JDB_LAUNCH(marker_scan_kernel, dim3(ctas), dim3(256), 0, s, src, n, ends, max_ends, count);
	return jdb_rt_check_launch("marker_scan_kernel");
}

/* decode mode of the chunk-parallel path: like jdb_inflate_batch (one warp per item, output written),
 * but every item stops after the first empty stored block it reads -- status JDB_INF_ST_MARKER,
 * `error` = 1 when that block carried BFINAL */
extern "C" int jdb_inflate_chunks(const uint8_t* src_base, uint8_t* dst_base,
                                  const jdb_inflate_item* items, jdb_inflate_result* results,
                                  uint32_t count, uint32_t* counter, jdb_stream s)
{
	if (count == 0) return JDB_OK;
	int r = jdb_memset_async(counter, 0, sizeof(uint32_t), s);
	if (r != JDB_OK) return r;
	const size_t smem = sizeof(WarpMem) * INF_WARPS;
	JDB_CONFIGURE_SMEM(inflate_batch_kernel, smem);
	uint32_t ctas = (count + INF_WARPS - 1) / INF_WARPS;
	const uint32_t cap = (uint32_t) jdb_rt_sm_count();
	if (ctas > cap) ctas = cap;
	JDB_LAUNCH(inflate_batch_kernel, dim3(ctas), dim3(INF_THREADS), smem, s,
	           src_base, dst_base, items, results, (jdb_inflate_state*) 0, count, (uint32_t) JDB_FMT_RAW, 0u,
	           counter, 0u, 2u);
	return jdb_rt_check_launch("inflate_batch_kernel");
}
