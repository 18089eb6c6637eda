/*
 * inflate.cu -- table driven DEFLATE decoder for sm_100a, one warp per stream.
 *
 * Replaces the reference's serial decoder: buildtable (src/inflator.c:380-568),
 * decodednmc/readlengths (:1029-1190), decodestrd (:930-1019), the hot loop
 * decodefast (:1529-1823) and the resumable decodeblock/copybytes
 * (:1213-1518), plus updatewindow (:616-675) for streaming use.
 *
 * Work decomposition
 *   A DEFLATE stream is bit-serial, so parallelism comes from many streams
 *   (BASELINE config 3: 1 M independent records; our own output: independent
 *   chunks).  Each warp owns one stream at a time and pulls the next stream
 *   index from a global counter (dynamic load balance for 4-64 KiB records).
 *
 *   Per warp, in shared memory: the two-level lookup tables (lit/len root
 *   10 bits, distance root 8 bits -- the reference's LROOTBITS/DROOTBITS,
 *   src/inflator.c:30-32), built warp-cooperatively per block, and a 32-entry
 *   symbol queue.  Lane 0 runs the bit-serial Huffman decode and fills the
 *   queue; then all 32 lanes turn the queue into bytes: a warp prefix sum
 *   gives every symbol its output offset, literals and short far matches are
 *   written by their own lane, long matches and matches that read bytes
 *   produced inside the same batch are copied warp-cooperatively in order.
 *   Output bytes are read back with ld.global.cg (L2) because they were
 *   written by other lanes of the warp.
 *
 *   Algorithmic traffic: C compressed bytes read + N bytes written per stream
 *   (match sources are re-read from L2).
 *
 * Error model: the INFLT_* codes of jdeflate/inflator.h with the acceptance
 * rules of the reference (see oracle/jd_oracle.c for the restatement); the two
 * places where the reference accepts an invalid stream (lit/len symbols
 * 286/287 and distance symbols 30/31 of the fixed code, SURVEY / DESIGN.md
 * "deviations") are rejected with INFLT_EBADCODE like zlib does.
 */
#include "common.cuh"

#ifndef INF_WARPS
#define INF_WARPS        16
#endif
#define INF_THREADS      (INF_WARPS * 32)
#ifndef LIT_ROOT
#define LIT_ROOT         10
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
struct BuildMem {
	uint16_t code[320];
	uint8_t  len[320];
	uint16_t count[16];
	uint16_t next[16];
	uint32_t scratch[8];
};

struct WarpMem {
	uint32_t lit[LIT_TABLE];
	uint32_t dist[DIST_TABLE];
	uint32_t queue[QUEUE];
	uint32_t inbuf[INW];           /* ring of staged input words (symbol loop) */
	BuildMem bm;
	uint8_t  ring[RING];
};

enum { KIND_LIT = 0, KIND_DIST = 1, KIND_PRE = 2 };

/*
 * Build one two-level table from code lengths in m->len[0..n).
 * Returns 0, or 1 when the length set is not acceptable (rules of
 * src/inflator.c:428-474).  All lanes call it; all lanes get the result.
 */
static __device__ int
build_table(BuildMem* m, uint32_t* table, int n, int kind, int lenoff)
{
	const unsigned lane = jdb_lane();
	const int root = kind == KIND_LIT ? LIT_ROOT : kind == KIND_DIST ? DIST_ROOT : 7;
	const int limit = kind == KIND_LIT ? LIT_TABLE : kind == KIND_DIST ? DIST_TABLE : 128;
	const uint8_t* len = m->len + lenoff;

	for (int i = lane; i < limit; i += 32) table[i] = 0;
	if (lane < 16) m->count[lane] = 0;
	__syncwarp();

	int rc = 0;
	if (lane == 0) {
		for (int i = 0; i < n; i++) m->count[len[i]]++;
		if (m->count[0] == n) {
			rc = kind == KIND_DIST ? 2 : 1;         /* 2: empty distance code is legal */
		} else {
			m->count[0] = 0;
			int mlen = 15;
			while (m->count[mlen] == 0) mlen--;
			int left = 1;
			for (int l = 1; l <= 15; l++) {
				left = (left << 1) - m->count[l];
				if (left < 0) { rc = 1; break; }
			}
			if (rc == 0 && left && !(mlen == 1 && kind == KIND_DIST)) rc = 1;
			if (rc == 0) {
				uint32_t c = 0;
				m->next[0] = 0;
				for (int l = 1; l <= 15; l++) {
					c = (c + m->count[l - 1]) << 1;
					m->next[l] = (uint16_t) c;
				}
				/* canonical code of every symbol, bit reversed (LSB-first stream) */
				for (int i = 0; i < n; i++) {
					int l = len[i];
					if (l) m->code[i] = (uint16_t) (__brev((uint32_t) m->next[l]++) >> (32 - l));
				}
			}
		}
	}
	rc = __shfl_sync(JDB_FULL_MASK, rc, 0);
	if (rc) return rc == 2 ? 0 : 1;

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
parse_block_header(BuildMem* bm, uint32_t* lit, uint32_t* dist, Bits& b, uint32_t& type_out, uint32_t& lb_out)
{
	const unsigned lane = jdb_lane();
	uint32_t r = 0;        /* 0 ok, 1 starved, 2+ : INFLT error code + 1 */
	uint32_t type = 0, lb = 0, hlit = 0, hdist = 0;
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
	uint8_t* ring;              /* shared-memory copy of the newest RING output bytes */
	int64_t  ring_lo;           /* positions >= ring_lo are served from the ring */
	uint32_t status, error;
};

/* byte at output position `pos` relative to this call's dst: the newest bytes
 * come from the shared-memory ring, older ones from L2, and negative positions
 * from the history ring of earlier calls */
static __device__ __forceinline__ uint8_t
out_byte(const Stream& s, int64_t pos)
{
	if (pos >= s.ring_lo) return s.ring[(uint32_t) pos & (RING - 1)];
	if (pos >= 0) return __ldcg(s.dst + pos);
	uint64_t abs = s.total_before + (uint64_t) pos;
	return __ldcg(s.st->history + (abs & (JDB_INFLATE_HISTORY - 1)));
}

static __device__ __forceinline__ void
put_byte(const Stream& s, uint64_t pos, uint8_t v)
{
	s.dst[pos] = v;
	if (s.ring) s.ring[(uint32_t) pos & (RING - 1)] = v;
}

/* warp-cooperative copy of `len` bytes with source `dist` back from `pos` */
static __device__ __forceinline__ void
copy_match(const Stream& s, uint64_t pos, uint32_t len, uint32_t dist)
{
	/* every source byte lies before `pos` (k < dist), so all loads can be
	 * issued before the first store */
	const unsigned lane = jdb_lane();
	uint8_t tmp[9];                      /* 258 bytes / 32 lanes */
#pragma unroll
	for (int i = 0; i < 9; i++) {
		uint32_t j = lane + 32u * i;
		if (j < len) {
			uint32_t k = dist >= len ? j : j % dist;
			tmp[i] = out_byte(s, (int64_t) (pos - dist) + k);
		}
	}
#pragma unroll
	for (int i = 0; i < 9; i++) {
		uint32_t j = lane + 32u * i;
		if (j < len) put_byte(s, pos + j, tmp[i]);
	}
}

/*
 * All lanes: turn `nq` queued symbols (literal: byte value; match: len << 16 |
 * dist) of stream `s` into output bytes starting at s.out.  A warp prefix sum
 * gives every symbol its offset; literals and short far matches are written by
 * their own lane, long matches and matches that read bytes produced by the
 * same queue are copied cooperatively, in order.  A match cut by the end of
 * the target is reported through pend_len / pend_dist.
 */
static __device__ __forceinline__ void
emit_queue(Stream& s, BuildMem* bm, const uint32_t* queue, uint32_t nq, uint32_t& pend_len, uint32_t& pend_dist)
{
	const unsigned lane = jdb_lane();
	uint32_t q = lane < nq ? queue[lane] : 0;
	uint32_t len = lane < nq ? ((q >> 16) ? (q >> 16) : 1u) : 0u;
	const bool is_match = lane < nq && (q >> 16) != 0;
	uint32_t dist = q & 0xffffu;
	uint32_t incl = len;
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(JDB_FULL_MASK, incl, o);
		if ((int) lane >= o) incl += t;
	}
	const uint64_t base = s.out;
	if (s.ring) {
		s.ring_lo = (int64_t) base - (int64_t) RING_KEEP;
		if (s.ring_lo < 0) s.ring_lo = 0;
	} else {
		s.ring_lo = 0x7fffffffffffffffll;     /* no shared-memory mirror: re-read from L2 */
	}
	const uint64_t pos = base + incl - len;
	const uint32_t total = __shfl_sync(JDB_FULL_MASK, incl, 31);
	/* clip the last symbol to the target capacity */
	uint32_t emit = len;
	if (pos + len > s.dst_cap) emit = (uint32_t) (s.dst_cap - pos);
	if (lane < nq && emit < len) {
		bm->scratch[5] = len - emit;
		bm->scratch[6] = dist;
	}
	const bool dependent = is_match && ((int64_t) pos - (int64_t) dist + (int64_t) len > (int64_t) base || dist < len);
	const bool longm = is_match && !dependent && emit > 16;
	if (lane < nq && !is_match) put_byte(s, pos, (uint8_t) q);
	if (is_match && !dependent && !longm) {
		/* short far match: all loads first, then the stores */
		uint8_t tmp[16];
#pragma unroll
		for (int j = 0; j < 16; j++)
			if ((uint32_t) j < emit) tmp[j] = out_byte(s, (int64_t) (pos - dist) + j);
#pragma unroll
		for (int j = 0; j < 16; j++)
			if ((uint32_t) j < emit) put_byte(s, pos + j, tmp[j]);
	}
	unsigned lm = __ballot_sync(JDB_FULL_MASK, longm);
	while (lm) {
		int src = __ffs(lm) - 1;
		lm &= lm - 1;
		uint64_t p2 = __shfl_sync(JDB_FULL_MASK, pos, src);
		uint32_t l2 = __shfl_sync(JDB_FULL_MASK, emit, src);
		uint32_t d2 = __shfl_sync(JDB_FULL_MASK, dist, src);
		copy_match(s, p2, l2, d2);
	}
	__syncwarp();
	unsigned dm = __ballot_sync(JDB_FULL_MASK, dependent);
	while (dm) {
		int src = __ffs(dm) - 1;
		dm &= dm - 1;
		uint64_t p2 = __shfl_sync(JDB_FULL_MASK, pos, src);
		uint32_t l2 = __shfl_sync(JDB_FULL_MASK, emit, src);
		uint32_t d2 = __shfl_sync(JDB_FULL_MASK, dist, src);
		copy_match(s, p2, l2, d2);
		__syncwarp();
	}
	uint64_t done = base + total;
	if (done > s.dst_cap) {
		done = s.dst_cap;
		pend_len = bm->scratch[5];
		pend_dist = bm->scratch[6];
	}
	s.out = done;
	__syncwarp();
}

/*
 * Decode one stream (or one call's worth of a streaming decode).
 * Called by all 32 lanes of a warp with identical arguments.
 */
static __device__ void
inflate_stream(WarpMem* m, Stream& s)
{
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

	/* finish a match that did not fit into the previous target window */
	if (pend_len) {
		uint32_t n = pend_len;
		if ((uint64_t) n > s.dst_cap) n = (uint32_t) s.dst_cap;
		copy_match(s, 0, n, pend_dist);
		__syncwarp();
		s.out = n;
		pend_len -= n;
		if (pend_len) { s.status = ST_TGTEXH; goto finish; }
	}

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
			if (!s.count_only)
				for (uint64_t j = lane; j < n; j += 32) put_byte(s, s.out + j, s.src[pos + j]);
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

			for (;;) {
				/* ---- all lanes: keep >= 3 KiBit of input ahead of `o` in the ring ---- */
				if (o >= filled) filled = o & ~1023u;
				while (filled < o + 3072u && filled < fill_limit) {
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

				/* ---- decode up to QUEUE symbols (uniform; lane 0 writes the queue) ---- */
				uint32_t nq = 0;
				uint32_t qbytes = 0;
				ev = 0;
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
								if (lane == 0) m->queue[nq] = e >> 16;
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
								if (lane == 0) m->queue[nq] = (flen << 16) | fdist;
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
							if (lane == 0) m->queue[nq] = e >> 16;    /* len field 0: literal */
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
						if (lane == 0) m->queue[nq] = (len << 16) | dist;       /* len <= 258, dist <= 32768 */
						nq++;
						qbytes += len;
						if (qbytes > room) { ev = 3; break; }        /* partially fits: split below */
					}
				}
				__syncwarp();

				/* ---- all lanes: turn the queue into bytes ---- */
				if (nq) {
					if (s.count_only) s.out += qbytes;
					else emit_queue(s, &m->bm, m->queue, nq, pend_len, pend_dist);
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

finish:
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
			inflate_stream(m, s);
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
