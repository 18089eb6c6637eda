/*
 * jd_oracle.c -- CPU restatement of the jdeflate hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, loaded by
 * or shipped with the product (jdeflate_b200/lib/libjdeflate.so).  Only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may use it,
 * and there only as the checker.
 *
 * What it is: a plain C99, one-shot (non streaming) restatement of what the
 * reference computes on the compress / decompress / checksum path, each
 * function citing the reference file:line it follows (paths relative to the
 * reference tree):
 *
 *   jdo_inflate        raw DEFLATE decode with the reference's acceptance
 *                      rules and INFLT_* error codes      src/inflator.c
 *   jdo_deflate        levels 0-9, token-for-token the reference's parser,
 *                      Huffman construction and bit emission, so the output
 *                      is BYTE-IDENTICAL to deflator_deflate(DEFLT_END) on a
 *                      fresh instance given the whole input   src/deflator.c
 *   jdo_crc32_update, jdo_crc32_combine                       src/zstrm.c
 *   jdo_adler32_update RFC 1950 Adler-32 (the reference's C fallback is wrong
 *                      for ~44 % of sizes, src/zstrm.c:1386-1396; zlib's
 *                      adler32 is the oracle for this quantity)
 *
 * Pinning (tests/test_oracle.py): the reference ships no tests or golden
 * vectors, so the oracle is pinned against outputs of
 * the reference itself -- oracle/_ref/libjdeflate_ref.so, compiled from the
 * unmodified sources by oracle/Makefile -- against zlib 1.3, and against the
 * committed fixtures of tests/golden/ generated from that compiled reference
 * by tests/golden/make_golden.py.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define JDO_API __attribute__((visibility("default")))

/* result / error codes: jdeflate/inflator.h:48-66, jdeflate/deflator.h:48-70 */
enum { JDO_OK = 0, JDO_SRCEXHSTD = 1, JDO_TGTEXHSTD = 2, JDO_ERROR = 3 };
enum {
	JDO_EBADSTATE = 1, JDO_EBADCODE = 2, JDO_EBADTREE = 3, JDO_EFAROFFSET = 4,
	JDO_EBADBLOCK = 5, JDO_EINPUTEND = 6
};


/* ==========================================================================
 * Checksums
 * ========================================================================== */

#define CRCPOLY 0xEDB88320u

static uint32_t crc_tab[8][256];
static uint32_t crc_zop[32][32];    /* operator "append 2^i zero bytes" */
static int crc_ready;

static uint32_t
gf2_apply(const uint32_t* m, uint32_t v)
{
	/* src/zstrm.c:1413-1425 (GF2_matrixtimes) */
	uint32_t r = 0;
	for (; v; v >>= 1, m++) {
		if (v & 1u) r ^= *m;
	}
	return r;
}

static void
crc_setup(void)
{
	uint32_t i, k;

	if (crc_ready) return;
	for (i = 0; i < 256; i++) {
		uint32_t c = i;
		for (k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ CRCPOLY : (c >> 1);
		crc_tab[0][i] = c;
	}
	for (k = 1; k < 8; k++) {
		for (i = 0; i < 256; i++) {
			uint32_t c = crc_tab[k - 1][i];
			crc_tab[k][i] = (c >> 8) ^ crc_tab[0][c & 0xff];
		}
	}
	/* operator for one zero byte: row b = image of bit b; then square it
	 * (the reference stores these 32 matrices precomputed,
	 * src/zstrm.c:2027-2319) */
	for (i = 0; i < 32; i++) {
		uint32_t c = 1u << i;
		for (k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ CRCPOLY : (c >> 1);
		crc_zop[0][i] = c;
	}
	for (k = 1; k < 32; k++) {
		for (i = 0; i < 32; i++) {
			crc_zop[k][i] = gf2_apply(crc_zop[k - 1], crc_zop[k - 1][i]);
		}
	}
	crc_ready = 1;
}

/* src/zstrm.c:1489-1526: slice-by-8 over 32-bit little endian loads; takes
 * and returns the non-finalised register */
JDO_API uint32_t
jdo_crc32_update(uint32_t crc, const uint8_t* p, size_t n)
{
	crc_setup();
	for (; n && ((uintptr_t) p & 7); n--) {
		crc = (crc >> 8) ^ crc_tab[0][(crc ^ *p++) & 0xff];
	}
	for (; n >= 8; n -= 8, p += 8) {
		uint32_t a, b;
		memcpy(&a, p, 4);
		memcpy(&b, p + 4, 4);
		a ^= crc;
		crc = crc_tab[7][a & 0xff] ^ crc_tab[6][(a >> 8) & 0xff] ^
		      crc_tab[5][(a >> 16) & 0xff] ^ crc_tab[4][a >> 24] ^
		      crc_tab[3][b & 0xff] ^ crc_tab[2][(b >> 8) & 0xff] ^
		      crc_tab[1][(b >> 16) & 0xff] ^ crc_tab[0][b >> 24];
	}
	for (; n; n--) {
		crc = (crc >> 8) ^ crc_tab[0][(crc ^ *p++) & 0xff];
	}
	return crc;
}

/* src/zstrm.c:1427-1443 (crc32_ncombine): finalised crcs; walks the set bits
 * of len2 applying the "2^i zero bytes" operators.  64-bit length here. */
JDO_API uint32_t
jdo_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2)
{
	int i;

	crc_setup();
	for (i = 0; len2; i++, len2 >>= 1) {
		if (len2 & 1u) {
			if (i < 32) {
				crc1 = gf2_apply(crc_zop[i], crc1);
			} else {
				/* 2^i zero bytes for i >= 32: apply 2^31 twice per doubling */
				uint64_t reps = (uint64_t) 1 << (i - 31);
				while (reps--) crc1 = gf2_apply(crc_zop[31], crc1);
			}
		}
	}
	return crc1 ^ crc2;
}

/* RFC 1950 section 8.2 / zlib adler32; src/zstrm.c:1346-1399 is the (defective)
 * reference counterpart */
JDO_API uint32_t
jdo_adler32_update(uint32_t adler, const uint8_t* p, size_t n)
{
	uint32_t a = adler & 0xffff, b = (adler >> 16) & 0xffff;

	while (n) {
		size_t k = n < 5552 ? n : 5552;
		n -= k;
		while (k--) {
			a += *p++;
			b += a;
		}
		a %= 65521u;
		b %= 65521u;
	}
	return (b << 16) | a;
}


/* ==========================================================================
 * Inflate
 * ========================================================================== */

/* length / distance base + extra bits: src/inflator.c:336-373 */
static const uint16_t len_base[29] = {
	3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
	67, 83, 99, 115, 131, 163, 195, 227, 258
};
static const uint8_t len_extra[29] = {
	0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0
};
static const uint16_t dist_base[30] = {
	1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
	1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577
};
static const uint8_t dist_extra[30] = {
	0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13
};
static const uint8_t precode_order[19] = {
	16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15
};

typedef struct {
	uint16_t count[16];    /* codes of each length     */
	uint16_t symbol[288];  /* symbols ordered by code  */
	int      empty;
} jdo_hcode;

enum { HC_LITLEN, HC_DIST, HC_PRECODE };

/*
 * Validation rules of buildtable (src/inflator.c:380-474): all-zero lengths
 * are an error except for the distance code; over-subscribed sets are an
 * error; incomplete sets are an error unless it is a distance code whose
 * longest (only) length is 1.
 */
static int
hcode_build(jdo_hcode* h, const uint16_t* lengths, int n, int kind)
{
	int offs[16];
	int i, left, mlen;

	memset(h->count, 0, sizeof(h->count));
	h->empty = 0;
	for (i = 0; i < n; i++) h->count[lengths[i]]++;
	if (h->count[0] == n) {
		if (kind == HC_DIST) {
			h->empty = 1;
			return 0;
		}
		return -1;
	}
	h->count[0] = 0;
	for (mlen = 15; h->count[mlen] == 0; mlen--);
	left = 1;
	for (i = 1; i <= 15; i++) {
		left = (left << 1) - h->count[i];
		if (left < 0) return -1;
	}
	if (left) {
		if (mlen != 1 || kind != HC_DIST) return -1;
	}
	offs[1] = 0;
	for (i = 1; i < 15; i++) offs[i + 1] = offs[i] + h->count[i];
	for (i = 0; i < n; i++) {
		if (lengths[i]) h->symbol[offs[lengths[i]]++] = (uint16_t) i;
	}
	return 0;
}

typedef struct {
	const uint8_t* src;
	size_t n, pos;
	uint64_t bb;
	unsigned bc;
} jdo_bits;

/* make at least `need` (<= 32) bits available; 0 when the input ends first */
static int
bits_need(jdo_bits* b, unsigned need)
{
	while (b->bc < need) {
		if (b->pos >= b->n) return 0;
		b->bb |= (uint64_t) b->src[b->pos++] << b->bc;
		b->bc += 8;
	}
	return 1;
}

static uint32_t
bits_take(jdo_bits* b, unsigned k)
{
	uint32_t v = (uint32_t) (b->bb & (((uint64_t) 1 << k) - 1));
	b->bb >>= k;
	b->bc -= k;
	return v;
}

/*
 * Canonical decode, one bit at a time (equivalent to the two level tables of
 * src/inflator.c:476-568).  Returns the symbol, -1 when the input ended inside
 * the code, -2 for a bit pattern no code uses (reference: table entry with
 * length 0 -> INFLT_EBADCODE, src/inflator.c:1426-1431, 1630-1634).
 */
static int
hcode_decode(jdo_bits* b, const jdo_hcode* h)
{
	int code = 0, first = 0, index = 0, len;
	jdo_bits save = *b;

	if (h->empty) return -2;
	for (len = 1; len <= 15; len++) {
		int count;
		if (!bits_need(b, 1)) {
			*b = save;
			return -1;
		}
		code |= (int) bits_take(b, 1);
		count = h->count[len];
		if (code - count < first) return h->symbol[index + (code - first)];
		index += count;
		first += count;
		first <<= 1;
		code <<= 1;
	}
	return -2;
}

/*
 * One-shot decode of a raw DEFLATE stream (inflator_inflate on a fresh
 * instance: src/inflator.c:764-903).  `final` says no more input can follow.
 * Outputs: *consumed = bytes of src used (whole bytes still unread are given
 * back, i.e. the exact end of the stream), *produced, *error (INFLT_E*).
 */
JDO_API int
jdo_inflate(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, int final,
            size_t* consumed, size_t* produced, int* error)
{
	jdo_bits b;
	jdo_hcode lit, dist, pre;
	uint16_t lengths[320];
	size_t out = 0;
	int last = 0, status = JDO_OK, err = 0;
	jdo_bits mark;

	b.src = src; b.n = n; b.pos = 0; b.bb = 0; b.bc = 0;
	mark = b;

#define FAIL(E) do { err = (E); status = JDO_ERROR; goto done; } while (0)
#define STARVED() do { b = mark; if (final) FAIL(JDO_EINPUTEND); status = JDO_SRCEXHSTD; goto done; } while (0)

	while (!last) {
		uint32_t type;

		mark = b;
		/* block header, src/inflator.c:829-851 */
		if (!bits_need(&b, 3)) STARVED();
		last = (int) bits_take(&b, 1);
		type = bits_take(&b, 2);

		if (type == 0) {
			/* stored, src/inflator.c:930-1019 */
			uint32_t len, nlen;
			bits_take(&b, b.bc & 7);
			if (!bits_need(&b, 32)) STARVED();
			len = bits_take(&b, 16);
			nlen = bits_take(&b, 16);
			if ((uint16_t) ~len != nlen) FAIL(JDO_EBADBLOCK);
			/* bb is empty now (bc is a multiple of 8 and was drained) */
			while (len) {
				if (b.bc) {
					if (out >= cap) { status = JDO_TGTEXHSTD; goto done; }
					dst[out++] = (uint8_t) bits_take(&b, 8);
					len--;
					continue;
				}
				if (b.pos >= b.n) {
					if (final) FAIL(JDO_EINPUTEND);
					status = JDO_SRCEXHSTD;
					goto done;
				}
				if (out >= cap) { status = JDO_TGTEXHSTD; goto done; }
				dst[out++] = b.src[b.pos++];
				len--;
			}
			continue;
		}
		if (type == 3) FAIL(JDO_EBADBLOCK);        /* src/inflator.c:888 */

		if (type == 1) {
			/* fixed codes, src/inflator.c:685-726 */
			int i;
			for (i = 0; i < 144; i++) lengths[i] = 8;
			for (; i < 256; i++) lengths[i] = 9;
			for (; i < 280; i++) lengths[i] = 7;
			for (; i < 288; i++) lengths[i] = 8;
			hcode_build(&lit, lengths, 288, HC_LITLEN);
			for (i = 0; i < 32; i++) lengths[i] = 5;
			hcode_build(&dist, lengths, 32, HC_DIST);
		} else {
			/* dynamic header, src/inflator.c:1103-1190 */
			uint32_t hlit, hdist, hclen, i;
			uint16_t plen[19];

			if (!bits_need(&b, 14)) STARVED();
			hlit = bits_take(&b, 5) + 257;
			hdist = bits_take(&b, 5) + 1;
			hclen = bits_take(&b, 4) + 4;
			if (hlit > 286 || hdist > 30) FAIL(JDO_EBADTREE);
			memset(plen, 0, sizeof(plen));
			for (i = 0; i < hclen; i++) {
				if (!bits_need(&b, 3)) STARVED();
				plen[precode_order[i]] = (uint16_t) bits_take(&b, 3);
			}
			if (hcode_build(&pre, plen, 19, HC_PRECODE)) FAIL(JDO_EBADTREE);

			/* readlengths, src/inflator.c:1029-1101 */
			for (i = 0; i < hlit + hdist;) {
				int sym = hcode_decode(&b, &pre);
				uint32_t rep, val;
				if (sym == -1) STARVED();
				if (sym < 0) FAIL(JDO_EBADCODE);
				if (sym < 16) {
					lengths[i++] = (uint16_t) sym;
					continue;
				}
				if (sym == 16) {
					if (!bits_need(&b, 2)) STARVED();
					rep = 3 + bits_take(&b, 2);
					if (i == 0) FAIL(JDO_EBADTREE);
					val = lengths[i - 1];
				} else if (sym == 17) {
					if (!bits_need(&b, 3)) STARVED();
					rep = 3 + bits_take(&b, 3);
					val = 0;
				} else {
					if (!bits_need(&b, 7)) STARVED();
					rep = 11 + bits_take(&b, 7);
					val = 0;
				}
				/* the reference bounds runs by the array size (320), not by
				 * hlit + hdist: src/inflator.c:1090-1093 */
				if (i + rep > 320) FAIL(JDO_EBADTREE);
				while (rep--) lengths[i++] = (uint16_t) val;
			}
			if (lengths[256] == 0) FAIL(JDO_EBADTREE);   /* :1171-1174 */
			if (hcode_build(&lit, lengths, (int) hlit, HC_LITLEN)) FAIL(JDO_EBADTREE);
			if (hcode_build(&dist, lengths + hlit, (int) hdist, HC_DIST)) FAIL(JDO_EBADTREE);
		}

		/* symbols, src/inflator.c:1329-1518 / 1529-1823 */
		for (;;) {
			int sym;
			uint32_t len, d, eb;

			mark = b;
			sym = hcode_decode(&b, &lit);
			if (sym == -1) STARVED();
			if (sym < 0) FAIL(JDO_EBADCODE);
			if (sym < 256) {
				if (out >= cap) { b = mark; status = JDO_TGTEXHSTD; goto done; }
				dst[out++] = (uint8_t) sym;
				continue;
			}
			if (sym == 256) break;
			if (sym > 285) FAIL(JDO_EBADCODE);     /* lnsinfo entries 286/287 have length 0 */
			eb = len_extra[sym - 257];
			if (!bits_need(&b, eb)) STARVED();
			len = len_base[sym - 257] + bits_take(&b, eb);

			sym = hcode_decode(&b, &dist);
			if (sym == -1) STARVED();
			if (sym < 0 || sym > 29) FAIL(JDO_EBADCODE);
			eb = dist_extra[sym];
			if (!bits_need(&b, eb)) STARVED();
			d = dist_base[sym] + bits_take(&b, eb);

			if (d > out) FAIL(JDO_EFAROFFSET);     /* no window: fresh instance */
			while (len) {
				if (out >= cap) { status = JDO_TGTEXHSTD; goto done; }
				dst[out] = dst[out - d];
				out++;
				len--;
			}
		}
	}

done:
	/* whole bytes still sitting in the bit buffer were not consumed */
	b.pos -= b.bc >> 3;
	if (consumed) *consumed = b.pos;
	if (produced) *produced = out;
	if (error) *error = err;
	return status;
#undef FAIL
#undef STARVED
}


/* ==========================================================================
 * Deflate
 * ========================================================================== */

#define WND      32768
#define MINMATCH 3
#define MAXMATCH 258
#define LOOKAHEAD (MINMATCH + MAXMATCH)     /* src/deflator.c:2329 */
#define GUARD    304                         /* src/deflator.c:320-324 */

typedef struct {
	/* output bits (src/deflator.c:563-607) */
	uint8_t* dst;
	size_t   cap, out;
	uint64_t bb;
	unsigned bc;
	int      overflow;

	int      level, fixedonly;
	uint32_t good, nice, chain;

	/* window, src/deflator.c:1817-1897 */
	uint8_t* win;
	size_t   wsize;
	size_t   inend;      /* filled part of the window */
	size_t   cursor;
	int64_t  whence3, whence4;
	const uint8_t* src;
	size_t   srcleft;

	int16_t*  h4head;    /* 65536 */
	int16_t*  h4prev;    /* 32768 */
	uint16_t* h3head;    /* 16384 */
	uint16_t* h3prev;    /* 16384 */

	uint16_t* tok;
	size_t    ntok, tokcap;

	uint32_t lfreq[288], dfreq[32], cfreq[19];

	/* block split statistics, src/deflator.c:2527-2596 */
	uint32_t cur[32], prv[32];
	uint32_t obscount, newcount, obstotal;
} jdo_enc;

static void
put(jdo_enc* e, uint32_t bits, unsigned n)
{
	e->bb |= (uint64_t) bits << e->bc;
	e->bc += n;
	while (e->bc >= 8) {
		if (e->out < e->cap) e->dst[e->out] = (uint8_t) e->bb;
		else e->overflow = 1;
		e->out++;
		e->bb >>= 8;
		e->bc -= 8;
	}
}

static void
put_align(jdo_enc* e)
{
	if (e->bc) put(e, 0, 8 - e->bc);
}

static uint32_t
revbits(uint32_t code, unsigned len)
{
	uint32_t r = 0;
	while (len--) {
		r = (r << 1) | (code & 1u);
		code >>= 1;
	}
	return r;
}

/* level -> (good, nice, chain): src/deflator.c:241-263 */
static void
set_level(jdo_enc* e, int level)
{
	static const uint16_t t[10][3] = {
		{0, 0, 0}, {8, 4, 2}, {8, 8, 8}, {8, 16, 16}, {8, 32, 32}, {8, 64, 128},
		{16, 16, 48}, {32, 64, 128}, {64, 128, 320}, {192, 256, 512}
	};
	e->level = level;
	e->good = t[level][0];
	e->nice = t[level][1];
	e->chain = t[level][2];
}

/* ---- Huffman construction ------------------------------------------------ */

typedef struct { uint8_t len; uint16_t code; } jdo_code;

/* Moffat & Katajainen, "In-place calculation of minimum-redundancy codes"
 * (src/deflator.c:1032-1081): a[] holds n frequencies in ascending order and
 * is overwritten by code lengths. */
static void
mr_lengths(uint32_t* a, int n)
{
	int root, leaf, next, avail, used, depth;

	if (n == 1) { a[0] = 1; return; }
	/* phase 1: combine */
	root = 0; leaf = 0;
	for (next = 0; next < n - 1; next++) {
		if (leaf >= n || (root < next && a[root] < a[leaf])) {
			a[next] = a[root];
			a[root++] = (uint32_t) next;
		} else {
			a[next] = a[leaf++];
		}
		if (leaf >= n || (root < next && a[root] < a[leaf])) {
			a[next] += a[root];
			a[root++] = (uint32_t) next;
		} else {
			a[next] += a[leaf++];
		}
	}
	/* phases 2+3 fused the way the reference does it: walk the internal
	 * nodes from the root counting how many sit at each depth */
	{
		int prev = n - 2, tree = n - 2, k = n - 1;
		avail = 2;
		for (depth = 1; k > 0; depth++) {
			int j;
			for (used = 0; tree && a[tree - 1] >= (uint32_t) prev;) {
				tree--;
				used++;
			}
			for (j = avail - used; j; j--) a[k--] = (uint32_t) depth;
			avail = used << 1;
			prev = tree;
		}
	}
}

/* Kraft repair after clamping, src/deflator.c:991-1028 */
static void
limit_lengths(uint32_t* len, int n, uint32_t maxlen)
{
	int64_t k = 0;
	int i;

	for (i = 0; i < n; i++) {
		if (len[i] > maxlen) len[i] = maxlen;
		k += (int64_t) 1 << (15 - len[i]);
	}
	for (i = 0; i < n; i++) {
		while (len[i] < maxlen && k > 0x8000) {
			len[i]++;
			k -= (int64_t) 1 << (15 - len[i]);
		}
	}
	for (i = n - 1; i >= 0; i--) {
		while (k + ((int64_t) 1 << (15 - len[i])) <= 0x8000) {
			k += (int64_t) 1 << (15 - len[i]);
			len[i]--;
		}
	}
}

/*
 * setuptable + computelengths (src/deflator.c:1138-1285): freq[] (size n) is
 * replaced by code lengths; codes[] receives bit-reversed canonical codes.
 * Returns last used symbol + 1.
 */
static int
make_code(uint32_t* freq, int n, uint32_t maxlen, jdo_code* codes)
{
	int map[288];
	uint32_t work[288];
	uint16_t count[16], next[16];
	int used = 0, i, j, last = 0;

	for (i = 0; i < n; i++) used += freq[i] != 0;
	/* at least two codes, src/deflator.c:1149-1162 */
	if (used == 0) {
		freq[0] = freq[1] = 1;
	} else if (used == 1) {
		if (freq[0]) freq[1] = 1;
		else freq[0] = 1;
	}
	used = 0;
	for (i = 0; i < n; i++) {
		if (freq[i]) map[used++] = i;
	}
	/* ascending by (frequency, symbol): what the reference's heapsort yields
	 * (src/deflator.c:933-989); insertion sort on a total order is identical */
	for (i = 1; i < used; i++) {
		int s = map[i];
		for (j = i; j > 0; j--) {
			int t = map[j - 1];
			if (freq[t] < freq[s] || (freq[t] == freq[s] && t < s)) break;
			map[j] = t;
		}
		map[j] = s;
	}
	for (i = 0; i < used; i++) work[i] = freq[map[i]];
	mr_lengths(work, used);
	limit_lengths(work, used, maxlen);

	memset(count, 0, sizeof(count));
	for (i = 0; i < used; i++) {
		count[work[i]]++;
		freq[map[i]] = work[i];
	}
	next[0] = 0;
	for (i = 1; i <= 15; i++) next[i] = (uint16_t) ((count[i - 1] + next[i - 1]) << 1);
	for (i = 0; i < n; i++) {
		uint32_t l = freq[i];
		codes[i].len = (uint8_t) l;
		codes[i].code = 0;
		if (l == 0) continue;
		codes[i].code = (uint16_t) revbits(next[l]++, l);
		last = i;
	}
	return last + 1;
}

/* countprecodes, src/deflator.c:1287-1354: run-length code one length array in
 * place; a[size] is the element after the last used symbol (always 0 here) */
static void
rle_lengths(uint32_t* a, int size, uint32_t* cfreq)
{
	uint32_t p = 0xffff, n, s;
	int i, j = 0, count = 0, maxrun = 0, breakrun;

	a[size + 1] = 0xffff;
	for (i = 0; i <= size; i++) {
		n = a[i];
		if (n == p) {
			count++;
			if (count < maxrun) continue;
			breakrun = 1;
		} else {
			breakrun = 0;
		}
		if (count > 2) {
			s = p ? 16 : (count > 10 ? 18 : 17);
			cfreq[s]++;
			a[j++] = s;
			a[j++] = (uint32_t) count;
			if (breakrun) {
				count = 0;
				continue;
			}
		} else if (count) {
			cfreq[p] += (uint32_t) count;
			while (count) {
				a[j++] = p;
				count--;
			}
		}
		cfreq[n]++;
		maxrun = n ? 6 : 136;
		a[j++] = p = n;
		count = 0;
	}
	a[j - 1] = 0xffff;
}

static void
emit_rle(jdo_enc* e, const uint32_t* a, const jdo_code* pc)
{
	/* emittrees L_STATE2, src/deflator.c:1674-1713 */
	int i = 0;
	while (a[i] != 0xffff) {
		uint32_t s = a[i++];
		put(e, pc[s].code, pc[s].len);
		if (s == 16) put(e, a[i++] - 3, 2);
		else if (s == 17) put(e, a[i++] - 3, 3);
		else if (s == 18) put(e, a[i++] - 11, 7);
	}
}

/* distance / length symbol: src/deflator.c:2169-2284 */
static uint32_t
dsym_of(uint32_t d)
{
	uint32_t s = 0;
	while (s < 29 && dist_base[s + 1] <= d) s++;
	return s;
}

static uint32_t
lsym_of(uint32_t l)
{
	uint32_t s = 0;
	while (s < 28 && len_base[s + 1] <= l) s++;
	return s;
}

/* flushblock, src/deflator.c:1724-1805 */
static void
flush_block(jdo_enc* e)
{
	jdo_code lit[288], dst[32], pre[19];
	uint32_t total = (uint32_t) e->ntok;
	int dynamic, i;
	size_t t;

	if (total == 0) return;
	e->lfreq[256]++;
	e->tok[e->ntok++] = 256;

	dynamic = !(e->level == 1 || e->fixedonly || total < 0x400);
	put(e, 0, 1);
	put(e, dynamic ? 2 : 1, 2);

	if (dynamic) {
		/* buildtables, src/deflator.c:1361-1390 */
		int lmax = make_code(e->lfreq, 288, 15, lit);
		int dmax = make_code(e->dfreq, 32, 15, dst);
		int cmax;
		memset(e->cfreq, 0, sizeof(e->cfreq));
		/* both arrays need two slack slots after the last symbol */
		{
			uint32_t la[290], da[34];
			memcpy(la, e->lfreq, sizeof(e->lfreq));
			memcpy(da, e->dfreq, sizeof(e->dfreq));
			la[288] = la[289] = 0;
			da[32] = da[33] = 0;
			rle_lengths(la, lmax, e->cfreq);
			rle_lengths(da, dmax, e->cfreq);
			make_code(e->cfreq, 19, 7, pre);
			for (cmax = 18; cmax >= 3; cmax--) {
				if (e->cfreq[precode_order[cmax]]) break;
			}
			cmax++;
			/* emittrees, src/deflator.c:1633-1722 */
			put(e, (uint32_t) lmax - 257, 5);
			put(e, (uint32_t) dmax - 1, 5);
			put(e, (uint32_t) cmax - 4, 4);
			for (i = 0; i < cmax; i++) put(e, e->cfreq[precode_order[i]], 3);
			emit_rle(e, la, pre);
			emit_rle(e, da, pre);
		}
	} else {
		/* fixed codes, RFC 1951 3.2.6 (tables at src/deflator.c:2987-3110) */
		for (i = 0; i < 288; i++) {
			unsigned l = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
			uint32_t c = i < 144 ? 0x30 + i : i < 256 ? 0x190 + (i - 144)
			           : i < 280 ? (uint32_t) (i - 256) : 0xc0 + (i - 280);
			lit[i].len = (uint8_t) l;
			lit[i].code = (uint16_t) revbits(c, l);
		}
		for (i = 0; i < 32; i++) {
			dst[i].len = 5;
			dst[i].code = (uint16_t) revbits((uint32_t) i, 5);
		}
	}

	/* emitlz, src/deflator.c:1421-1631 */
	for (t = 0; t < e->ntok;) {
		uint16_t a = e->tok[t];
		if (a < 0x8000) {
			put(e, lit[a].code, lit[a].len);
			t++;
		} else {
			uint32_t len = a - 0x8000u, d = e->tok[t + 1];
			uint32_t ls = e->tok[t + 2] >> 8, ds = e->tok[t + 2] & 0xff;
			put(e, lit[257 + ls].code, lit[257 + ls].len);
			if (len_extra[ls]) put(e, len - len_base[ls], len_extra[ls]);
			put(e, dst[ds].code, dst[ds].len);
			if (dist_extra[ds]) put(e, d - dist_base[ds], dist_extra[ds]);
			t += 3;
		}
	}
	e->ntok = 0;
}

/* endstream, src/deflator.c:609-654 */
static void
end_stream(jdo_enc* e, int final)
{
	put(e, final ? 1 : 0, 1);
	put(e, 0, 2);
	put_align(e);
	put(e, 0x0000, 16);
	put(e, 0xffff, 16);
}

/* ---- window ---------------------------------------------------------------- */

/* fillwindow + slidewindow, src/deflator.c:1817-1897 */
static size_t
fill_window(jdo_enc* e)
{
	size_t left = e->wsize - e->inend;
	size_t total = e->srcleft;

	if (total > left && left < 0x400) {
		size_t from = e->cursor - WND;
		size_t r = ((uintptr_t) (e->win + from)) & 7;
		size_t moved;
		from -= r;
		moved = e->inend - from;
		memmove(e->win, e->win + from, moved);
		e->inend = moved;
		e->cursor = WND + r;
		e->whence3 -= (int64_t) from;
		e->whence4 -= (int64_t) from;
		left = e->wsize - e->inend;
	}
	if (total > left) total = left;
	if (total) {
		memcpy(e->win + e->inend, e->src, total);
		e->src += total;
		e->srcleft -= total;
		e->inend += total;
	}
	return total;
}

/* slidehash, src/deflator.c:1899-1911 */
static void
slide_hash(jdo_enc* e)
{
	int i;
	for (i = 0; i < 65536; i++) {
		int16_t v = e->h4head[i];
		e->h4head[i] = (int16_t) (0x8000 | (v & ~(v >> 15)));
	}
	for (i = 0; i < 32768; i++) {
		int16_t v = e->h4prev[i];
		e->h4prev[i] = (int16_t) (0x8000 | (v & ~(v >> 15)));
	}
}

/* gethead + gethash, src/deflator.c:1930-1947: big endian 4 bytes */
static uint32_t
head_at(const jdo_enc* e, size_t pos)
{
	const uint8_t* p = e->win + pos;
	return ((uint32_t) p[0] << 24) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 8) | p[3];
}

static uint32_t
hash_of(uint32_t head, unsigned bits)
{
	return (uint32_t) (head * 0x1e35a7bdu) >> (32 - bits);
}

/* getmatchlength, src/deflator.c:1977-2059: common prefix capped at 258.  The
 * reference compares 8 bytes at a time and may read past the cap; the result
 * is the exact prefix length when a difference exists inside the compared
 * words, else 258 -- the 8-byte stride makes lengths above 258 possible only
 * through the cap return, so a byte loop capped at 258 is NOT the same: the
 * loop checks the limit only every 16 bytes after the first 32. */
static uint32_t
match_len(const uint8_t* a, const uint8_t* b)
{
	uint32_t n = 0;
	/* first 32 bytes: four 8-byte words, exact */
	while (n < 32) {
		if (a[n] != b[n]) return n;
		n++;
	}
	for (;;) {
		uint32_t k;
		if (n >= 258) return 258;       /* limit test at the top of each pair */
		for (k = 0; k < 16; k++) {
			if (a[n + k] != b[n + k]) return n + k;
		}
		n += 16;
	}
}

typedef struct { uint32_t len, off; } jdo_match;

static void
insert4(jdo_enc* e, uint32_t h4, uint32_t* position4)
{
	uint32_t p4 = (uint16_t) ((int64_t) e->cursor - e->whence4);
	if (p4 == WND) {
		slide_hash(e);
		e->whence4 += WND;
		p4 = 0;
	}
	e->h4prev[p4 & 32767] = e->h4head[h4];
	e->h4head[h4] = (int16_t) p4;
	*position4 = p4;
}

/* getmatch1, src/deflator.c:2335-2400 */
static jdo_match
find1(jdo_enc* e, uint32_t length, uint32_t* hash)
{
	const uint8_t* s = e->win + e->cursor;
	const uint8_t* send = s + MAXMATCH;
	const uint8_t* best = s;
	uint32_t p4, chain;
	int16_t next, limit;
	jdo_match m;

	if (send > e->win + e->inend) send = e->win + e->inend;
	next = 0;
	{
		uint32_t h4 = hash[0];
		uint32_t q = (uint16_t) ((int64_t) e->cursor - e->whence4);
		if (q == WND) {
			slide_hash(e);
			e->whence4 += WND;
		}
		next = e->h4head[h4];
		insert4(e, h4, &p4);
	}
	hash[0] = hash_of(head_at(e, e->cursor + 1), 16);

	limit = (int16_t) (p4 - WND);
	for (chain = e->chain; chain; chain--) {
		const uint8_t* c;
		if (next <= limit) break;
		c = e->win + (e->whence4 + next);
		if (s[length] == c[length]) {
			uint32_t n = match_len(s, c);
			if (n > length) {
				length = n;
				best = c;
				if (length >= e->nice) break;
			}
		}
		next = e->h4prev[(uint32_t) next & 32767];
	}
	if (s + length > send) length -= (uint32_t) ((s + length) - send);
	m.len = length;
	m.off = (uint32_t) (s - best);
	return m;
}

/* skipbytes1, src/deflator.c:2402-2428 */
static void
skip1(jdo_enc* e, uint32_t skip, uint32_t total, uint32_t* hash)
{
	uint32_t h4 = hash[0], p4;
	for (; skip < total; skip++) {
		e->cursor++;
		insert4(e, h4, &p4);
		h4 = hash_of(head_at(e, e->cursor + 1), 16);
	}
	hash[0] = h4;
}

static void
add_literal(jdo_enc* e, uint32_t c)
{
	e->tok[e->ntok++] = (uint16_t) c;
	e->lfreq[c]++;
}

static void
add_match(jdo_enc* e, jdo_match m)
{
	/* addmatch, src/deflator.c:2293-2305 */
	uint32_t ls = lsym_of(m.len), ds = dsym_of(m.off);
	e->lfreq[257 + ls]++;
	e->dfreq[ds]++;
	e->tok[e->ntok++] = (uint16_t) (m.len | 0x8000);
	e->tok[e->ntok++] = (uint16_t) m.off;
	e->tok[e->ntok++] = (uint16_t) ((ls << 8) | ds);
}

static void
reset_freqs(jdo_enc* e)
{
	memset(e->lfreq, 0, sizeof(e->lfreq));
	memset(e->dfreq, 0, sizeof(e->dfreq));
}

/* where the parser has to stop for now: src/deflator.c:2450-2468, 2803-2821.
 * returns 0 when it may not run at all yet */
static int
parse_limit(const jdo_enc* e, size_t* limit)
{
	size_t lim = e->inend;
	if (lim - e->cursor > LOOKAHEAD + 1) {
		if (e->srcleft) lim -= LOOKAHEAD;
	} else if (e->srcleft) {
		lim = e->cursor;
	}
	/* one-shot: flush is DEFLT_END from the first call on */
	*limit = lim;
	return 1;
}

/* compress1, src/deflator.c:2430-2520 (greedy, levels 1-5) */
static void
run_greedy(jdo_enc* e)
{
	uint32_t hash[1] = {0};
	size_t limit;

	reset_freqs(e);
	for (;;) {
		parse_limit(e, &limit);
		while (limit > e->cursor) {
			jdo_match m = find1(e, MINMATCH, hash);
			if (m.len > MINMATCH) {
				add_match(e, m);
				skip1(e, 1, m.len, hash);
			} else {
				add_literal(e, e->win[e->cursor]);
			}
			e->cursor++;
			if (e->ntok + 4 > e->tokcap) {
				flush_block(e);
				reset_freqs(e);
			}
		}
		if (fill_window(e) == 0) break;
	}
	flush_block(e);
}

/* ---- lazy parser (levels 6-9) -------------------------------------------- */

static void
reset_obs(jdo_enc* e)
{
	memset(e->cur, 0, sizeof(e->cur));
	memset(e->prv, 0, sizeof(e->prv));
	e->obscount = e->newcount = e->obstotal = 0;
}

/* shouldsplit, src/deflator.c:2556-2596 */
static int
should_split(jdo_enc* e)
{
	int j;
	if (e->obscount > 0) {
		uint32_t delta = 0;
		for (j = 0; j < 32; j++) {
			uint32_t a = e->prv[j], b = e->cur[j];
			delta += a > b ? a - b : b - a;
		}
		if (delta >= 320 && e->obstotal >= 7168) {
			reset_obs(e);
			return 1;
		}
	}
	for (j = 0; j < 32; j++) {
		e->prv[j] = (e->prv[j] >> 1) + (e->cur[j] >> 1);
		e->cur[j] = 0;
	}
	e->obscount += e->newcount;
	e->newcount = 0;
	return 0;
}

static uint32_t
load3(const uint8_t* p)
{
	return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16);
}

/* getmatch2, src/deflator.c:2605-2725 */
static jdo_match
find2(jdo_enc* e, uint32_t length, uint32_t* hash, int shrt)
{
	const uint8_t* s = e->win + e->cursor;
	const uint8_t* send = s + MAXMATCH;
	const uint8_t* best = s;
	uint32_t p4, p3, chain, h3, h4;
	int16_t next4, limit;
	uint16_t next3;
	jdo_match m;

	if (send > e->win + e->inend) send = e->win + e->inend;

	p3 = (uint16_t) ((int64_t) e->cursor - e->whence3);
	h3 = hash[0];
	h4 = hash[1];
	{
		uint32_t q = (uint16_t) ((int64_t) e->cursor - e->whence4);
		if (q == WND) {
			slide_hash(e);
			e->whence4 += WND;
		}
	}
	next3 = e->h3head[h3];
	next4 = e->h4head[h4];
	insert4(e, h4, &p4);
	e->h3prev[p3 & 16383] = e->h3head[h3];
	e->h3head[h3] = (uint16_t) p3;

	{
		uint32_t head = head_at(e, e->cursor + 1);
		hash[0] = hash_of(head >> 8, 14);
		hash[1] = hash_of(head, 16);
	}

	chain = e->chain;
	if (length >= 3) chain >>= 1;

	limit = (int16_t) (p4 - WND);
	for (; chain; chain--) {
		const uint8_t* c;
		if (next4 <= limit) break;
		c = e->win + (e->whence4 + next4);
		if (s[length] == c[length]) {
			uint32_t n = match_len(s, c);
			if (n > length) {
				length = n;
				best = c;
				if (length >= e->nice) goto out;
			}
		}
		next4 = e->h4prev[(uint32_t) next4 & 32767];
	}

	/* hash-3 probes, src/deflator.c:2676-2711 */
	if (shrt && length < 3) {
		uint32_t s1 = load3(s);
		int probe;
		for (probe = 0; probe < 2; probe++) {
			uint32_t noff;
			const uint8_t* c;
			if (next3 == 0) break;
			noff = (uint16_t) (p3 - next3);
			if (noff > WND || noff == 0) break;
			c = s - noff;
			if (load3(c) == s1) {
				length = 3;
				best = c;
				break;
			}
			next3 = e->h3prev[next3 & 16383];
		}
	}
out:
	if (s + length > send) length -= (uint32_t) ((s + length) - send);
	m.len = length;
	m.off = (uint32_t) (s - best);
	return m;
}

/* skipbytes2, src/deflator.c:2729-2764 */
static void
skip2(jdo_enc* e, uint32_t skip, uint32_t total, uint32_t* hash)
{
	uint32_t h3 = hash[0], h4 = hash[1], p4, p3;
	for (; skip < total; skip++) {
		uint32_t head;
		e->cursor++;
		p3 = (uint16_t) ((int64_t) e->cursor - e->whence3);
		insert4(e, h4, &p4);
		e->h3prev[p3 & 16383] = e->h3head[h3];
		e->h3head[h3] = (uint16_t) p3;
		head = head_at(e, e->cursor + 1);
		h3 = hash_of(head >> 8, 14);
		h4 = hash_of(head, 16);
	}
	hash[0] = h3;
	hash[1] = h4;
}

static uint32_t
ilog2(uint32_t v)
{
	uint32_t r = 0;
	while (v >>= 1) r++;
	return r;
}

static void
obs_literal(jdo_enc* e, uint32_t c)
{
	e->cur[c >> 4]++;
	e->newcount++;
	e->obstotal++;
}

static void
obs_match(jdo_enc* e, jdo_match m)
{
	e->cur[16 + (lsym_of(m.len) >> 1)]++;
	e->newcount++;
	e->obstotal += m.len;
}

/* compress2, src/deflator.c:2766-2973 */
static void
run_lazy(jdo_enc* e)
{
	uint32_t hash[2] = {0, 0};
	jdo_match m = {0, 0}, prev;
	int hasmatch = 0, doshort = 0;
	size_t limit;

	reset_freqs(e);
	reset_obs(e);
	for (;;) {
		parse_limit(e, &limit);
		while (limit > e->cursor) {
			if (!hasmatch) {
				m = find2(e, MINMATCH - 1, hash, doshort);
				if (m.len == MINMATCH && m.off > 8192) m.len = MINMATCH - 1;
				if (m.len >= MINMATCH) {
					if (m.len >= e->good) {
						skip2(e, 1, m.len, hash);
						add_match(e, m);
						obs_match(e, m);
					} else {
						hasmatch = 1;
					}
				} else {
					uint32_t c = e->win[e->cursor];
					add_literal(e, c);
					obs_literal(e, c);
				}
			} else {
				int accept = 0;
				prev = m;
				m = find2(e, prev.len - 1, hash, 0);
				if (m.len >= prev.len) {
					int32_t d = (int32_t) m.len - (int32_t) prev.len;
					if (d > 4) {
						accept = 1;
					} else {
						int32_t l1 = (int32_t) ilog2(prev.off), l2 = (int32_t) ilog2(m.off);
						accept = (d << 2) + (l1 - l2) >= 2;
					}
				}
				if (accept) {
					uint32_t c = e->win[e->cursor - 1];
					add_literal(e, c);
					obs_literal(e, c);
				} else {
					skip2(e, 2, prev.len, hash);
					add_match(e, prev);
					obs_match(e, prev);
					hasmatch = 0;
				}
			}
			e->cursor++;

			if (e->ntok + 4 > e->tokcap) {
				reset_obs(e);
				flush_block(e);
				reset_freqs(e);
				reset_obs(e);
				continue;
			}
			if (e->newcount >= 512 && e->obstotal >= 4096) {
				doshort = e->cur[0] >= 16;
				if (should_split(e)) {
					flush_block(e);
					reset_freqs(e);
					reset_obs(e);
				}
			}
		}
		if (fill_window(e) == 0) break;
	}
	flush_block(e);
}

/* compress0, src/deflator.c:796-926: stored blocks of at most 65535 bytes */
static void
run_stored(jdo_enc* e, const uint8_t* src, size_t n)
{
	while (n) {
		size_t k = n < 65535 ? n : 65535;
		size_t i;
		put(e, 0, 3);
		put_align(e);
		put(e, (uint32_t) k, 16);
		put(e, (uint32_t) (~k & 0xffff), 16);
		for (i = 0; i < k; i++) put(e, src[i], 8);
		src += k;
		n -= k;
	}
}

/*
 * One-shot deflator_deflate(state, DEFLT_END) on a fresh instance
 * (src/deflator.c:690-786).  flags bit 0 = DEFLT_FIXEDCODES.  Returns 0, or -1
 * for a bad level, -2 when dst is too small (outlen then holds the needed
 * size), -3 on allocation failure.
 */
JDO_API int
jdo_deflate(int level, unsigned flags, const uint8_t* src, size_t n,
            uint8_t* dst, size_t cap, size_t* outlen)
{
	jdo_enc e;
	int rc = 0;

	if (level < 0 || level > 9) return -1;
	memset(&e, 0, sizeof(e));
	e.dst = dst;
	e.cap = cap;
	e.fixedonly = (flags & 1u) != 0;
	set_level(&e, level);

	if (level == 0) {
		run_stored(&e, src, n);
	} else {
		/* buffer sizes, src/deflator.c:209-230 */
		unsigned wbits = level > 5 ? 17 : 16;
		unsigned tbits = level == 1 ? 14 : level <= 5 ? 15 : level <= 7 ? 16 : 17;
		int i;

		e.wsize = (size_t) 1 << wbits;
		e.tokcap = (size_t) 1 << tbits;
		e.win = calloc(e.wsize + GUARD, 1);
		e.tok = malloc((e.tokcap + 8) * sizeof(uint16_t));
		e.h4head = malloc(65536 * sizeof(int16_t));
		e.h4prev = malloc(32768 * sizeof(int16_t));
		e.h3head = calloc(16384, sizeof(uint16_t));
		e.h3prev = calloc(16384, sizeof(uint16_t));
		if (!e.win || !e.tok || !e.h4head || !e.h4prev || !e.h3head || !e.h3prev) {
			rc = -3;
			goto out;
		}
		/* resetcache, src/deflator.c:418-441 */
		for (i = 0; i < 65536; i++) e.h4head[i] = (int16_t) -WND;
		for (i = 0; i < 32768; i++) e.h4prev[i] = (int16_t) -WND;
		e.src = src;
		e.srcleft = n;
		if (level <= 5) run_greedy(&e);
		else run_lazy(&e);
	}
	end_stream(&e, 1);
	if (e.overflow) rc = -2;
out:
	if (outlen) *outlen = e.out;
	free(e.win); free(e.tok); free(e.h4head); free(e.h4prev); free(e.h3head); free(e.h3prev);
	return rc;
}
