/*
 * checksum.cu -- chunk-parallel CRC-32 and Adler-32 for sm_100a.
 *
 * Replaces the reference's serial zstrm_crc32update (slice-by-8,
 * reference src/zstrm.c:1489-1526, asm/x86-64-crc32.asm) and
 * zstrm_adler32update (src/zstrm.c:1346-1399, asm/x86-64-adler32.asm) plus
 * the GF(2) combine crc32_ncombine (src/zstrm.c:1413-1443).
 *
 * HBM layout / algorithm
 *   The 16-byte aligned body of the input is cut into G contiguous spans, one
 *   per CTA.  Inside a span the 256 threads read 16-byte vectors interleaved
 *   (thread t takes vector t, t+256, ...: fully coalesced, no staging).  CRC
 *   is linear over GF(2), so every thread keeps four independent 32-bit
 *   accumulators (one per word slot of its vector) and advances them by one
 *   whole row (4096 bytes) with a 4-lookup table multiply by x^(8*4096) mod P
 *   before xoring the next word in.  The 1024 accumulators of the CTA then
 *   form a virtual 4 KiB row that is reduced the same way by warp 0 (row of
 *   128 bytes) and finally by one lane with the classic table step.  Each CTA
 *   multiplies its result by x^(8*bytes-after-span) so the combine across CTAs
 *   is a plain xor; Adler-32 partials are position weighted the same way so
 *   their combine is a plain sum modulo 65521.
 *   Algorithmic traffic: N bytes read, 16 bytes written per CTA.
 *
 *   Large inputs (>= CKW_MIN_BYTES) take ck_wide_kernel: the same scheme with
 *   1024 threads and one CTA per SM, whose row table is replicated 32 times in
 *   shared memory (entry e of lane l lives at word e * 32 + l, i.e. in bank l)
 *   so that the 16 data dependent lookups per vector never collide.  The
 *   256-thread kernel loses two thirds of its shared-memory wavefronts to bank
 *   conflicts (profiles/r1_checksum_kernel_ncu_summary.txt).
 */
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

#define CK_THREADS   256
#define CK_ROW_BYTES (CK_THREADS * 16)       /* 4096 */
#define CK_MAX_CTAS  4096
#define CRC_POLY     0xEDB88320u
#define ADLER_MOD    65521u

/* table set indices */
#define ZT_4    0      /* advance by 4 bytes    (classic slice-by-4 tables) */
#define ZT_ROW  1      /* advance by 4096 bytes */
#define ZT_W0   2      /* advance by 128 bytes  */
#define ZT_WROW 3      /* advance by 16384 bytes (row of the wide kernel) */
#define ZT_SETS 4

#define CKW_THREADS   1024
#define CKW_ROW_BYTES (CKW_THREADS * 16)     /* 16384 */
#define CKW_REP_WORDS (4 * 256 * 32)         /* row table, one copy per bank */
#define CKW_SMEM      ((CKW_REP_WORDS + CKW_THREADS * 4 + 3 * 1024) * 4)

struct CkTables {
	uint32_t z[ZT_SETS][4][256];
	uint32_t xpow[64];          /* x^(8*2^k) mod P, reflected */
};

struct CkPartial {
	uint32_t crc;       /* F(span) * x^(8*after) */
	uint32_t a;         /* sum of bytes mod 65521 */
	uint32_t b;         /* position weighted sum incl. bytes after the span */
	uint32_t pad;
};

/* ---- GF(2) polynomial helpers (reflected bit order) ------------------- */

static __host__ __device__ inline uint32_t gf2_mulmod(uint32_t a, uint32_t b)
{
	uint32_t p = 0;
	for (int i = 0; i < 32; i++) {
		if (a & (0x80000000u >> i)) p ^= b;
		b = (b & 1u) ? (b >> 1) ^ CRC_POLY : (b >> 1);
	}
	return p;
}

/* x^(8*n) mod P using the table of x^(8*2^k) */
static __host__ __device__ inline uint32_t gf2_xpow8(const uint32_t* xpow, uint64_t n)
{
	uint32_t r = 0x80000000u;   /* x^0 */
	for (int k = 0; n; k++, n >>= 1)
		if (n & 1u) r = gf2_mulmod(r, xpow[k]);
	return r;
}

static CkTables* g_dev_tables[64];

static void build_tables(CkTables* t)
{
	uint32_t t0[256];
	for (uint32_t i = 0; i < 256; i++) {
		uint32_t c = i;
		for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ CRC_POLY : (c >> 1);
		t0[i] = c;
	}
	uint32_t x8 = 0x80000000u;
	for (int k = 0; k < 8; k++) x8 = (x8 & 1u) ? (x8 >> 1) ^ CRC_POLY : (x8 >> 1);
	t->xpow[0] = x8;
	for (int k = 1; k < 64; k++) t->xpow[k] = gf2_mulmod(t->xpow[k - 1], t->xpow[k - 1]);

	const uint64_t adv[ZT_SETS] = { 4, CK_ROW_BYTES, 128, CKW_ROW_BYTES };
	for (int s = 0; s < ZT_SETS; s++) {
		uint32_t m = gf2_xpow8(t->xpow, adv[s]);
		for (int k = 0; k < 4; k++)
			for (uint32_t b = 0; b < 256; b++)
				t->z[s][k][b] = gf2_mulmod(b << (8 * k), m);
	}
	/* self check of the table identity against the byte-wise definition */
	uint32_t v = 0x12345678u, w = v;
	for (int i = 0; i < 4; i++) w = (w >> 8) ^ t0[w & 0xff];
	uint32_t z = t->z[ZT_4][0][v & 0xff] ^ t->z[ZT_4][1][(v >> 8) & 0xff] ^
	             t->z[ZT_4][2][(v >> 16) & 0xff] ^ t->z[ZT_4][3][v >> 24];
	if (z != w) { jdb_rt_set_error("checksum: crc table self check failed"); abort(); }
}

static const CkTables* device_tables(jdb_stream s)
{
	int dev = jdb_rt_get_device();
	if (dev < 0 || dev >= 64) dev = 0;
	if (g_dev_tables[dev]) return g_dev_tables[dev];
	CkTables* h = (CkTables*) malloc(sizeof(CkTables));
	if (!h) return NULL;
	build_tables(h);
	CkTables* d = (CkTables*) jdb_dev_alloc(sizeof(CkTables));
	if (!d) { free(h); return NULL; }
	if (jdb_copy_async(d, h, sizeof(CkTables), s) != JDB_OK || jdb_stream_sync(s) != JDB_OK) {
		free(h);
		jdb_dev_free(d);
		return NULL;
	}
	free(h);
	g_dev_tables[dev] = d;
	return d;
}

/* ---- kernels ----------------------------------------------------------- */

#define ZMUL(T, r) ((T)[0][(r) & 0xffu] ^ (T)[1][((r) >> 8) & 0xffu] ^ \
                    (T)[2][((r) >> 16) & 0xffu] ^ (T)[3][(r) >> 24])

template <bool DO_CRC, bool DO_ADLER>
__global__ void __launch_bounds__(CK_THREADS)
ck_partial_kernel(const uint4* __restrict__ body, uint64_t nvec, uint32_t nctas,
                  const CkTables* __restrict__ tables, CkPartial* __restrict__ out)
{
	__shared__ uint32_t zrow[4][256];
	__shared__ uint32_t z4[4][256];
	__shared__ uint32_t zw0[4][256];
	__shared__ uint32_t row[CK_THREADS * 4];
	__shared__ unsigned long long red[3][CK_THREADS / 32];

	const uint32_t tid = threadIdx.x;
	const uint32_t g = blockIdx.x;

	if (DO_CRC) {
		for (uint32_t i = tid; i < 1024; i += CK_THREADS) {
			(&zrow[0][0])[i] = (&tables->z[ZT_ROW][0][0])[i];
			(&z4[0][0])[i]   = (&tables->z[ZT_4][0][0])[i];
			(&zw0[0][0])[i]  = (&tables->z[ZT_W0][0][0])[i];
		}
	}
	__syncthreads();

	/* contiguous span of vectors owned by this CTA */
	const uint64_t v0 = nvec * g / nctas;
	const uint64_t v1 = nvec * (g + 1) / nctas;
	const uint64_t span = v1 - v0;
	const uint64_t rows = (span + CK_THREADS - 1) / CK_THREADS;
	/* right-align the span in a rows x 256 grid: leading zero vectors do not
	 * change a CRC that starts from register 0 */
	const uint64_t pad = rows * CK_THREADS - span;
	const uint64_t span_bytes = span * 16;

	uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
	unsigned long long sa = 0, sb = 0, sw = 0;

	for (uint64_t j = 0; j < rows; j++) {
		const uint64_t slot = j * CK_THREADS + tid;
		if (DO_CRC) {
			c0 = ZMUL(zrow, c0);
			c1 = ZMUL(zrow, c1);
			c2 = ZMUL(zrow, c2);
			c3 = ZMUL(zrow, c3);
		}
		if (slot >= pad) {
			const uint64_t vi = slot - pad;
			const uint4 v = __ldg(body + v0 + vi);
			if (DO_CRC) { c0 ^= v.x; c1 ^= v.y; c2 ^= v.z; c3 ^= v.w; }
			if (DO_ADLER) {
				uint32_t s = __dp4a(v.x, 0x01010101u, 0u);
				s = __dp4a(v.y, 0x01010101u, s);
				s = __dp4a(v.z, 0x01010101u, s);
				s = __dp4a(v.w, 0x01010101u, s);
				uint32_t w = __dp4a(v.x, 0x03020100u, 0u);
				w = __dp4a(v.y, 0x07060504u, w);
				w = __dp4a(v.z, 0x0b0a0908u, w);
				w = __dp4a(v.w, 0x0f0e0d0cu, w);
				sa += s;
				sw += w;
				/* every byte d_k at span offset o+k weighs (span_bytes - o - k) */
				sb += (unsigned long long) (span_bytes - vi * 16) * s;
			}
		}
	}

	uint32_t crc = 0;
	if (DO_CRC) {
		row[tid * 4 + 0] = c0;
		row[tid * 4 + 1] = c1;
		row[tid * 4 + 2] = c2;
		row[tid * 4 + 3] = c3;
		__syncthreads();
		if (tid < 32) {
			/* 1024 words -> 32 columns of a 128 byte row */
			uint32_t r = 0;
			for (int j = 0; j < 32; j++) {
				r = ZMUL(zw0, r);
				r ^= row[j * 32 + tid];
			}
			__syncwarp();
			row[tid] = r;
			__syncwarp();
			if (tid == 0) {
				uint32_t f = 0;
				for (int l = 0; l < 32; l++) {
					f ^= row[l];
					f = ZMUL(z4, f);
				}
				/* shift to the end of the whole body */
				const uint64_t after = (nvec - v1) * 16;
				crc = gf2_mulmod(f, gf2_xpow8(tables->xpow, after));
			}
		}
	}

	if (DO_ADLER) {
		/* reduce modulo 65521 early so the block sums stay small */
		unsigned long long b = (sb - sw) % ADLER_MOD;
		unsigned long long a = sa % ADLER_MOD;
		for (int o = 16; o; o >>= 1) {
			a += __shfl_down_sync(JDB_FULL_MASK, a, o);
			b += __shfl_down_sync(JDB_FULL_MASK, b, o);
		}
		if ((tid & 31) == 0) { red[0][tid >> 5] = a; red[1][tid >> 5] = b; }
		__syncthreads();
		if (tid == 0) {
			a = 0; b = 0;
			for (int w = 0; w < CK_THREADS / 32; w++) { a += red[0][w]; b += red[1][w]; }
			a %= ADLER_MOD;
			const uint64_t after = (nvec - v1) * 16;
			b = (b + (after % ADLER_MOD) * a) % ADLER_MOD;
			out[g].a = (uint32_t) a;
			out[g].b = (uint32_t) b;
		}
	}
	if (tid == 0) out[g].crc = crc;
}

/* lookups into the bank-replicated row table: L already points at the lane's bank */
/* byte k of r selects the entry: PRMT isolates the byte (ALU pipe), the scaled add
 * to the lane's bank goes to the FMA pipe as an IMAD -- the shift/mask/or form kept
 * the ALU pipe busy 3 ops per lookup and capped the kernel at 4.8 TB/s */
#ifdef JDB_SIMT_EMU
#define ZIDX(r, k, lb) (((((r) >> (8 * (k))) & 0xffu) << 7) + (lb))
#else
static __device__ __forceinline__ uint32_t zidx_(uint32_t r, uint32_t sel, uint32_t lb)
{
	uint32_t b, a;
	asm("prmt.b32 %0, %1, 0, %2;" : "=r"(b) : "r"(r), "r"(sel));
	asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(a) : "r"(b), "r"(lb));
	return a;
}
#define ZIDX(r, k, lb) zidx_((r), 0x4440u + (k), (lb))
#endif
/* lb = byte offset of the lane's bank inside the replicated table */
#define ZLD(R, off) (*(const uint32_t*) ((const unsigned char*) (R) + (off)))
#define ZREP(R, lb, r) (ZLD(R, ZIDX(r, 0, lb)) ^ ZLD(R, 32768u + ZIDX(r, 1, lb)) ^ \
                        ZLD(R, 65536u + ZIDX(r, 2, lb)) ^ ZLD(R, 98304u + ZIDX(r, 3, lb)))

template <bool DO_CRC, bool DO_ADLER>
__global__ void __launch_bounds__(CKW_THREADS, 1)
ck_wide_kernel(const uint4* __restrict__ body, uint64_t nvec, uint32_t nctas,
               const CkTables* __restrict__ tables, CkPartial* __restrict__ out)
{
	JDB_DYN_SMEM(dyn);
	uint32_t* rep = (uint32_t*) dyn;                                   /* [4][256][32] */
	uint32_t* row = rep + CKW_REP_WORDS;                               /* [4096]       */
	uint32_t (*zrow)[256] = (uint32_t (*)[256]) (row + CKW_THREADS * 4);
	uint32_t (*z4)[256]   = zrow + 4;
	uint32_t (*zw0)[256]  = z4 + 4;
	__shared__ unsigned long long red[2][CKW_THREADS / 32];

	const uint32_t tid = threadIdx.x;
	const uint32_t g = blockIdx.x;

	if (DO_CRC) {
		const uint32_t* src = &tables->z[ZT_WROW][0][0];
		for (uint32_t i = tid; i < CKW_REP_WORDS; i += CKW_THREADS) rep[i] = src[i >> 5];
		(&zrow[0][0])[tid] = (&tables->z[ZT_ROW][0][0])[tid];
		(&z4[0][0])[tid]   = (&tables->z[ZT_4][0][0])[tid];
		(&zw0[0][0])[tid]  = (&tables->z[ZT_W0][0][0])[tid];
	}
	__syncthreads();

	const uint64_t v0 = nvec * g / nctas;
	const uint64_t v1 = nvec * (g + 1) / nctas;
	const uint64_t span = v1 - v0;
	const uint64_t rows = (span + CKW_THREADS - 1) / CKW_THREADS;
	const uint64_t pad = rows * CKW_THREADS - span;       /* right aligned, see above */
	const uint32_t lb = (tid & 31u) * 4;

	uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
	unsigned long long sa = 0, sb = 0, sw = 0;

	/* Adler-32: a byte at span offset o weighs (span_bytes - o).  The weight of a
	 * thread's vector drops by one row (16384) per iteration, so with sa the running
	 * byte sum, sb += sa before each row leaves sb = sum_j s_j * (rows after j) and
	 * the weighted sum is w_last * sa + 16384 * sb - (in-vector offsets, sw) */
#define CKW_ADLER(v) do { \
		uint32_t s_ = __dp4a((v).x, 0x01010101u, 0u); \
		s_ = __dp4a((v).y, 0x01010101u, s_); \
		s_ = __dp4a((v).z, 0x01010101u, s_); \
		s_ = __dp4a((v).w, 0x01010101u, s_); \
		uint32_t w_ = __dp4a((v).x, 0x03020100u, 0u); \
		w_ = __dp4a((v).y, 0x07060504u, w_); \
		w_ = __dp4a((v).z, 0x0b0a0908u, w_); \
		w_ = __dp4a((v).w, 0x0f0e0d0cu, w_); \
		sb += sa; sa += s_; sw += w_; \
	} while (0)

	if (rows) {
		/* first row: the only one with empty slots */
		if (tid >= pad) {
			const uint4 v = __ldg(body + v0 + (tid - pad));
			if (DO_CRC) { c0 = v.x; c1 = v.y; c2 = v.z; c3 = v.w; }
			if (DO_ADLER) CKW_ADLER(v);
		}
		const uint4* p = body + v0 + (CKW_THREADS - pad) + tid;
		uint64_t j = 1;
		for (; j + 4 <= rows; j += 4) {
			const uint4 va = __ldg(p);
			const uint4 vb = __ldg(p + CKW_THREADS);
			const uint4 vc = __ldg(p + 2 * CKW_THREADS);
			const uint4 vd = __ldg(p + 3 * CKW_THREADS);
			p += 4 * CKW_THREADS;
			if (DO_CRC) {
				c0 = ZREP(rep, lb, c0) ^ va.x; c1 = ZREP(rep, lb, c1) ^ va.y; c2 = ZREP(rep, lb, c2) ^ va.z; c3 = ZREP(rep, lb, c3) ^ va.w;
				c0 = ZREP(rep, lb, c0) ^ vb.x; c1 = ZREP(rep, lb, c1) ^ vb.y; c2 = ZREP(rep, lb, c2) ^ vb.z; c3 = ZREP(rep, lb, c3) ^ vb.w;
				c0 = ZREP(rep, lb, c0) ^ vc.x; c1 = ZREP(rep, lb, c1) ^ vc.y; c2 = ZREP(rep, lb, c2) ^ vc.z; c3 = ZREP(rep, lb, c3) ^ vc.w;
				c0 = ZREP(rep, lb, c0) ^ vd.x; c1 = ZREP(rep, lb, c1) ^ vd.y; c2 = ZREP(rep, lb, c2) ^ vd.z; c3 = ZREP(rep, lb, c3) ^ vd.w;
			}
			if (DO_ADLER) {
				CKW_ADLER(va);
				CKW_ADLER(vb);
				CKW_ADLER(vc);
				CKW_ADLER(vd);
			}
		}
		for (; j < rows; j++) {
			const uint4 v = __ldg(p);
			p += CKW_THREADS;
			if (DO_CRC) { c0 = ZREP(rep, lb, c0) ^ v.x; c1 = ZREP(rep, lb, c1) ^ v.y; c2 = ZREP(rep, lb, c2) ^ v.z; c3 = ZREP(rep, lb, c3) ^ v.w; }
			if (DO_ADLER) CKW_ADLER(v);
		}
	}
#undef CKW_ADLER

	uint32_t crc = 0;
	if (DO_CRC) {
		row[tid * 4 + 0] = c0;
		row[tid * 4 + 1] = c1;
		row[tid * 4 + 2] = c2;
		row[tid * 4 + 3] = c3;
		__syncthreads();
		/* 4096 words -> 1024 columns of a 4 KiB row */
		uint32_t r = 0;
		for (int k = 0; k < 4; k++) {
			r = ZMUL(zrow, r);
			r ^= row[k * CKW_THREADS + tid];
		}
		__syncthreads();
		row[tid] = r;
		__syncthreads();
		if (tid < 32) {
			/* 1024 words -> 32 columns of a 128 byte row */
			r = 0;
			for (int k = 0; k < 32; k++) {
				r = ZMUL(zw0, r);
				r ^= row[k * 32 + tid];
			}
			__syncwarp();
			row[tid] = r;
			__syncwarp();
			if (tid == 0) {
				uint32_t f = 0;
				for (int l = 0; l < 32; l++) {
					f ^= row[l];
					f = ZMUL(z4, f);
				}
				const uint64_t after = (nvec - v1) * 16;
				crc = gf2_mulmod(f, gf2_xpow8(tables->xpow, after));
			}
		}
	}

	if (DO_ADLER) {
		const unsigned long long wlast = (unsigned long long) (CKW_THREADS - tid) * 16;
		unsigned long long b = (wlast * sa + (sb % ADLER_MOD) * CKW_ROW_BYTES - sw) % ADLER_MOD;
		unsigned long long a = sa % ADLER_MOD;
		for (int o = 16; o; o >>= 1) {
			a += __shfl_down_sync(JDB_FULL_MASK, a, o);
			b += __shfl_down_sync(JDB_FULL_MASK, b, o);
		}
		if ((tid & 31) == 0) { red[0][tid >> 5] = a; red[1][tid >> 5] = b; }
		__syncthreads();
		if (tid == 0) {
			a = 0; b = 0;
			for (int w = 0; w < CKW_THREADS / 32; w++) { a += red[0][w]; b += red[1][w]; }
			a %= ADLER_MOD;
			const uint64_t after = (nvec - v1) * 16;
			b = (b + (after % ADLER_MOD) * a) % ADLER_MOD;
			out[g].a = (uint32_t) a;
			out[g].b = (uint32_t) b;
		}
	}
	if (tid == 0) out[g].crc = crc;
}

/*
 * Ordered combine: head bytes (before 16-byte alignment), the xor / sum of the
 * CTA partials, tail bytes.  One warp; the serial parts touch < 32 bytes.
 */
__global__ void __launch_bounds__(32)
ck_combine_kernel(const uint8_t* data, uint32_t nhead, uint64_t nbody, uint32_t ntail,
                  const CkPartial* __restrict__ parts, uint32_t nparts,
                  const CkTables* __restrict__ tables, int which,
                  uint32_t* crc_io, uint32_t* adler_io)
{
	const uint32_t lane = threadIdx.x;
	uint32_t fx = 0;
	unsigned long long sa = 0, sb = 0;
	for (uint32_t i = lane; i < nparts; i += 32) {
		fx ^= parts[i].crc;
		sa += parts[i].a;
		sb += parts[i].b;
	}
	for (int o = 16; o; o >>= 1) {
		fx ^= __shfl_down_sync(JDB_FULL_MASK, fx, o);
		sa += __shfl_down_sync(JDB_FULL_MASK, sa, o);
		sb += __shfl_down_sync(JDB_FULL_MASK, sb, o);
	}
	if (lane != 0) return;

	const uint32_t (*t0)[256] = tables->z[ZT_4];
	const uint8_t* tail = data + nhead + nbody;

	if (which & JDB_CK_CRC32) {
		uint32_t r = *crc_io;
		/* byte step: advance by one byte = z4-table 3 on the low byte */
		for (uint32_t i = 0; i < nhead; i++) r = (r >> 8) ^ t0[3][(r ^ data[i]) & 0xffu];
		if (nbody) r = gf2_mulmod(r, gf2_xpow8(tables->xpow, nbody)) ^ fx;
		for (uint32_t i = 0; i < ntail; i++) r = (r >> 8) ^ t0[3][(r ^ tail[i]) & 0xffu];
		*crc_io = r;
	}
	if (which & JDB_CK_ADLER32) {
		unsigned long long a = *adler_io & 0xffffu;
		unsigned long long b = (*adler_io >> 16) & 0xffffu;
		for (uint32_t i = 0; i < nhead; i++) { a += data[i]; b += a; }
		a %= ADLER_MOD; b %= ADLER_MOD;
		if (nbody) {
			b = (b + (nbody % ADLER_MOD) * a + sb) % ADLER_MOD;
			a = (a + sa) % ADLER_MOD;
		}
		for (uint32_t i = 0; i < ntail; i++) { a += tail[i]; b += a; }
		a %= ADLER_MOD; b %= ADLER_MOD;
		*adler_io = (uint32_t) ((b << 16) | a);
	}
}

/* ---- launcher ---------------------------------------------------------- */

/* smallest body that takes the wide kernel (its 128 KiB table fill has to pay) */
static uint64_t ckw_min_bytes(void)
{
	/* read once per process */
	static const uint64_t v = [] {
		const char* e = getenv("JDB200_CK_WIDE_MIN_KIB");
		return e && atoll(e) > 0 ? (uint64_t) atoll(e) << 10 : (uint64_t) 16 << 20;
	}();
	return v;
}

extern "C" size_t jdb_checksum_workspace_bytes(void)
{
	return sizeof(CkPartial) * CK_MAX_CTAS;
}

extern "C" int jdb_checksum(const uint8_t* data, size_t n, int which,
                            uint32_t* crc_io, uint32_t* adler_io,
                            void* work, jdb_stream s)
{
	if (n == 0 || (which & (JDB_CK_CRC32 | JDB_CK_ADLER32)) == 0) return JDB_OK;
	const CkTables* tables = device_tables(s);
	if (!tables) return JDB_ENOMEM;

	uint32_t nhead = (uint32_t) ((16 - ((uintptr_t) data & 15)) & 15);
	if (nhead > n) nhead = (uint32_t) n;
	uint64_t nbody = (n - nhead) & ~(uint64_t) 15;
	uint32_t ntail = (uint32_t) (n - nhead - nbody);
	uint64_t nvec = nbody / 16;

	uint32_t nctas = 0;
	if (nvec && nbody >= ckw_min_bytes()) {
		/* one CTA per SM, at least 128 KiB each */
		JDB_CONFIGURE_SMEM((ck_wide_kernel<true, false>), CKW_SMEM);
		JDB_CONFIGURE_SMEM((ck_wide_kernel<true, true>), CKW_SMEM);
		JDB_CONFIGURE_SMEM((ck_wide_kernel<false, true>), CKW_SMEM);
		uint64_t want = nbody >> 17;
		uint64_t cap = (uint64_t) jdb_rt_sm_count();
		nctas = (uint32_t) (want < cap ? want : cap);
		if (nctas == 0) nctas = 1;
		const uint4* body = (const uint4*) (data + nhead);
		CkPartial* parts = (CkPartial*) work;
		if ((which & JDB_CK_CRC32) == 0)
			JDB_LAUNCH((ck_wide_kernel<false, true>), dim3(nctas), dim3(CKW_THREADS), CKW_SMEM, s, body, nvec, nctas, tables, parts);
		else if (which & JDB_CK_ADLER32)
			JDB_LAUNCH((ck_wide_kernel<true, true>), dim3(nctas), dim3(CKW_THREADS), CKW_SMEM, s, body, nvec, nctas, tables, parts);
		else
			JDB_LAUNCH((ck_wide_kernel<true, false>), dim3(nctas), dim3(CKW_THREADS), CKW_SMEM, s, body, nvec, nctas, tables, parts);
		int r = jdb_rt_check_launch("ck_wide_kernel");
		if (r != JDB_OK) return r;
	} else if (nvec) {
		/* at least 64 KiB per CTA, at most 4 CTAs per SM */
		uint64_t want = (nbody + 65535) / 65536;
		uint64_t cap = (uint64_t) jdb_rt_sm_count() * 4;
		if (cap > CK_MAX_CTAS) cap = CK_MAX_CTAS;
		nctas = (uint32_t) (want < cap ? want : cap);
		const uint4* body = (const uint4*) (data + nhead);
		CkPartial* parts = (CkPartial*) work;
		const bool c = (which & JDB_CK_CRC32) != 0, a = (which & JDB_CK_ADLER32) != 0;
		if (c && a)
			JDB_LAUNCH((ck_partial_kernel<true, true>), dim3(nctas), dim3(CK_THREADS), 0, s, body, nvec, nctas, tables, parts);
		else if (c)
			JDB_LAUNCH((ck_partial_kernel<true, false>), dim3(nctas), dim3(CK_THREADS), 0, s, body, nvec, nctas, tables, parts);
		else
			JDB_LAUNCH((ck_partial_kernel<false, true>), dim3(nctas), dim3(CK_THREADS), 0, s, body, nvec, nctas, tables, parts);
		int r = jdb_rt_check_launch("ck_partial_kernel");
		if (r != JDB_OK) return r;
	}
	JDB_LAUNCH(ck_combine_kernel, dim3(1), dim3(32), 0, s, data, nhead, nbody, ntail,
	           (const CkPartial*) work, nctas, tables, which, crc_io, adler_io);
	return jdb_rt_check_launch("ck_combine_kernel");
}
