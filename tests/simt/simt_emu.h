/*
 * simt_emu.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A tiny single-process SIMT emulator that lets the .cu kernel sources of
 * jdeflate_b200/csrc/device be compiled with g++ (-DJDB_SIMT_EMU) and executed
 * on the CPU under ASan/UBSan/gdb in the build container, which has no GPU.
 * It exists to debug kernel LOGIC (indexing, warp collectives, barriers)
 * before spending GPU-box minutes.  It is never part of libjdeflate.so, is not
 * importable from the jdeflate_b200 package, and is not a fallback: the
 * product library refuses to work without a CUDA device.
 *
 * Model: CTAs run one after another; the threads of a CTA are cooperative
 * fibers switched round-robin at every barrier / warp collective.  Warp
 * collectives rendezvous all lanes named in the mask.  Atomics are plain
 * read-modify-write (there is no concurrency).
 */
#ifndef JDB_SIMT_EMU_H
#define JDB_SIMT_EMU_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <functional>
#include <vector>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) alignas(n)

struct uint3 { unsigned x, y, z; };
struct dim3 {
	unsigned x, y, z;
	dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct uchar4 { unsigned char x, y, z, w; };
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }

namespace simt {

struct Fiber {
	void*  sp;
	char*  stack;
	bool   done;
	uint3  tid;
	int    lane, warp;
	uint64_t coll[32];      /* snapshot of the last warp collective */
	unsigned coll_mask;
	bool   coll_ready;
};

struct Warp {
	uint64_t vals[32];
	unsigned arrived;
};

struct Cta {
	std::vector<Fiber> f;
	int      n;
	int      alive;
	int      bar_arrived;
	unsigned bar_gen;
	int      or_flag;              /* __syncthreads_or */
	int      nb_arrived[16];       /* named barriers (bar.sync id, count) */
	unsigned nb_gen[16];
	Warp     warps[32];
	uint3    bid;
	dim3     bdim, gdim;
	std::function<void()> body;
	void*    main_sp;
	int      cur;
};

extern Cta* g_cta;
extern unsigned char* g_dsmem;

extern "C" void simt_switch(void** save_sp, void* new_sp);

static inline Fiber& self() { return g_cta->f[g_cta->cur]; }

/* hand the CPU to the next unfinished fiber (or back to main when none) */
void yield();
void run_grid(dim3 grid, dim3 block, size_t smem, std::function<void()> body);
void collective(unsigned mask, uint64_t v);

} /* namespace simt */

#define threadIdx (simt::self().tid)
#define blockIdx  (simt::g_cta->bid)
#define blockDim  (simt::g_cta->bdim)
#define gridDim   (simt::g_cta->gdim)
#define warpSize  32

#define JDB_DYN_SMEM(name) unsigned char* name = simt::g_dsmem

static inline void __syncthreads()
{
	simt::Cta* c = simt::g_cta;
	unsigned gen = c->bar_gen;
	if (++c->bar_arrived >= c->alive) {
		c->bar_arrived = 0;
		c->bar_gen++;
		return;
	}
	while (c->bar_gen == gen)
		simt::yield();
}

static inline int __syncthreads_or(int pred)
{
	simt::Cta* c = simt::g_cta;
	if (pred) c->or_flag = 1;
	__syncthreads();
	const int r = c->or_flag;
	__syncthreads();
	if (simt::g_cta->cur == 0 || simt::self().tid.x + simt::self().tid.y + simt::self().tid.z == 0) c->or_flag = 0;
	__syncthreads();
	return r;
}

/* bar.sync id, count: `count` threads of the CTA meet at named barrier `id` */
static inline void simt_named_barrier(unsigned id, int count)
{
	simt::Cta* c = simt::g_cta;
	unsigned gen = c->nb_gen[id];
	if (++c->nb_arrived[id] >= count) {
		c->nb_arrived[id] = 0;
		c->nb_gen[id]++;
		return;
	}
	while (c->nb_gen[id] == gen)
		simt::yield();
}

static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::collective(mask, 0); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) {}
static inline long long clock64() { return 0; }

static inline unsigned __ballot_sync(unsigned mask, int pred)
{
	simt::collective(mask, pred ? 1 : 0);
	simt::Fiber& f = simt::self();
	unsigned r = 0;
	for (int i = 0; i < 32; i++)
		if ((mask >> i) & 1u) r |= (unsigned) (f.coll[i] & 1) << i;
	return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }

template <typename T> static inline uint64_t simt_pack(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> static inline T simt_unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
	simt::collective(mask, simt_pack(v));
	simt::Fiber& f = simt::self();
	int base = f.lane & ~(width - 1);
	int s = base + (src & (width - 1));
	return simt_unpack<T>(f.coll[s]);
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{
	simt::collective(mask, simt_pack(v));
	simt::Fiber& f = simt::self();
	int base = f.lane & ~(width - 1);
	int s = f.lane - (int) d;
	if (s < base) s = f.lane;
	return simt_unpack<T>(f.coll[s]);
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32)
{
	simt::collective(mask, simt_pack(v));
	simt::Fiber& f = simt::self();
	int base = f.lane & ~(width - 1);
	int s = f.lane + (int) d;
	if (s >= base + width) s = f.lane;
	return simt_unpack<T>(f.coll[s]);
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32)
{
	simt::collective(mask, simt_pack(v));
	simt::Fiber& f = simt::self();
	int s = f.lane ^ x;
	(void) width;
	return simt_unpack<T>(f.coll[s]);
}
template <typename T> static inline unsigned __match_any_sync(unsigned mask, T v)
{
	simt::collective(mask, simt_pack(v));
	simt::Fiber& f = simt::self();
	unsigned r = 0;
	uint64_t mine = f.coll[f.lane];
	for (int i = 0; i < 32; i++)
		if (((mask >> i) & 1u) && f.coll[i] == mine) r |= 1u << i;
	return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v)
{
	simt::collective(mask, v);
	simt::Fiber& f = simt::self();
	unsigned r = 0;
	for (int i = 0; i < 32; i++) if ((mask >> i) & 1u) r += (unsigned) f.coll[i];
	return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v)
{
	simt::collective(mask, v);
	simt::Fiber& f = simt::self();
	unsigned r = 0;
	for (int i = 0; i < 32; i++) if ((mask >> i) & 1u) r = std::max(r, (unsigned) f.coll[i]);
	return r;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
	simt::collective(mask, v);
	simt::Fiber& f = simt::self();
	unsigned r = 0;
	for (int i = 0; i < 32; i++) if ((mask >> i) & 1u) r |= (unsigned) f.coll[i];
	return r;
}

/* integer intrinsics */
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned) v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long) v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
/* position of the offset-th set bit of mask at or above base (offset >= 1), 0xffffffff if none */
static inline unsigned __fns(unsigned mask, unsigned base, int offset)
{
	for (unsigned i = base; i < 32; i++)
		if ((mask >> i) & 1u) { if (--offset == 0) return i; }
	return 0xffffffffu;
}
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v)
{
	v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
	v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
	v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
	return __builtin_bswap32(v);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
	uint64_t t = ((uint64_t) b << 32) | a;
	unsigned r = 0;
	for (int i = 0; i < 4; i++) {
		unsigned sel = (s >> (4 * i)) & 7u;
		r |= (unsigned) ((t >> (8 * sel)) & 0xff) << (8 * i);
	}
	return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh)
{
	uint64_t t = ((uint64_t) hi << 32) | lo;
	return (unsigned) (t >> (sh & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh)
{
	uint64_t t = ((uint64_t) hi << 32) | lo;
	return (unsigned) ((t << (sh & 31)) >> 32);
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned) (((uint64_t) a * b) >> 32); }
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c)
{
	for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
	return c;
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcg(T* p, T v) { *p = v; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }

template <typename T, typename U> static inline T atomicAdd(T* p, U v) { T o = *p; *p = (T) (o + (T) v); return o; }
template <typename T, typename U> static inline T atomicOr(T* p, U v)  { T o = *p; *p = (T) (o | (T) v); return o; }
template <typename T, typename U> static inline T atomicAnd(T* p, U v) { T o = *p; *p = (T) (o & (T) v); return o; }
template <typename T, typename U> static inline T atomicMax(T* p, U v) { T o = *p; if ((T) v > o) *p = (T) v; return o; }
template <typename T, typename U> static inline T atomicMin(T* p, U v) { T o = *p; if ((T) v < o) *p = (T) v; return o; }
template <typename T, typename U> static inline T atomicExch(T* p, U v) { T o = *p; *p = (T) v; return o; }
template <typename T, typename U> static inline T atomicCAS(T* p, U c, U v) { T o = *p; if (o == (T) c) *p = (T) v; return o; }

#ifndef min
using std::min;
using std::max;
#endif
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

/* launch: the stream argument is ignored (everything is synchronous) */
#define JDB_LAUNCH(kernel, grid, block, smem, stream, ...) \
	simt::run_grid((grid), (block), (size_t) (smem), [=]() { kernel(__VA_ARGS__); })

#endif
