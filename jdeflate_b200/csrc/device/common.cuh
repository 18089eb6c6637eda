/* common.cuh -- shared device-side helpers for the sm_100a kernels */
#ifndef JDB_COMMON_CUH
#define JDB_COMMON_CUH

#include <stdint.h>
#include <stddef.h>

#ifdef JDB_SIMT_EMU
	/* CPU debugging build, see tests/simt/simt_emu.h (test infrastructure) */
	#include "simt_emu.h"
	#define JDB_CONFIGURE_SMEM(kernel, bytes) do { } while (0)
#else
	#include <cuda_runtime.h>
	/* every kernel launch of the library goes through here: launches are counted
	 * per kernel and, when profiling is on (jdb200_profile), bracketed by CUDA
	 * events on the launching stream */
	#define JDB_LAUNCH(kernel, grid, block, smem, stream, ...) \
		do { \
			static int jdb_prof_site_ = -1; \
			int jdb_prof_slot_ = jdb_prof_begin(#kernel, &jdb_prof_site_, (stream)); \
			kernel<<<(grid), (block), (smem), (cudaStream_t) (stream)>>>(__VA_ARGS__); \
			jdb_prof_end(jdb_prof_slot_, (stream)); \
		} while (0)
	#define JDB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
	/* opt a kernel in to its dynamic shared memory once per device (the device the calling
	 * thread is bound to, i.e. the one the launch goes to); idempotent, so a race between two
	 * threads that both find the flag clear is harmless */
	#define JDB_CONFIGURE_SMEM(kernel, bytes) \
		do { \
			static int jdb_cfg_done_[64]; \
			const int jdb_cfg_dev_ = jdb_rt_current_device(); \
			if (jdb_cfg_dev_ >= 0 && jdb_cfg_dev_ < 64 && !__atomic_load_n(&jdb_cfg_done_[jdb_cfg_dev_], __ATOMIC_ACQUIRE)) { \
				cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (bytes)); \
				__atomic_store_n(&jdb_cfg_done_[jdb_cfg_dev_], 1, __ATOMIC_RELEASE); \
			} \
		} while (0)
#endif

#include "jdb_device.h"

#define JDB_FULL_MASK 0xffffffffu

/* set by runtime.cu / runtime_emu.cpp */
extern "C" int  jdb_rt_check_launch(const char* what);
extern "C" void jdb_rt_set_error(const char* fmt, ...);
extern "C" int  jdb_prof_begin(const char* kernel, int* site, jdb_stream s);
extern "C" void jdb_prof_end(int slot, jdb_stream s);

static __device__ __forceinline__ unsigned jdb_lane() { return threadIdx.x & 31u; }
static __device__ __forceinline__ unsigned jdb_warp() { return threadIdx.x >> 5; }

/* A pointer into shared memory whose address the compiler takes from a register instead of
 * rebuilding it where it is used: sm_100 shared addresses carry the CTA's rank in its cluster
 * (S2R SR_CgaCtaId), and ptxas is happy to re-read that special register inside a hot loop. */
template <typename T> static __device__ __forceinline__ T* jdb_pin_shared(T* p)
{
#ifndef JDB_SIMT_EMU
	uint32_t a = (uint32_t) __cvta_generic_to_shared((const void*) p);
	asm volatile("" : "+r"(a));
	return (T*) __cvta_shared_to_generic((size_t) a);
#else
	return p;
#endif
}

/* unaligned little-endian 32 bit read assembled from two aligned words */
static __device__ __forceinline__ uint32_t jdb_ld32u(const uint8_t* base, uint32_t off)
{
	const uint32_t* w = (const uint32_t*) (base + (off & ~3u));
	return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
}

#endif
