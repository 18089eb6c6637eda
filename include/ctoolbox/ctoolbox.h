/*
 * ctoolbox/ctoolbox.h -- minimal stand-in for the (un-vendored) ctoolbox
 * dependency of jdeflate.
 *
 * The reference pulls ctoolbox through a meson wrap at `revision = master`
 * (reference subprojects/ctoolbox.wrap:1-3), so no pinned copy exists.  This
 * header supplies exactly the names the public jdeflate headers and the
 * reference sources use (list in SURVEY.md section 8c): fixed width integer
 * typedefs, pointer-width `uintxx/intxx`, and the CTB_* helper macros.
 * Nothing DEFLATE specific lives here.
 */
#ifndef JDB200_CTOOLBOX_CTOOLBOX_H
#define JDB200_CTOOLBOX_CTOOLBOX_H

#include <stddef.h>
#include <stdint.h>
#include <assert.h>

#if !defined(__cplusplus)
	#include <stdbool.h>
#endif

typedef uint8_t  uint8;
typedef uint16_t uint16;
typedef uint32_t uint32;
typedef uint64_t uint64;
typedef int8_t   int8;
typedef int16_t  int16;
typedef int32_t  int32;
typedef int64_t  int64;

/* pointer sized integers */
typedef uintptr_t uintxx;
typedef intptr_t  intxx;

#if UINTPTR_MAX > 0xffffffffu
	#define CTB_ENV64 1
#endif

/* this build only targets little endian hosts (x86-64 / aarch64 + B200) */
#define CTB_IS_LITTLEENDIAN 1
#define CTB_IS_BIGENDIAN    0
#define CTB_FASTUNALIGNED   1

#if defined(__GNUC__)
	#define CTB_INLINE      static inline __attribute__((unused))
	#define CTB_FORCEINLINE static inline __attribute__((always_inline, unused))
	#define CTB_EXPECT1(C)  __builtin_expect(!!(C), 1)
	#define CTB_EXPECT0(C)  __builtin_expect(!!(C), 0)
#else
	#define CTB_INLINE      static inline
	#define CTB_FORCEINLINE static inline
	#define CTB_EXPECT1(C)  (C)
	#define CTB_EXPECT0(C)  (C)
#endif

#if defined(NDEBUG)
	#define CTB_ASSERT(C) ((void) 0)
#else
	#define CTB_ASSERT(C) assert(C)
#endif

#define CTB_CONSTCAST(P) ((void*) (uintptr_t) (P))

/* byte order helpers: identity on little endian, swap on big endian */
#define CTB_SWAP32(X) __builtin_bswap32(X)
#define CTB_SWAP64(X) __builtin_bswap64(X)
#define CTB_SWAP32ONLE(X) CTB_SWAP32(X)
#define CTB_SWAP32ONBE(X) (X)
#define CTB_SWAP64ONLE(X) CTB_SWAP64(X)
#define CTB_SWAP64ONBE(X) (X)

#endif
