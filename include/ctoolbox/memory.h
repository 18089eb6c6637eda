/*
 * ctoolbox/memory.h -- allocator interface stand-in.
 *
 * The reference only ever calls `request(size, user)`, `dispose(ptr, size,
 * user)` and reads `user` (reference src/deflator.c:283,391,526,544;
 * src/inflator.c:174,235,258; src/zstrm.c:123,193).  The member ORDER of the
 * real TAllocator cannot be recovered from the reference tree, so the ABI for
 * custom allocators is unpinned; passing NULL selects
 * the default allocator and is what every supported configuration does.
 */
#ifndef JDB200_CTOOLBOX_MEMORY_H
#define JDB200_CTOOLBOX_MEMORY_H

#include <string.h>
#include <stdlib.h>
#include "ctoolbox.h"

struct TAllocator {
	void* (*request)(uintxx size, void* user);
	void  (*dispose)(void* memory, uintxx size, void* user);
	void* user;
};

typedef struct TAllocator TAllocator;

static inline __attribute__((unused)) void*
ctb_default_request_(uintxx size, void* user)
{
	(void) user;
	return malloc((size_t) size);
}

static inline __attribute__((unused)) void
ctb_default_dispose_(void* memory, uintxx size, void* user)
{
	(void) size;
	(void) user;
	free(memory);
}

static inline __attribute__((unused)) const TAllocator*
ctb_getdefaultallocator(void)
{
	static const TAllocator dflt = {
		ctb_default_request_, ctb_default_dispose_, NULL
	};
	return &dflt;
}

#define ctb_memcpy(D, S, N) memcpy((D), (S), (size_t) (N))
#define ctb_memset(D, V, N) memset((D), (V), (size_t) (N))

#endif
