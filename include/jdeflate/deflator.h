/*
 * jdeflate/deflator.h -- raw DEFLATE encoder, B200-native implementation.
 *
 * Drop-in for the reference header of the same name
 * (reference jdeflate/deflator.h:47-203).  Everything a caller can observe is
 * kept: enumerator names and values, the 72 byte public cursor block of
 * struct TDeflator (five u32 then six pointers), the poison value 0xDEADBEEF,
 * and the four header-inline accessors that store/read the cursor fields
 * directly.  The accessors never dereference the buffers, so both host and
 * device (cudaMalloc) pointers may be handed to setsrc/settgt; the library
 * detects which one it got.
 *
 * Typical loop (same as the reference):
 *
 *     deflator_setsrc(d, src, n);
 *     do {
 *         deflator_settgt(d, out, cap);
 *         r = deflator_deflate(d, last ? DEFLT_END : DEFLT_NOFLUSH);
 *         consume(out, deflator_tgtend(d));
 *     } while (r == DEFLT_TGTEXHSTD);
 */
#ifndef JDB200_JDEFLATE_DEFLATOR_H
#define JDB200_JDEFLATE_DEFLATOR_H

#include <ctoolbox/ctoolbox.h>
#include <ctoolbox/memory.h>
#include <jdeflate/config/config.h>

#ifdef __cplusplus
extern "C" {
#endif

/* deflator_deflate() results (reference jdeflate/deflator.h:48-53) */
typedef enum {
	DEFLT_OK        = 0,   /* stream (or flush) complete                 */
	DEFLT_SRCEXHSTD = 1,   /* all source consumed, feed more             */
	DEFLT_TGTEXHSTD = 2,   /* target window full, provide a new one      */
	DEFLT_ERROR     = 3    /* see TDeflator.error                        */
} eDEFLTResult;

/* flush argument (reference jdeflate/deflator.h:57-61) */
typedef enum {
	DEFLT_NOFLUSH = 0,
	DEFLT_END     = 1,     /* finish: last marker carries BFINAL=1       */
	DEFLT_FLUSH   = 2      /* sync flush: byte aligned empty stored block */
} eDEFLTFlush;

/* TDeflator.error values (reference jdeflate/deflator.h:65-70) */
typedef enum {
	DEFLT_EBADSTATE     = 1,
	DEFLT_EOOM          = 2,
	DEFLT_ELEVEL        = 3,
	DEFLT_EINCORRECTUSE = 4
} eDEFLTError;

/* creation flags (reference jdeflate/deflator.h:74-76) */
typedef enum {
	DEFLT_FIXEDCODES = 0x01
} eDEFLTFlags;

/*
 * Public part of an encoder instance (reference jdeflate/deflator.h:81-99).
 * `state`/`error` open the struct because the inline setters below poison
 * them through a two-u32 overlay.
 */
struct TDeflator {
	const uint32 state;
	const uint32 error;
	const uint32 flags;
	const uint32 flush;
	const uint32 status;     /* last value returned by deflator_deflate */

	const uint8* source;     /* read cursor  */
	const uint8* sbgn;
	const uint8* send;

	uint8* target;           /* write cursor */
	uint8* tbgn;
	uint8* tend;
};

typedef struct TDeflator TDeflator;

/* level 0..9; bad level or allocation failure returns NULL
 * (reference src/deflator.c:375-416) */
JDEFLATE_API
TDeflator* deflator_create(uintxx flags, intxx level, const TAllocator*);

JDEFLATE_API
void deflator_destroy(TDeflator*);

/*
 * Compress from the source window into the target window.  With DEFLT_FLUSH or
 * DEFLT_END everything buffered is written out followed by an empty stored
 * block (the Z_SYNC_FLUSH marker 00 00 FF FF); no new source may be set until
 * the call returns DEFLT_OK (reference jdeflate/deflator.h:131-141).
 */
JDEFLATE_API
eDEFLTResult deflator_deflate(TDeflator*, eDEFLTFlush flush);

/* preset dictionary (last 32 KiB are used); must precede the first deflate */
JDEFLATE_API
void deflator_setdctnr(TDeflator*, const uint8* dict, uintxx size);

JDEFLATE_API
void deflator_reset(TDeflator*);


/* ---- header-inline accessors (part of the ABI) ------------------------- */

struct TJDStateHeader_ {
	uint32 state;
	uint32 error;
};

CTB_INLINE void
deflator_setsrc(TDeflator* d, const uint8* source, uintxx size)
{
	CTB_ASSERT(d && source && size);

	/* feeding input after a flush was latched is a usage error
	 * (reference jdeflate/deflator.h:164-176) */
	if (CTB_EXPECT0(d->flush != 0)) {
		if (d->error == 0) {
			struct TJDStateHeader_* h = (struct TJDStateHeader_*) d;
			h->error = DEFLT_EINCORRECTUSE;
			h->state = 0xDEADBEEF;
		}
		return;
	}
	d->sbgn   = source;
	d->source = source;
	d->send   = source + size;
}

CTB_INLINE void
deflator_settgt(TDeflator* d, uint8* target, uintxx size)
{
	CTB_ASSERT(d && target && size);
	d->tbgn   = target;
	d->target = target;
	d->tend   = target + size;
}

/* bytes consumed since the last setsrc */
CTB_INLINE uintxx
deflator_srcend(TDeflator* d)
{
	CTB_ASSERT(d);
	return (uintxx) (d->source - d->sbgn);
}

/* bytes produced since the last settgt */
CTB_INLINE uintxx
deflator_tgtend(TDeflator* d)
{
	CTB_ASSERT(d);
	return (uintxx) (d->target - d->tbgn);
}

#ifdef __cplusplus
}
#endif

#endif
