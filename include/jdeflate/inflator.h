/*
 * jdeflate/inflator.h -- raw DEFLATE decoder, B200-native implementation.
 *
 * Drop-in for the reference header (reference jdeflate/inflator.h:47-189):
 * same enumerators, same 72 byte public cursor block, same inline accessors.
 * Host or device pointers may be given to setsrc/settgt.
 *
 *     inflator_setsrc(s, src, n);
 *     do {
 *         inflator_settgt(s, out, cap);
 *         r = inflator_inflate(s, last);
 *         consume(out, inflator_tgtend(s));
 *     } while (r == INFLT_TGTEXHSTD);
 */
#ifndef JDB200_JDEFLATE_INFLATOR_H
#define JDB200_JDEFLATE_INFLATOR_H

#include <ctoolbox/ctoolbox.h>
#include <ctoolbox/memory.h>
#include <jdeflate/config/config.h>

#ifdef __cplusplus
extern "C" {
#endif

/* inflator_inflate() results (reference jdeflate/inflator.h:48-53) */
typedef enum {
	INFLT_OK        = 0,   /* final block decoded                   */
	INFLT_SRCEXHSTD = 1,
	INFLT_TGTEXHSTD = 2,
	INFLT_ERROR     = 3
} eINFLTResult;

/* TInflator.error values (reference jdeflate/inflator.h:57-66) */
typedef enum {
	INFLT_EBADSTATE     = 1,
	INFLT_EBADCODE      = 2,   /* bit pattern with no code assigned        */
	INFLT_EBADTREE      = 3,   /* invalid dynamic block header             */
	INFLT_EFAROFFSET    = 4,   /* distance reaches before start of output  */
	INFLT_EBADBLOCK     = 5,   /* BTYPE 3 or stored LEN/NLEN mismatch      */
	INFLT_EINPUTEND     = 6,   /* final input ended inside the stream      */
	INFLT_EOOM          = 7,
	INFLT_EINCORRECTUSE = 8
} eINFLTError;

/* Public part of a decoder instance (reference jdeflate/inflator.h:71-89). */
struct TInflator {
	const uint32 state;
	const uint32 error;
	const uint32 flags;
	const uint32 finalinput;
	const uint32 status;     /* last value returned by inflator_inflate */

	const uint8* source;
	const uint8* sbgn;
	const uint8* send;

	uint8* target;
	uint8* tbgn;
	uint8* tend;
};

typedef struct TInflator TInflator;

JDEFLATE_API
TInflator* inflator_create(uintxx flags, const TAllocator*);

JDEFLATE_API
void inflator_destroy(TInflator*);

/* `final` non-zero: no more input will follow the current source window */
JDEFLATE_API
eINFLTResult inflator_inflate(TInflator*, uint32 final);

JDEFLATE_API
void inflator_setdctnr(TInflator*, const uint8* dict, uintxx size);

JDEFLATE_API
void inflator_reset(TInflator*);


/* ---- header-inline accessors (part of the ABI) ------------------------- */

struct TJIStateHeader_ {
	uint32 state;
	uint32 error;
};

CTB_INLINE void
inflator_setsrc(TInflator* s, const uint8* source, uintxx size)
{
	CTB_ASSERT(s && source && size);

	/* new input after `final` was announced poisons the instance
	 * (reference jdeflate/inflator.h:150-162) */
	if (CTB_EXPECT0(s->finalinput != 0)) {
		if (s->error == 0) {
			struct TJIStateHeader_* h = (struct TJIStateHeader_*) s;
			h->error = INFLT_EINCORRECTUSE;
			h->state = 0xDEADBEEF;
		}
		return;
	}
	s->sbgn   = source;
	s->source = source;
	s->send   = source + size;
}

CTB_INLINE void
inflator_settgt(TInflator* s, uint8* target, uintxx size)
{
	CTB_ASSERT(s && target && size);
	s->tbgn   = target;
	s->target = target;
	s->tend   = target + size;
}

CTB_INLINE uintxx
inflator_srcend(TInflator* s)
{
	CTB_ASSERT(s);
	return (uintxx) (s->source - s->sbgn);
}

CTB_INLINE uintxx
inflator_tgtend(TInflator* s)
{
	CTB_ASSERT(s);
	return (uintxx) (s->target - s->tbgn);
}

#ifdef __cplusplus
}
#endif

#endif
