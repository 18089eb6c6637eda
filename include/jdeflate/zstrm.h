/*
 * jdeflate/zstrm.h -- gzip / zlib / raw container layer with callback I/O,
 * B200-native implementation.
 *
 * Drop-in for the reference header (reference jdeflate/zstrm.h:37-223): same
 * flag, error and state enumerators, same 56 byte public TZStrm block, same
 * thirteen entry points.  Differences in behaviour are the documented defect
 * fixes of DESIGN.md (valid zlib FCHECK, RFC-correct Adler-32, exact input
 * accounting, zstrm_crc32combine actually defined).
 */
#ifndef JDB200_JDEFLATE_ZSTRM_H
#define JDB200_JDEFLATE_ZSTRM_H

#include <ctoolbox/ctoolbox.h>
#include <ctoolbox/memory.h>
#include "deflator.h"
#include "inflator.h"

#ifdef __cplusplus
extern "C" {
#endif

/* direction, one of (reference jdeflate/zstrm.h:38-41) */
typedef enum {
	ZSTRM_INFLATE = 0x00010000,
	ZSTRM_DEFLATE = 0x00020000
} eZSTRMMode;

/* container; exactly one for deflate, any subset (0 = all) for inflate */
typedef enum {
	ZSTRM_DFLT = 0x00100000,
	ZSTRM_ZLIB = 0x00200000,
	ZSTRM_GZIP = 0x00400000
} eZSTRMType;

typedef enum {
	ZSTRM_DOCRC   = 0x01000000,   /* compute even if the container has none */
	ZSTRM_DOADLER = 0x02000000,
	ZSTRM_NOCRC   = 0x04000000,   /* inflate only: skip verification        */
	ZSTRM_NOADLER = 0x08000000
} eZSTRMFlags;

typedef enum {
	ZSTRM_OK            =  0,
	ZSTRM_EIOERROR      =  1,
	ZSTRM_EOOM          =  2,
	ZSTRM_EBADDATA      =  3,
	ZSTRM_ECHECKSUM     =  4,
	ZSTRM_EFORMAT       =  5,
	ZSTRM_EMISSINGDICT  =  6,
	ZSTRM_ESRCEXHSTD    =  7,
	ZSTRM_ETGTEXHSTD    =  8,
	ZSTRM_EDEFLATE      =  9,
	ZSTRM_EBADDICT      = 10,
	ZSTRM_ELIMIT        = 11,
	ZSTRM_EINCORRECTUSE = 12
} eZSTRMError;

typedef enum {
	ZSTRM_NOTSET   = 0,
	ZSTRM_READY    = 1,
	ZSTRM_NEEDDICT = 2,
	ZSTRM_NORMAL   = 3,
	ZSTRM_END      = 4
} eZSTRMState;

/*
 * I/O callbacks: return the number of bytes read / written; a source callback
 * returns 0 at end of input; negative means failure
 * (reference jdeflate/zstrm.h:92-101).  A target callback has to take every
 * byte it is offered.
 */
typedef intxx (*TZStrmIFn)(      uint8* buffer, uintxx size, void* user);
typedef intxx (*TZStrmOFn)(const uint8* buffer, uintxx size, void* user);

/* Public state, handed out as `const TZStrm*`
 * (reference jdeflate/zstrm.h:105-132). */
struct TZStrm {
	uint32 state;      /* eZSTRMState */
	uint32 error;      /* eZSTRMError */
	uint32 flags;
	uint32 smode;      /* eZSTRMMode  */
	uint32 stype;      /* eZSTRMType  */
	 int32 level;

	uintxx total;      /* uncompressed bytes read or written so far */

	uint32 dictid;     /* Adler-32 of the preset dictionary */
	uint32 dict;

	uint32 crc;
	uint32 adler;

	uintxx usedinput;  /* compressed bytes consumed (inflate mode) */
};

typedef struct TZStrm TZStrm;

JDEFLATE_API
const TZStrm* zstrm_create(uintxx flags, intxx level, const TAllocator*);

JDEFLATE_API
void zstrm_destroy(const TZStrm*);

/* inflate input: either one memory buffer ... */
JDEFLATE_API
void zstrm_setsource(const TZStrm*, const uint8* source, uintxx size);

/* ... or a pull callback */
JDEFLATE_API
void zstrm_setsourcefn(const TZStrm*, TZStrmIFn fn, void* user);

/* deflate output: push callback */
JDEFLATE_API
void zstrm_settargetfn(const TZStrm*, TZStrmOFn fn, void* user);

JDEFLATE_API
void zstrm_setdctnr(const TZStrm*, const uint8* dict, uintxx size);

/* decompress up to n (< 2^31) bytes into target; returns bytes produced */
JDEFLATE_API
uintxx zstrm_inflate(const TZStrm*, void* target, uintxx n);

/* compress n (< 2^31) bytes; returns n unless the target callback failed */
JDEFLATE_API
uintxx zstrm_deflate(const TZStrm*, const void* source, uintxx n);

/* push out everything buffered; final != 0 also writes the trailer */
JDEFLATE_API
void zstrm_flush(const TZStrm*, uint32 final);

JDEFLATE_API
void zstrm_reset(const TZStrm*);


/* ---- stand-alone checksum helpers (reference jdeflate/zstrm.h:203-223) -- */

/* crc of A||B from the finalised crc(A), crc(B) and len(B) */
JDEFLATE_API
uint32 zstrm_crc32combine(uint32 crc1, uint32 crc2, uintxx size2);

/* takes and returns the NON-finalised register (start 0xFFFFFFFF, xor at end) */
JDEFLATE_API
uint32 zstrm_crc32update(uint32 chcksm, const uint8* source, uintxx size);

/* takes and returns the plain Adler-32 value (start 1) */
JDEFLATE_API
uint32 zstrm_adler32update(uint32 chcksm, const uint8* source, uintxx size);

/* the name the reference object actually exports for the combine
 * (reference src/zstrm.c:1427-1443); kept so binaries linked against it load */
JDEFLATE_API
uint32 crc32_ncombine(uint32 crc1, uint32 crc2, uint32 size2);

#ifdef __cplusplus
}
#endif

#endif
