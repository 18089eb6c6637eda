/*
 * jdeflate/b200.h -- ADDITIVE entry points of the B200-native build.
 *
 * Nothing here exists in the reference; nothing in the reference API changes.
 * These calls express what the single-stream, synchronous jdeflate API cannot
 * (last row): decoding or encoding a batch of independent
 * streams in one launch, and choosing the CUDA device.
 *
 * All pointers may be host or device (cudaMalloc) addresses; device-resident
 * arguments are used in place, host ones are staged.
 */
#ifndef JDB200_JDEFLATE_B200_H
#define JDB200_JDEFLATE_B200_H

#include <ctoolbox/ctoolbox.h>
#include <jdeflate/config/config.h>
#include <jdeflate/inflator.h>

#ifdef __cplusplus
extern "C" {
#endif

/* container of every stream in a batch */
typedef enum {
	JDB200_RAW  = 0,     /* raw DEFLATE (what inflator_inflate decodes)          */
	JDB200_ZLIB = 1      /* RFC 1950: 2 byte header, Adler-32 trailer (verified) */
} eJDB200Format;

/* one stream: byte ranges relative to the batch base pointers */
typedef struct TJDB200Item {
	uint64 srcoffset;
	uint64 tgtoffset;
	uint64 srcsize;
	uint64 tgtsize;
} TJDB200Item;

/*
 * Outcome of one stream, same meaning as a fresh TInflator fed the stream with
 * final = 1: `status` is an eINFLTResult, `error` an eINFLTError.  `zerror` is
 * 0 or a ZSTRM_E* code for container problems (ZSTRM_EBADDATA bad header /
 * missing trailer, ZSTRM_ECHECKSUM, ZSTRM_EMISSINGDICT).
 */
typedef struct TJDB200Result {
	uint32 status;
	uint32 error;
	uint32 zerror;
	uint32 checksum;     /* Adler-32 of the output for JDB200_ZLIB */
	uint64 srcused;      /* trailer included */
	uint64 tgtused;
} TJDB200Result;

/*
 * Decode `count` independent streams: one GPU warp per stream, or one thread
 * block per stream when there are no more streams than the device has SMs
 * (a few long streams: about seven times faster each).  Same results either way.
 * Returns 0 when the batch ran (look at the per stream results), non-zero for
 * a runtime failure (no device, out of memory).
 */
JDEFLATE_API
int jdb200_inflate_batch(const uint8* source, uint8* target,
                         const TJDB200Item* items, TJDB200Result* results,
                         uintxx count, eJDB200Format format);

/*
 * Compress `count` independent records, each into a complete stream of its own
 * (raw DEFLATE, or zlib with header and Adler-32 trailer), in one pass of the
 * chunk-parallel encoder per group of records -- what the reference does as a loop
 * of deflator_reset + deflator_setsrc + deflator_deflate(DEFLT_END)
 * (jdeflate/deflator.h:104-153) plus the zlib framing of zstrm (src/zstrm.c:1033-1110).
 * Item i is read from source[srcoffset, +srcsize) and written to
 * target[tgtoffset, +tgtsize).  Results: `status` DEFLT_OK with `srcused` = srcsize and
 * `tgtused` = bytes written, or DEFLT_TGTEXHSTD when the stream does not fit tgtsize
 * (nothing is written, srcused = tgtused = 0; srcsize + srcsize / 64 + 80 always fits);
 * `checksum` is the Adler-32 of the record for JDB200_ZLIB.  `level` as deflator_create.
 * Returns 0 when the batch ran, non-zero for a runtime failure.
 */
JDEFLATE_API
int jdb200_deflate_batch(const uint8* source, uint8* target,
                         const TJDB200Item* items, TJDB200Result* results,
                         uintxx count, eJDB200Format format, intxx level);

/* CUDA device used by instances created afterwards by this thread's process
 * (default: $JDB200_DEVICE, else $LOCAL_RANK, else 0) */
JDEFLATE_API
int jdb200_set_device(int ordinal);

JDEFLATE_API
int jdb200_device_count(void);

/*
 * Launch accounting.  Every kernel launch of the library is counted per kernel;
 * with profiling enabled each launch is also bracketed by CUDA events on the
 * stream it runs on and the durations are summed (used by bench.py for the
 * roofline of the dominant kernel).  jdb200_profile() resets the counters.
 */
typedef struct TJDB200KernelStat {
	char   name[48];
	uint64 launches;
	double ms;
} TJDB200KernelStat;

JDEFLATE_API
int jdb200_profile(int enable);

/* fills up to `max` entries, returns how many; call after the work has finished */
JDEFLATE_API
int jdb200_profile_read(TJDB200KernelStat* stats, int max);

/* last runtime error text of the calling thread ("" when none) */
JDEFLATE_API
const char* jdb200_last_error(void);

#ifdef __cplusplus
}
#endif

#endif
