/*
 * jdb_device.h -- internal C ABI between the C99 host layer (csrc/host/*.c)
 * and the CUDA translation units (csrc/device/*.cu).
 *
 * Plain pointers and sizes only.  Every function returns 0 on success or a
 * JDB_E* code; nothing here falls back to the CPU: when no CUDA device is
 * usable jdb_rt_init() fails and the public constructors return NULL.
 */
#ifndef JDB_DEVICE_H
#define JDB_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
	JDB_OK       = 0,
	JDB_ENODEV   = 1,   /* no CUDA device / driver            */
	JDB_ENOMEM   = 2,   /* cudaMalloc / cudaHostAlloc failed  */
	JDB_ECUDA    = 3,   /* any other CUDA runtime error       */
	JDB_EARG     = 4
};

typedef void* jdb_stream;

/* ---- runtime ---------------------------------------------------------- */
int         jdb_rt_init(void);                 /* idempotent */
const char* jdb_rt_last_error(void);
int         jdb_rt_sm_count(void);
int         jdb_rt_set_device(int ordinal);
int         jdb_rt_get_device(void);          /* default device of new instances */
int         jdb_rt_use_device(int ordinal);   /* bind the calling thread to an instance's device */
int         jdb_rt_current_device(void);      /* device the calling thread is bound to */
int         jdb_rt_device_count(void);

/* per-kernel launch counters and (optional) CUDA-event timing */
typedef struct jdb_kernel_stat {
	char     name[48];
	uint64_t launches;
	double   ms;          /* sum of event-timed durations (profiling on) */
} jdb_kernel_stat;
int  jdb_prof_enable(int on);                          /* resets the counters */
int  jdb_prof_read(jdb_kernel_stat* out, int max);     /* caller has synchronised */

void* jdb_dev_alloc(size_t bytes);
void  jdb_dev_free(void* p);
void* jdb_pinned_alloc(size_t bytes);
void  jdb_pinned_free(void* p);
/* 1: device (or managed) memory, 0: host memory (pageable or pinned) */
int   jdb_ptr_is_device(const void* p);
/* 1 when a host pointer is page-locked (async copies really are async) */
int   jdb_ptr_is_pinned(const void* p);

int   jdb_stream_create(jdb_stream* s);
void  jdb_stream_destroy(jdb_stream s);
int   jdb_stream_sync(jdb_stream s);
/* direction inferred from the pointers (cudaMemcpyDefault) */
int   jdb_copy_async(void* dst, const void* src, size_t bytes, jdb_stream s);
int   jdb_memset_async(void* dst, int value, size_t bytes, jdb_stream s);

/* events for pipelining batches across two streams */
typedef void* jdb_event;
int   jdb_event_create(jdb_event* e);
void  jdb_event_destroy(jdb_event e);
int   jdb_event_record(jdb_event e, jdb_stream s);
int   jdb_stream_wait_event(jdb_stream s, jdb_event e);
int   jdb_event_sync(jdb_event e);              /* host waits for the event */

/* ---- checksums (checksum.cu) ------------------------------------------ */
/*
 * Workspace needed by jdb_checksum(): partial sums of every CTA.
 */
size_t jdb_checksum_workspace_bytes(void);

enum { JDB_CK_CRC32 = 1, JDB_CK_ADLER32 = 2 };

/*
 * Chunk-parallel CRC-32 / Adler-32 of `n` bytes at device pointer `data`.
 *   which     JDB_CK_* mask
 *   crc_io    device u32: in = running NON-finalised crc register,
 *             out = register after the bytes   (reference zstrm_crc32update
 *             semantics, src/zstrm.c:1489-1526)
 *   adler_io  device u32: in/out plain Adler-32 value (src/zstrm.c:1346-1399,
 *             but RFC 1950 exact for every size)
 *   work      device workspace of jdb_checksum_workspace_bytes()
 * Two launches on `s`: per-CTA partials, then the ordered combine.
 */
int jdb_checksum(const uint8_t* data, size_t n, int which,
                 uint32_t* crc_io, uint32_t* adler_io,
                 void* work, jdb_stream s);


/* ---- inflate (inflate.cu) ---------------------------------------------- */
#define JDB_INFLATE_HISTORY 32768u     /* DEFLATE window, power of two */
#define JDB_INF_LIT_ROOT    9          /* the reference uses 10 (LROOTBITS, src/inflator.c:30): 9 halves the table, which
                                        * buys a 4 KiB output ring per warp at 16 warps per SM */
#define JDB_INF_LIT_TABLE   864        /* >= 852 = zlib's ENOUGH for 286 symbols, root 9, 15-bit codes */
#define JDB_INF_DIST_TABLE  416        /* >= ENOUGHD 400  (root 8),  reference src/inflator.c:50-62 */

enum { JDB_INF_HEADER = 0, JDB_INF_STORED = 1, JDB_INF_SYMBOLS = 2 };
enum { JDB_FMT_RAW = 0, JDB_FMT_ZLIB = 1 };
/* container level problems use the ZSTRM_E* numbering of jdeflate/zstrm.h */
enum { JDB_ZERR_BADDATA = 3, JDB_ZERR_CHECKSUM = 4, JDB_ZERR_MISSINGDICT = 6 };

/* one stream of a batch: byte ranges relative to the two base pointers */
typedef struct jdb_inflate_item {
	uint64_t src_off;
	uint64_t dst_off;
	uint64_t src_len;
	uint64_t dst_cap;
} jdb_inflate_item;

/* per stream outcome: INFLT_* status / error codes of jdeflate/inflator.h */
typedef struct jdb_inflate_result {
	uint32_t status;
	uint32_t error;
	uint32_t zerror;      /* 0 or JDB_ZERR_* (container formats only) */
	uint32_t checksum;    /* Adler-32 of the output (JDB_FMT_ZLIB)    */
	uint64_t consumed;    /* source bytes used, trailer included      */
	uint64_t produced;
} jdb_inflate_result;

/*
 * Device resident state of a resumable (streaming) decode: what the reference
 * keeps in TINFLTPrvt between inflator_inflate calls (src/inflator.c:73-149):
 * bit buffer, position in the block structure, a match cut by the end of the
 * target, the tables of the block in progress and the last 32 KiB of output.
 */
typedef struct jdb_inflate_state {
	uint64_t bitbuf;
	uint64_t total_out;       /* absolute output position (ring index base) */
	uint64_t hist_avail;      /* bytes of history usable as match source    */
	uint32_t bitcnt;
	uint32_t phase;           /* JDB_INF_*                                   */
	uint32_t lastblock;
	uint32_t stored_left;
	uint32_t pend_len;
	uint32_t pend_dist;
	uint32_t lit[JDB_INF_LIT_TABLE];
	uint32_t dist[JDB_INF_DIST_TABLE];
	uint8_t  history[JDB_INFLATE_HISTORY];
} jdb_inflate_state;

/*
 * Decode `count` independent streams, one warp each (dynamic assignment).
 *   states   NULL for one-shot streams; else one jdb_inflate_state per stream
 *            (zero-initialised before the first call) for resumable decoding
 *   final    non-zero: running out of input is INFLT_EINPUTEND, not SRCEXHSTD
 *   counter  device u32 scratch (work distribution)
 */
int jdb_inflate_batch(const uint8_t* src_base, uint8_t* dst_base,
                      const jdb_inflate_item* items, jdb_inflate_result* results,
                      jdb_inflate_state* states, uint32_t count, uint32_t format,
                      uint32_t final, uint32_t* counter, jdb_stream s);

/*
 * The same for streams that deserve a whole CTA each (inflate_wide_kernel): one resumable raw
 * DEFLATE (or zlib) stream per item, `count` CTAs.  This is what inflator_inflate runs for a
 * stream that carries no chunk markers, and jdb200_inflate_batch for a batch of fewer streams
 * than the device has SMs.
 */
int jdb_inflate_wide(const uint8_t* src_base, uint8_t* dst_base,
                     const jdb_inflate_item* items, jdb_inflate_result* results,
                     jdb_inflate_state* states, uint32_t count, uint32_t format, uint32_t final, jdb_stream s);

/* chunk discovery for the parallel decode of one large stream (inflate.cu) */
int jdb_marker_scan(const uint8_t* src, uint64_t n, uint32_t* ends, uint32_t max_ends,
                    uint32_t* count, jdb_stream s);
#define JDB_INF_ST_MARKER 4u       /* jdb_inflate_chunks: stopped after an empty stored block */
/* every item is decoded into its own dst range and stops after the first empty stored block it
 * reads: status JDB_INF_ST_MARKER, `consumed` up to and including the marker, `produced` bytes
 * written, `error` = 1 when the marker carried BFINAL */
int jdb_inflate_chunks(const uint8_t* src_base, uint8_t* dst_base,
                       const jdb_inflate_item* items, jdb_inflate_result* results,
                       uint32_t count, uint32_t* counter, jdb_stream s);

/* ---- deflate pipeline stages (lz.cu, huffman.cu, pack.cu) ---------------- */
#define JDB_SEG 8192u      /* LZ segment: positions per CTA, histogram granule */

/* links to the previous same-hash position (u16 distance, 0 = none), one per
 * input byte; `range` (divides chunk_bytes) is the work item of one warp */
int jdb_lz_chain(const uint8_t* in, uint64_t n, uint32_t chunk_bytes, uint32_t range,
                 const uint32_t* chunk_len, uint16_t* prev, uint16_t* heads, jdb_stream s);
/* `heads` (jdb_lz_chain_heads_bytes, or NULL): scratch for the head tables the ranges leave behind;
 * with it no range replays the 32 KiB in front of it (chain_fix_kernel links across ranges) */
size_t jdb_lz_chain_heads_bytes(uint64_t n, uint32_t chunk_bytes);

/* match search + parse: tokens of segment k at tok[k*JDB_SEG ...), their count
 * in seg_ntok[k], the 320-bin symbol histogram in seg_hist[k*320 ...) */
int jdb_lz_parse(const uint8_t* in, uint64_t n, uint32_t chunk_bytes,
                 const uint32_t* chunk_len, const uint16_t* prev,
                 uint32_t good, uint32_t nice, uint32_t chain, uint32_t lazy,
                 uint32_t skip_segs, uint32_t hist_min,
                 uint32_t* tok, uint32_t* seg_ntok, uint32_t* seg_hist, jdb_stream s);

/* per block Huffman code construction + block type choice */
int jdb_huffman_blocks(const uint32_t* seg_ntok, const uint32_t* seg_hist, uint64_t n,
                       uint32_t chunk_bytes, uint32_t block_segs, uint32_t nblocks,
                       uint32_t level, uint32_t fixedonly, uint32_t skip_blocks,
                       const uint32_t* chunk_len, void* blocks, jdb_stream s);

/* ---- whole pipeline (deflate.cu) ---------------------------------------- */
typedef struct jdb_deflate_cfg {
	uint32_t level;          /* 0..9 (0 = stored only)                              */
	uint32_t fixedonly;      /* DEFLT_FIXEDCODES                                    */
	uint32_t good, nice, chain, lazy;   /* reference setparameters(), src/deflator.c:241-263 */
	uint32_t chunk_bytes;    /* independent chunk, multiple of JDB_SEG              */
	uint32_t block_segs;     /* segments per DEFLATE block, 1..16                   */
	uint32_t chain_range;    /* positions per chain-building warp (0: whole chunk)  */
	uint32_t final;          /* last chunk closes the stream (BFINAL = 1)           */
	/* preset dictionary (first batch of a stream only): the batch starts with
	 * `dict_region` bytes -- zero padding, then the dictionary -- that are history
	 * for the first chunk but produce no output; a multiple of the block size */
	uint32_t dict_region;
	uint32_t dict_pad;       /* bytes of padding in front of the dictionary          */
	/* ragged chunks (records, see deflate.cuh): device array of n / chunk_bytes entries,
	 * or NULL; wrap_head / wrap_tail bytes are left free in the output in front of the
	 * first / after the last chunk of every record (container header and trailer) */
	const uint32_t* chunk_len;
	uint32_t wrap_head, wrap_tail;
} jdb_deflate_cfg;

size_t jdb_deflate_workspace_bytes(uint64_t n, const jdb_deflate_cfg* cfg);

/*
 * Compress the batch in[0..n) (device memory, 16-byte aligned).  On return
 * *out points at the compressed bytes inside `work` and *total_dev at the u64
 * byte count, both valid once the stream has been synchronised.
 */
int jdb_deflate_run(const uint8_t* in, uint64_t n, const jdb_deflate_cfg* cfg,
                    void* work, uint8_t** out, uint64_t** total_dev, jdb_stream s);

/* ---- batched records (records.cu) ---------------------------------------- */
size_t jdb_records_workspace_bytes(uint64_t slot_bytes, const jdb_deflate_cfg* cfg);
int jdb_records_deflate(const uint8_t* src_base, uint8_t* tgt_base,
                        const jdb_inflate_item* items, jdb_inflate_result* results,
                        const uint32_t* first_chunk, uint32_t nrec,
                        uint32_t chunk_base, uint32_t nchunks, uint32_t format,
                        const jdb_deflate_cfg* cfg, uint8_t* slots, uint32_t* chunk_len,
                        void* work, jdb_stream s);

#ifdef __cplusplus
}
#endif

#endif
