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
int         jdb_rt_get_device(void);
int         jdb_rt_device_count(void);

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

#ifdef __cplusplus
}
#endif

#endif
