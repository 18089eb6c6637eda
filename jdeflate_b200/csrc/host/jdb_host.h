/*
 * jdb_host.h -- helpers shared by the C99 host layer (deflator.c, inflator.c,
 * zstrm.c, checksum.c).  Host code only orchestrates: it owns the reference's
 * state machines and moves bytes; all codec arithmetic runs in the CUDA
 * kernels behind csrc/device/jdb_device.h.
 */
#ifndef JDB_HOST_H
#define JDB_HOST_H

#include <stddef.h>
#include <stdint.h>
#include "../device/jdb_device.h"

/* grow-only device buffer */
typedef struct {
	uint8_t* ptr;
	size_t   cap;
} jdb_dbuf;

/* returns 0 on success; contents are NOT preserved when the buffer grows */
int  jdb_dbuf_reserve(jdb_dbuf* b, size_t bytes);
void jdb_dbuf_release(jdb_dbuf* b);

/* level -> (good, nice, chain), lazy for 6..9: reference setparameters(), src/deflator.c:241-263 */
void jdb_level_params(jdb_deflate_cfg* cfg, int level);

/* loud failure used where the public API has no error channel */
void jdb_fatal(const char* what);

#endif
