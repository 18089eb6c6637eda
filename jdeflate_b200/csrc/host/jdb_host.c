/* jdb_host.c -- small host-side utilities, see jdb_host.h */
#include "jdb_host.h"
#include <stdio.h>
#include <stdlib.h>

int
jdb_dbuf_reserve(jdb_dbuf* b, size_t bytes)
{
	size_t cap;

	if (b->cap >= bytes && b->ptr != NULL) {
		return 0;
	}
	jdb_dev_free(b->ptr);
	b->ptr = NULL;
	b->cap = 0;

	cap = (bytes + 4095) & ~(size_t) 4095;
	b->ptr = (uint8_t*) jdb_dev_alloc(cap);
	if (b->ptr == NULL) {
		return JDB_ENOMEM;
	}
	b->cap = cap;
	return 0;
}

void
jdb_dbuf_release(jdb_dbuf* b)
{
	jdb_dev_free(b->ptr);
	b->ptr = NULL;
	b->cap = 0;
}

void
jdb_level_params(jdb_deflate_cfg* cfg, int level)
{
	static const uint16_t t[10][3] = {
		{0, 0, 0}, {8, 4, 2}, {8, 8, 8}, {8, 16, 16}, {8, 32, 32}, {8, 64, 128},
		{16, 16, 48}, {32, 64, 128}, {64, 128, 320}, {192, 256, 512}
	};
	cfg->level = (uint32_t) level;
	cfg->good = t[level][0];
	cfg->nice = t[level][1];
	cfg->chain = t[level][2];
	cfg->lazy = level >= 6;
}

void
jdb_fatal(const char* what)
{
	fprintf(stderr, "jdeflate-b200: fatal: %s (%s)\n", what, jdb_rt_last_error());
	abort();
}
