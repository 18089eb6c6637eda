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
jdb_fatal(const char* what)
{
	fprintf(stderr, "jdeflate-b200: fatal: %s (%s)\n", what, jdb_rt_last_error());
	abort();
}
