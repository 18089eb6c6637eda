/*
 * b200.c -- additive entry points declared in jdeflate/b200.h: batched inflate
 * and device selection.  Host orchestration only; decoding is inflate.cu.
 */
#include <jdeflate/b200.h>
#include <string.h>
#include "jdb_host.h"

typedef char jdb_item_layout_check[(sizeof(TJDB200Item) == sizeof(jdb_inflate_item)) ? 1 : -1];
typedef char jdb_result_layout_check[(sizeof(TJDB200Result) == sizeof(jdb_inflate_result)) ? 1 : -1];

static __thread struct {
	jdb_stream stream;
	jdb_dbuf src, dst, items, results;
	uint32_t* counter;
	int ready;
} bt;

static int
batch_prepare(void)
{
	if (jdb_rt_init() != JDB_OK) {
		return JDB_ENODEV;
	}
	if (bt.ready) {
		return JDB_OK;
	}
	if (jdb_stream_create(&bt.stream) != JDB_OK) {
		return JDB_ECUDA;
	}
	bt.counter = (uint32_t*) jdb_dev_alloc(64);
	if (bt.counter == NULL) {
		return JDB_ENOMEM;
	}
	bt.ready = 1;
	return JDB_OK;
}

int
jdb200_inflate_batch(const uint8* source, uint8* target,
                     const TJDB200Item* items, TJDB200Result* results,
                     uintxx count, eJDB200Format format)
{
	const uint8* dsrc = source;
	uint8* ddst = target;
	const jdb_inflate_item* ditems = (const jdb_inflate_item*) items;
	jdb_inflate_result* dresults = (jdb_inflate_result*) results;
	uint64 srcspan = 0, dstspan = 0;
	int hostsrc, hostdst, hostitems, hostresults;
	int r;

	if (count == 0) {
		return 0;
	}
	if (count > 0xffffffffu || source == NULL || target == NULL || items == NULL || results == NULL) {
		return JDB_EARG;
	}
	r = batch_prepare();
	if (r != JDB_OK) {
		return r;
	}

	hostsrc = !jdb_ptr_is_device(source);
	hostdst = !jdb_ptr_is_device(target);
	hostitems = !jdb_ptr_is_device(items);
	hostresults = !jdb_ptr_is_device(results);

	if (hostsrc || hostdst) {
		/* spans of the host buffers come from the item list, which then has to
		 * be readable here */
		uintxx i;
		if (!hostitems) {
			return JDB_EARG;
		}
		for (i = 0; i < count; i++) {
			uint64 se = items[i].srcoffset + items[i].srcsize;
			uint64 te = items[i].tgtoffset + items[i].tgtsize;
			if (se > srcspan) srcspan = se;
			if (te > dstspan) dstspan = te;
		}
	}
	if (hostsrc) {
		if (jdb_dbuf_reserve(&bt.src, (size_t) srcspan + 16) != 0) return JDB_ENOMEM;
		if (jdb_copy_async(bt.src.ptr, source, (size_t) srcspan, bt.stream) != JDB_OK) return JDB_ECUDA;
		dsrc = bt.src.ptr;
	}
	if (hostdst) {
		if (jdb_dbuf_reserve(&bt.dst, (size_t) dstspan + 16) != 0) return JDB_ENOMEM;
		ddst = bt.dst.ptr;
	}
	if (hostitems) {
		if (jdb_dbuf_reserve(&bt.items, count * sizeof(jdb_inflate_item)) != 0) return JDB_ENOMEM;
		if (jdb_copy_async(bt.items.ptr, items, count * sizeof(jdb_inflate_item), bt.stream) != JDB_OK) return JDB_ECUDA;
		ditems = (const jdb_inflate_item*) bt.items.ptr;
	}
	if (hostresults) {
		if (jdb_dbuf_reserve(&bt.results, count * sizeof(jdb_inflate_result)) != 0) return JDB_ENOMEM;
		dresults = (jdb_inflate_result*) bt.results.ptr;
	}

	r = jdb_inflate_batch(dsrc, ddst, ditems, dresults, NULL, (uint32_t) count,
	                      format == JDB200_ZLIB ? JDB_FMT_ZLIB : JDB_FMT_RAW, 1, bt.counter, bt.stream);
	if (r != JDB_OK) {
		return r;
	}
	if (hostresults) {
		if (jdb_copy_async(results, dresults, count * sizeof(jdb_inflate_result), bt.stream) != JDB_OK) return JDB_ECUDA;
	}
	if (hostdst) {
		if (jdb_copy_async(target, ddst, (size_t) dstspan, bt.stream) != JDB_OK) return JDB_ECUDA;
	}
	return jdb_stream_sync(bt.stream);
}

int
jdb200_set_device(int ordinal)
{
	return jdb_rt_set_device(ordinal);
}

int
jdb200_device_count(void)
{
	return jdb_rt_device_count();
}

const char*
jdb200_last_error(void)
{
	return jdb_rt_last_error();
}

typedef char jdb_stat_layout_check[(sizeof(TJDB200KernelStat) == sizeof(jdb_kernel_stat)) ? 1 : -1];

int
jdb200_profile(int enable)
{
	if (jdb_rt_init() != JDB_OK) {
		return JDB_ENODEV;
	}
	return jdb_prof_enable(enable);
}

int
jdb200_profile_read(TJDB200KernelStat* stats, int max)
{
	return jdb_prof_read((jdb_kernel_stat*) stats, max);
}
