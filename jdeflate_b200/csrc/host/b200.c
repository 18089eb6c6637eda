/*
 * b200.c -- additive entry points declared in jdeflate/b200.h: batched inflate
 * and device selection.  Host orchestration only; decoding is inflate.cu.
 */
#include <jdeflate/b200.h>
#include <string.h>
#include <stdlib.h>
#include "jdb_host.h"

typedef char jdb_item_layout_check[(sizeof(TJDB200Item) == sizeof(jdb_inflate_item)) ? 1 : -1];
typedef char jdb_result_layout_check[(sizeof(TJDB200Result) == sizeof(jdb_inflate_result)) ? 1 : -1];

static __thread struct {
	jdb_stream stream;
	jdb_dbuf src, dst, items, results;
	jdb_dbuf slots, chunklen, firstchunk, work;      /* jdb200_deflate_batch */
	uint32_t* counter;
	int ready;
	int device;                                      /* where stream and buffers live */
	int narrow;                                      /* JDB200_INFLATE_NARROW=1: one warp per stream whatever the count */
} bt;

static int
batch_prepare(void)
{
	if (jdb_rt_init() != JDB_OK) {
		return JDB_ENODEV;
	}
	if (bt.ready && bt.device == jdb_rt_current_device()) {
		return JDB_OK;
	}
	if (bt.ready) {
		/* the default device moved (jdb200_set_device): this thread's scratch moves with it */
		jdb_stream_sync(bt.stream);
		jdb_stream_destroy(bt.stream);
		jdb_dbuf_release(&bt.src);
		jdb_dbuf_release(&bt.dst);
		jdb_dbuf_release(&bt.items);
		jdb_dbuf_release(&bt.results);
		jdb_dbuf_release(&bt.slots);
		jdb_dbuf_release(&bt.chunklen);
		jdb_dbuf_release(&bt.firstchunk);
		jdb_dbuf_release(&bt.work);
		jdb_dev_free(bt.counter);
		bt.ready = 0;
	}
	if (jdb_stream_create(&bt.stream) != JDB_OK) {
		return JDB_ECUDA;
	}
	bt.counter = (uint32_t*) jdb_dev_alloc(64);
	if (bt.counter == NULL) {
		return JDB_ENOMEM;
	}
	bt.device = jdb_rt_current_device();
	{
		const char* e = getenv("JDB200_INFLATE_NARROW");
		bt.narrow = e != NULL && e[0] == '1';
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

	/* fewer streams than SMs: a thread block per stream (inflate_wide_kernel, ~7x a lone warp on a long
	 * stream) -- one warp per stream could not fill the device anyway */
	if (count <= (uintxx) jdb_rt_sm_count() && !bt.narrow) {
		r = jdb_inflate_wide(dsrc, ddst, ditems, dresults, NULL, (uint32_t) count,
		                     format == JDB200_ZLIB ? JDB_FMT_ZLIB : JDB_FMT_RAW, 1, bt.stream);
	}
	else {
		r = jdb_inflate_batch(dsrc, ddst, ditems, dresults, NULL, (uint32_t) count,
		                      format == JDB200_ZLIB ? JDB_FMT_ZLIB : JDB_FMT_RAW, 1, bt.counter, bt.stream);
	}
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

/*
 * Batched compression.  Records are laid into chunk slots (records.cu) and
 * compressed group by group, a group being as many records as fit the slot
 * budget; one stream, no host round trip inside a group.
 */
#define DB_GROUP_BYTES ((size_t) 256 << 20)
#define DB_MAX_CHUNK   ((size_t) 512 << 10)

int
jdb200_deflate_batch(const uint8* source, uint8* target,
                     const TJDB200Item* items, TJDB200Result* results,
                     uintxx count, eJDB200Format format, intxx level)
{
	const uint8* dsrc = source;
	uint8* ddst = target;
	const TJDB200Item* hitems = items;
	TJDB200Item* itemcopy = NULL;
	const jdb_inflate_item* ditems = (const jdb_inflate_item*) items;
	jdb_inflate_result* dresults = (jdb_inflate_result*) results;
	uint32_t* first = NULL;
	jdb_deflate_cfg cfg;
	uint64 srcspan = 0, dstspan = 0, totalbytes = 0, maxlen = 0;
	size_t chunk, budget, i;
	int hostsrc, hostdst, hostitems, hostresults;
	int r = JDB_OK;

	if (count == 0) {
		return 0;
	}
	if (count > 0x7fffffffu || source == NULL || target == NULL || items == NULL || results == NULL ||
	    level < 0 || level > 9 || (format != JDB200_RAW && format != JDB200_ZLIB)) {
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

	/* the slot plan is made here, so the item list has to be readable on the host */
	if (!hostitems) {
		itemcopy = (TJDB200Item*) malloc(count * sizeof(TJDB200Item));
		if (itemcopy == NULL) return JDB_ENOMEM;
		if (jdb_copy_async(itemcopy, items, count * sizeof(TJDB200Item), bt.stream) != JDB_OK ||
		    jdb_stream_sync(bt.stream) != JDB_OK) {
			free(itemcopy);
			return JDB_ECUDA;
		}
		hitems = itemcopy;
	}
	first = (uint32_t*) malloc(count * sizeof(uint32_t));
	if (first == NULL) {
		r = JDB_ENOMEM;
		goto L_DONE;
	}
	for (i = 0; i < count; i++) {
		uint64 se = hitems[i].srcoffset + hitems[i].srcsize;
		uint64 te = hitems[i].tgtoffset + hitems[i].tgtsize;
		if (se > srcspan) srcspan = se;
		if (te > dstspan) dstspan = te;
		if (hitems[i].srcsize > maxlen) maxlen = hitems[i].srcsize;
		totalbytes += hitems[i].srcsize;
	}
	/* chunk slot: twice the mean record, a power of two between one LZ segment and the
	 * chunk size of the streaming encoder (longer records span several chunks) */
	chunk = JDB_SEG;        /* one segment: an lz_kernel CTA then works on two records at a time */
	while (chunk < DB_MAX_CHUNK && chunk < 2 * (totalbytes / count)) chunk *= 2;
	{
		const char* e = getenv("JDB200_RECORD_CHUNK_KIB");
		if (e && atoi(e) > 0) {
			chunk = (size_t) atoi(e) << 10;
			if (chunk > JDB_SEG) chunk = chunk / (2 * JDB_SEG) * (2 * JDB_SEG);
		}
		if (chunk < JDB_SEG) chunk = JDB_SEG;
	}
	budget = DB_GROUP_BYTES;
	{
		const char* e = getenv("JDB200_BATCH_MIB");
		if (e && atoi(e) > 0) budget = (size_t) atoi(e) << 20;
	}
	{
		uint64 slot = 0;
		for (i = 0; i < count; i++) {
			uint64 nch = hitems[i].srcsize ? (hitems[i].srcsize + chunk - 1) / chunk : 1;
			if (slot + nch > 0xfffffff0u) { r = JDB_EARG; goto L_DONE; }
			first[i] = (uint32_t) slot;
			slot += nch;
		}
	}

	memset(&cfg, 0, sizeof cfg);
	jdb_level_params(&cfg, (int) level);
	cfg.chunk_bytes = (uint32_t) chunk;
	cfg.block_segs = 8;

	if (hostsrc) {
		if (jdb_dbuf_reserve(&bt.src, (size_t) srcspan + 16) != 0) { r = JDB_ENOMEM; goto L_DONE; }
		if (jdb_copy_async(bt.src.ptr, source, (size_t) srcspan, bt.stream) != JDB_OK) { r = JDB_ECUDA; goto L_DONE; }
		dsrc = bt.src.ptr;
	}
	if (hostdst) {
		if (jdb_dbuf_reserve(&bt.dst, (size_t) dstspan + 16) != 0) { r = JDB_ENOMEM; goto L_DONE; }
		ddst = bt.dst.ptr;
	}
	if (hostitems) {
		if (jdb_dbuf_reserve(&bt.items, count * sizeof(jdb_inflate_item)) != 0) { r = JDB_ENOMEM; goto L_DONE; }
		if (jdb_copy_async(bt.items.ptr, items, count * sizeof(jdb_inflate_item), bt.stream) != JDB_OK) { r = JDB_ECUDA; goto L_DONE; }
		ditems = (const jdb_inflate_item*) bt.items.ptr;
	}
	if (hostresults) {
		if (jdb_dbuf_reserve(&bt.results, count * sizeof(jdb_inflate_result)) != 0) { r = JDB_ENOMEM; goto L_DONE; }
		dresults = (jdb_inflate_result*) bt.results.ptr;
	}
	if (jdb_dbuf_reserve(&bt.firstchunk, count * sizeof(uint32_t)) != 0) { r = JDB_ENOMEM; goto L_DONE; }
	if (jdb_copy_async(bt.firstchunk.ptr, first, count * sizeof(uint32_t), bt.stream) != JDB_OK) { r = JDB_ECUDA; goto L_DONE; }

	/* groups of consecutive records within the slot budget (a record larger than the
	 * budget is a group of its own) */
	for (i = 0; i < count;) {
		size_t j = i;
		uint64 nch = 0, slotbytes;
		size_t wbytes;
		while (j < count) {
			uint64 k = hitems[j].srcsize ? (hitems[j].srcsize + chunk - 1) / chunk : 1;
			if (j > i && (nch + k) * chunk > budget) break;
			nch += k;
			j++;
		}
		slotbytes = nch * chunk;
		wbytes = jdb_records_workspace_bytes(slotbytes, &cfg);
		if (wbytes == 0 || nch > 0x7fffffffu) { r = JDB_EARG; goto L_DONE; }
		if (jdb_dbuf_reserve(&bt.slots, (size_t) slotbytes + 64) != 0 ||
		    jdb_dbuf_reserve(&bt.chunklen, (size_t) nch * 4) != 0 ||
		    jdb_dbuf_reserve(&bt.work, wbytes) != 0) {
			r = JDB_ENOMEM;
			goto L_DONE;
		}
		r = jdb_records_deflate(dsrc, ddst, ditems + i, dresults + i,
		                        (const uint32_t*) bt.firstchunk.ptr + i, (uint32_t) (j - i),
		                        first[i], (uint32_t) nch,
		                        format == JDB200_ZLIB ? JDB_FMT_ZLIB : JDB_FMT_RAW, &cfg,
		                        bt.slots.ptr, (uint32_t*) bt.chunklen.ptr, bt.work.ptr, bt.stream);
		if (r != JDB_OK) goto L_DONE;
		i = j;
	}
	if (hostresults) {
		if (jdb_copy_async(results, dresults, count * sizeof(jdb_inflate_result), bt.stream) != JDB_OK) { r = JDB_ECUDA; goto L_DONE; }
	}
	if (hostdst) {
		if (jdb_copy_async(target, ddst, (size_t) dstspan, bt.stream) != JDB_OK) { r = JDB_ECUDA; goto L_DONE; }
	}
	r = jdb_stream_sync(bt.stream);
L_DONE:
	if (r != JDB_OK) jdb_stream_sync(bt.stream);
	free(first);
	free(itemcopy);
	return r;
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
