/* runtime_emu.cpp -- TEST INFRASTRUCTURE ONLY: host-memory stand-in for
 * csrc/device/runtime.cu used by the SIMT emulator build (see simt_emu.h). */
#include "../../jdeflate_b200/csrc/device/jdb_device.h"
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <map>

static char g_err[256];
static std::map<const void*, size_t> g_dev;   /* "device" allocations */

extern "C" void jdb_rt_set_error(const char* fmt, ...)
{
	va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char* jdb_rt_last_error(void) { return g_err; }
extern "C" int jdb_rt_check_launch(const char*) { return JDB_OK; }
extern "C" int jdb_rt_init(void) { return JDB_OK; }
extern "C" int jdb_rt_sm_count(void) { const char* e = getenv("JDB_EMU_SMS"); return e ? atoi(e) : 4; }
extern "C" int jdb_rt_get_device(void) { return 0; }
extern "C" int jdb_rt_device_count(void) { return 1; }
extern "C" int jdb_rt_set_device(int) { return JDB_OK; }
extern "C" int jdb_rt_use_device(int) { return JDB_OK; }
extern "C" int jdb_rt_current_device(void) { return 0; }
extern "C" void* jdb_dev_alloc(size_t bytes)
{
	if (!bytes) bytes = 16;
	void* p = aligned_alloc(256, (bytes + 255) / 256 * 256);
	if (p) { memset(p, 0xa5, bytes); g_dev[p] = bytes; }
	return p;
}
extern "C" void jdb_dev_free(void* p) { if (p) { g_dev.erase(p); free(p); } }
extern "C" void* jdb_pinned_alloc(size_t bytes) { return malloc(bytes ? bytes : 16); }
extern "C" void jdb_pinned_free(void* p) { free(p); }
extern "C" int jdb_ptr_is_device(const void* p)
{
	auto it = g_dev.upper_bound(p);
	if (it == g_dev.begin()) return 0;
	--it;
	return (const char*) p < (const char*) it->first + it->second;
}
extern "C" int jdb_ptr_is_pinned(const void*) { return 0; }
extern "C" int jdb_stream_create(jdb_stream* s) { *s = (jdb_stream) 1; return JDB_OK; }
extern "C" void jdb_stream_destroy(jdb_stream) {}
extern "C" int jdb_stream_sync(jdb_stream) { return JDB_OK; }
extern "C" int jdb_copy_async(void* d, const void* s, size_t n, jdb_stream) { memmove(d, s, n); return JDB_OK; }
extern "C" int jdb_memset_async(void* d, int v, size_t n, jdb_stream) { memset(d, v, n); return JDB_OK; }
extern "C" int jdb_event_create(jdb_event* e) { *e = (jdb_event) 1; return JDB_OK; }
extern "C" void jdb_event_destroy(jdb_event) {}
extern "C" int jdb_event_record(jdb_event, jdb_stream) { return JDB_OK; }
extern "C" int jdb_stream_wait_event(jdb_stream, jdb_event) { return JDB_OK; }
extern "C" int jdb_event_sync(jdb_event) { return JDB_OK; }

extern "C" int  jdb_prof_begin(const char*, int*, jdb_stream) { return -1; }
extern "C" void jdb_prof_end(int, jdb_stream) {}
extern "C" int  jdb_prof_enable(int) { return JDB_OK; }
extern "C" int  jdb_prof_read(jdb_kernel_stat*, int) { return 0; }
