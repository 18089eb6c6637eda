/*
 * runtime.cu -- device discovery, memory, streams and copies for the
 * B200-native jdeflate build.  Thin on purpose: the host layer is C99 and
 * only sees the plain C ABI of jdb_device.h.
 */
#include "common.cuh"
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_err[256];
static int g_init_state = 0;      /* 0 unknown, 1 ok, -1 failed */
static int g_device = -1;
static int g_sm_count = 0;

extern "C" void jdb_rt_set_error(const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char* jdb_rt_last_error(void) { return g_err; }

static int check(cudaError_t e, const char* what)
{
	if (e == cudaSuccess) return JDB_OK;
	jdb_rt_set_error("%s: %s", what, cudaGetErrorString(e));
	if (e == cudaErrorMemoryAllocation) return JDB_ENOMEM;
	return JDB_ECUDA;
}

extern "C" int jdb_rt_check_launch(const char* what)
{
	return check(cudaGetLastError(), what);
}

extern "C" int jdb_rt_init(void)
{
	if (g_init_state == 1) {
		/* keep the calling thread on the chosen device */
		cudaSetDevice(g_device);
		return JDB_OK;
	}
	if (g_init_state == -1) return JDB_ENODEV;

	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count <= 0) {
		jdb_rt_set_error("jdeflate-b200: no usable CUDA device (%s); this library has no CPU path",
		                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
		fprintf(stderr, "%s\n", g_err);
		g_init_state = -1;
		return JDB_ENODEV;
	}
	int dev = 0;
	const char* env = getenv("JDB200_DEVICE");
	if (g_device >= 0) dev = g_device;
	else if (env && *env) dev = atoi(env);
	else if (getenv("LOCAL_RANK")) dev = atoi(getenv("LOCAL_RANK")) % count;
	if (dev < 0 || dev >= count) dev = 0;
	if (check(cudaSetDevice(dev), "cudaSetDevice") != JDB_OK) { g_init_state = -1; return JDB_ENODEV; }
	cudaDeviceProp prop;
	if (check(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties") != JDB_OK) {
		g_init_state = -1;
		return JDB_ENODEV;
	}
	g_sm_count = prop.multiProcessorCount;
	g_device = dev;
	g_init_state = 1;
	return JDB_OK;
}

extern "C" int jdb_rt_sm_count(void) { return g_sm_count > 0 ? g_sm_count : 148; }
extern "C" int jdb_rt_get_device(void) { return g_device; }

extern "C" int jdb_rt_device_count(void)
{
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
	return count;
}

extern "C" int jdb_rt_set_device(int ordinal)
{
	int count = jdb_rt_device_count();
	if (ordinal < 0 || ordinal >= count) return JDB_EARG;
	g_device = ordinal;
	if (g_init_state == 1) {
		cudaDeviceProp prop;
		if (check(cudaSetDevice(ordinal), "cudaSetDevice") != JDB_OK) return JDB_ECUDA;
		if (cudaGetDeviceProperties(&prop, ordinal) == cudaSuccess) g_sm_count = prop.multiProcessorCount;
		return JDB_OK;
	}
	return jdb_rt_init();
}

extern "C" void* jdb_dev_alloc(size_t bytes)
{
	void* p = NULL;
	if (bytes == 0) bytes = 16;
	if (check(cudaMalloc(&p, bytes), "cudaMalloc") != JDB_OK) return NULL;
	return p;
}

extern "C" void jdb_dev_free(void* p) { if (p) cudaFree(p); }

extern "C" void* jdb_pinned_alloc(size_t bytes)
{
	void* p = NULL;
	if (bytes == 0) bytes = 16;
	if (check(cudaHostAlloc(&p, bytes, cudaHostAllocDefault), "cudaHostAlloc") != JDB_OK) return NULL;
	return p;
}

extern "C" void jdb_pinned_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int jdb_ptr_is_device(const void* p)
{
	cudaPointerAttributes a;
	cudaError_t e = cudaPointerGetAttributes(&a, p);
	if (e != cudaSuccess) { cudaGetLastError(); return 0; }
	return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

extern "C" int jdb_ptr_is_pinned(const void* p)
{
	cudaPointerAttributes a;
	cudaError_t e = cudaPointerGetAttributes(&a, p);
	if (e != cudaSuccess) { cudaGetLastError(); return 0; }
	return a.type == cudaMemoryTypeHost;
}

extern "C" int jdb_stream_create(jdb_stream* s)
{
	cudaStream_t st;
	int r = check(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate");
	if (r != JDB_OK) return r;
	*s = (jdb_stream) st;
	return JDB_OK;
}

extern "C" void jdb_stream_destroy(jdb_stream s) { if (s) cudaStreamDestroy((cudaStream_t) s); }

extern "C" int jdb_stream_sync(jdb_stream s)
{
	return check(cudaStreamSynchronize((cudaStream_t) s), "cudaStreamSynchronize");
}

extern "C" int jdb_copy_async(void* dst, const void* src, size_t bytes, jdb_stream s)
{
	if (bytes == 0) return JDB_OK;
	return check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t) s), "cudaMemcpyAsync");
}

extern "C" int jdb_memset_async(void* dst, int value, size_t bytes, jdb_stream s)
{
	if (bytes == 0) return JDB_OK;
	return check(cudaMemsetAsync(dst, value, bytes, (cudaStream_t) s), "cudaMemsetAsync");
}

extern "C" int jdb_event_create(jdb_event* e)
{
	cudaEvent_t ev;
	int r = check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate");
	if (r != JDB_OK) return r;
	*e = (jdb_event) ev;
	return JDB_OK;
}

extern "C" void jdb_event_destroy(jdb_event e) { if (e) cudaEventDestroy((cudaEvent_t) e); }

extern "C" int jdb_event_record(jdb_event e, jdb_stream s)
{
	return check(cudaEventRecord((cudaEvent_t) e, (cudaStream_t) s), "cudaEventRecord");
}

extern "C" int jdb_event_sync(jdb_event e)
{
	return check(cudaEventSynchronize((cudaEvent_t) e), "cudaEventSynchronize");
}

extern "C" int jdb_stream_wait_event(jdb_stream s, jdb_event e)
{
	return check(cudaStreamWaitEvent((cudaStream_t) s, (cudaEvent_t) e, 0), "cudaStreamWaitEvent");
}


/* ---- launch accounting / profiling ------------------------------------------ */
#include <pthread.h>

#define PROF_KERNELS 24
#define PROF_PENDING 4096

static pthread_mutex_t g_prof_lock = PTHREAD_MUTEX_INITIALIZER;
static int g_prof_on = 0;
static int g_prof_n = 0;
static jdb_kernel_stat g_prof[PROF_KERNELS];
static struct { cudaEvent_t a, b; int kernel; int used; } g_pend[PROF_PENDING];
static int g_pend_n = 0;

static int prof_kernel_index(const char* name)
{
	/* `name` is the macro-stringified kernel expression: strip parentheses / template args */
	char clean[48];
	size_t k = 0;
	for (const char* p = name; *p && k + 1 < sizeof(clean); p++) {
		if (*p == '(' || *p == ')' || *p == ' ') continue;
		clean[k++] = *p;
	}
	clean[k] = 0;
	for (int i = 0; i < g_prof_n; i++)
		if (strcmp(g_prof[i].name, clean) == 0) return i;
	if (g_prof_n == PROF_KERNELS) return PROF_KERNELS - 1;
	memset(&g_prof[g_prof_n], 0, sizeof(g_prof[0]));
	strcpy(g_prof[g_prof_n].name, clean);
	return g_prof_n++;
}

static void prof_resolve_locked(void)
{
	for (int i = 0; i < g_pend_n; i++) {
		if (!g_pend[i].used) continue;
		float ms = 0;
		if (cudaEventSynchronize(g_pend[i].b) == cudaSuccess &&
		    cudaEventElapsedTime(&ms, g_pend[i].a, g_pend[i].b) == cudaSuccess)
			g_prof[g_pend[i].kernel].ms += ms;
		g_pend[i].used = 0;
	}
	g_pend_n = 0;
	cudaGetLastError();
}

extern "C" int jdb_prof_begin(const char* kernel, jdb_stream s)
{
	pthread_mutex_lock(&g_prof_lock);
	int k = prof_kernel_index(kernel);
	g_prof[k].launches++;
	int slot = -1;
	if (g_prof_on) {
		if (g_pend_n == PROF_PENDING) prof_resolve_locked();
		slot = g_pend_n++;
		if (!g_pend[slot].a) {
			cudaEventCreate(&g_pend[slot].a);
			cudaEventCreate(&g_pend[slot].b);
		}
		g_pend[slot].kernel = k;
		g_pend[slot].used = 1;
		cudaEventRecord(g_pend[slot].a, (cudaStream_t) s);
	}
	pthread_mutex_unlock(&g_prof_lock);
	return slot;
}

extern "C" void jdb_prof_end(int slot, jdb_stream s)
{
	if (slot < 0) return;
	cudaEventRecord(g_pend[slot].b, (cudaStream_t) s);
}

extern "C" int jdb_prof_enable(int on)
{
	pthread_mutex_lock(&g_prof_lock);
	prof_resolve_locked();
	for (int i = 0; i < g_prof_n; i++) { g_prof[i].launches = 0; g_prof[i].ms = 0; }
	g_prof_on = on != 0;
	pthread_mutex_unlock(&g_prof_lock);
	return JDB_OK;
}

extern "C" int jdb_prof_read(jdb_kernel_stat* out, int max)
{
	pthread_mutex_lock(&g_prof_lock);
	prof_resolve_locked();
	int n = g_prof_n < max ? g_prof_n : max;
	for (int i = 0; i < n; i++) out[i] = g_prof[i];
	pthread_mutex_unlock(&g_prof_lock);
	return n;
}
