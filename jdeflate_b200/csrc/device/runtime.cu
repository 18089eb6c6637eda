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

#include <pthread.h>

static __thread char g_err[256];
/* device discovery happens once, under a lock; afterwards the state is read-only except for
 * the default device, which jdb_rt_set_device() may move (new instances follow it, live
 * instances stay where they were created: they carry their own ordinal) */
static pthread_mutex_t g_rt_lock = PTHREAD_MUTEX_INITIALIZER;
static int g_init_state = 0;      /* 0 unknown, 1 ok, -1 failed */
static int g_device = -1;         /* default device of new instances */
static int g_device_count = 0;
static int g_sm_count[64];

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
	/* a failed call leaves its error behind for the next cudaGetLastError(): a launch that
	 * follows a failed allocation would otherwise be reported as failed, too */
	cudaGetLastError();
	if (e == cudaErrorMemoryAllocation) return JDB_ENOMEM;
	return JDB_ECUDA;
}

extern "C" int jdb_rt_check_launch(const char* what)
{
	return check(cudaGetLastError(), what);
}

static int rt_init_locked(void)
{
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count <= 0) {
		jdb_rt_set_error("jdeflate-b200: no usable CUDA device (%s); this library has no CPU path",
		                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
		fprintf(stderr, "%s\n", g_err);
		cudaGetLastError();
		return -1;
	}
	if (count > 64) count = 64;
	int dev = 0;
	const char* env = getenv("JDB200_DEVICE");
	const char* rank = getenv("LOCAL_RANK");
	if (g_device >= 0) dev = g_device;
	else if (env && *env) dev = atoi(env);
	else if (rank && *rank) dev = atoi(rank) % count;
	if (dev < 0 || dev >= count) dev = 0;
	for (int d = 0; d < count; d++) {
		int sms = 0;
		if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d) != cudaSuccess) { cudaGetLastError(); sms = 0; }
		g_sm_count[d] = sms;
	}
	if (check(cudaSetDevice(dev), "cudaSetDevice") != JDB_OK) return -1;
	g_device_count = count;
	g_device = dev;
	return 1;
}

extern "C" int jdb_rt_init(void)
{
	int st = __atomic_load_n(&g_init_state, __ATOMIC_ACQUIRE);
	if (st == 0) {
		pthread_mutex_lock(&g_rt_lock);
		if (g_init_state == 0) __atomic_store_n(&g_init_state, rt_init_locked(), __ATOMIC_RELEASE);
		st = g_init_state;
		pthread_mutex_unlock(&g_rt_lock);
	}
	if (st != 1) return JDB_ENODEV;
	/* keep the calling thread on the default device */
	return check(cudaSetDevice(__atomic_load_n(&g_device, __ATOMIC_RELAXED)), "cudaSetDevice");
}

/* bind the calling thread to the device an instance lives on */
extern "C" int jdb_rt_use_device(int ordinal)
{
	if (__atomic_load_n(&g_init_state, __ATOMIC_ACQUIRE) != 1 && jdb_rt_init() != JDB_OK) return JDB_ENODEV;
	if (ordinal < 0 || ordinal >= g_device_count) return JDB_EARG;
	return check(cudaSetDevice(ordinal), "cudaSetDevice");
}

extern "C" int jdb_rt_sm_count(void)
{
	int dev = -1;
	if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = g_device; }
	const int n = (dev >= 0 && dev < 64) ? g_sm_count[dev] : 0;
	return n > 0 ? n : 148;
}
extern "C" int jdb_rt_get_device(void) { return __atomic_load_n(&g_device, __ATOMIC_RELAXED); }

/* the device the calling thread is bound to (what a launch goes to) */
extern "C" int jdb_rt_current_device(void)
{
	int dev = -1;
	if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
	return dev;
}

extern "C" int jdb_rt_device_count(void)
{
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
	return count;
}

extern "C" int jdb_rt_set_device(int ordinal)
{
	int count = jdb_rt_device_count();
	if (ordinal < 0 || ordinal >= count || ordinal >= 64) return JDB_EARG;
	pthread_mutex_lock(&g_rt_lock);
	__atomic_store_n(&g_device, ordinal, __ATOMIC_RELAXED);
	pthread_mutex_unlock(&g_rt_lock);
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
	if (check(cudaHostAlloc(&p, bytes, cudaHostAllocPortable), "cudaHostAlloc") != JDB_OK) return NULL;
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

#define PROF_KERNELS 32
#define PROF_PENDING 4096

static pthread_mutex_t g_prof_lock = PTHREAD_MUTEX_INITIALIZER;
static int g_prof_on = 0;
static int g_prof_n = 0;
static jdb_kernel_stat g_prof[PROF_KERNELS];
/* event pairs of timed launches: a ring; a slot is free, open (begin recorded, the launch is
 * being issued by some thread) or complete (both events recorded, duration not read yet) */
enum { PS_FREE = 0, PS_OPEN = 1, PS_COMPLETE = 2 };
static struct { cudaEvent_t a, b; int kernel; int state; } g_pend[PROF_PENDING];
static int g_pend_next = 0;

/* index of a kernel by the macro-stringified launch expression (parentheses / blanks stripped),
 * -1 when the table is full */
static int prof_kernel_index_locked(const char* name)
{
	char clean[48];
	size_t k = 0;
	for (const char* p = name; *p && k + 1 < sizeof(clean); p++) {
		if (*p == '(' || *p == ')' || *p == ' ') continue;
		clean[k++] = *p;
	}
	clean[k] = 0;
	for (int i = 0; i < g_prof_n; i++)
		if (strcmp(g_prof[i].name, clean) == 0) return i;
	if (g_prof_n == PROF_KERNELS) return -1;
	memset(&g_prof[g_prof_n], 0, sizeof(g_prof[0]));
	strcpy(g_prof[g_prof_n].name, clean);
	return g_prof_n++;
}

static void prof_resolve_slot_locked(int i)
{
	float ms = 0;
	if (cudaEventSynchronize(g_pend[i].b) == cudaSuccess &&
	    cudaEventElapsedTime(&ms, g_pend[i].a, g_pend[i].b) == cudaSuccess)
		g_prof[g_pend[i].kernel].ms += ms;
	else
		cudaGetLastError();
	g_pend[i].state = PS_FREE;
}

static void prof_resolve_locked(void)
{
	for (int i = 0; i < PROF_PENDING; i++)
		if (g_pend[i].state == PS_COMPLETE) prof_resolve_slot_locked(i);
}

/*
 * Called around every kernel launch (JDB_LAUNCH).  `site` is a static int of the launch site
 * that caches the kernel's table index, so the common case -- profiling off -- is one atomic
 * increment and no lock.
 */
extern "C" int jdb_prof_begin(const char* kernel, int* site, jdb_stream s)
{
	int k = __atomic_load_n(site, __ATOMIC_RELAXED);
	if (k == -1) {
		pthread_mutex_lock(&g_prof_lock);
		k = prof_kernel_index_locked(kernel);
		pthread_mutex_unlock(&g_prof_lock);
		__atomic_store_n(site, k < 0 ? -2 : k, __ATOMIC_RELAXED);
	}
	if (k < 0) return -1;
	__atomic_fetch_add(&g_prof[k].launches, 1, __ATOMIC_RELAXED);
	if (!__atomic_load_n(&g_prof_on, __ATOMIC_RELAXED)) return -1;

	int slot = -1;
	pthread_mutex_lock(&g_prof_lock);
	if (g_prof_on) {
		const int i = g_pend_next;
		if (g_pend[i].state == PS_COMPLETE) prof_resolve_slot_locked(i);
		if (g_pend[i].state == PS_FREE) {       /* an open slot of another thread: this launch goes untimed */
			if (!g_pend[i].a) {
				cudaEventCreate(&g_pend[i].a);
				cudaEventCreate(&g_pend[i].b);
			}
			g_pend[i].kernel = k;
			g_pend[i].state = PS_OPEN;
			cudaEventRecord(g_pend[i].a, (cudaStream_t) s);
			slot = i;
		}
		g_pend_next = (i + 1) % PROF_PENDING;
	}
	pthread_mutex_unlock(&g_prof_lock);
	return slot;
}

extern "C" void jdb_prof_end(int slot, jdb_stream s)
{
	if (slot < 0) return;
	pthread_mutex_lock(&g_prof_lock);
	cudaEventRecord(g_pend[slot].b, (cudaStream_t) s);
	g_pend[slot].state = PS_COMPLETE;
	pthread_mutex_unlock(&g_prof_lock);
}

extern "C" int jdb_prof_enable(int on)
{
	pthread_mutex_lock(&g_prof_lock);
	prof_resolve_locked();
	for (int i = 0; i < g_prof_n; i++) { g_prof[i].launches = 0; g_prof[i].ms = 0; }
	__atomic_store_n(&g_prof_on, on != 0, __ATOMIC_RELAXED);
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
