/*
 * cpu_baseline.c -- times the reference's CPU codec on the host cores
 * (BENCH INFRASTRUCTURE ONLY; used by bench.py's cpu_baseline / --impl
 * reference legs, never by the product).
 *
 * The implementation under test is whatever shared object is passed in: the
 * compiled reference oracle/_ref/libjdeflate_ref.so (kind "reference").  It is
 * dlopen'ed RTLD_LOCAL and driven through the public jdeflate API exactly as
 * README.md:98-184 of the reference shows: one instance per thread, one-shot
 * deflator_deflate(DEFLT_END) / inflator_inflate(final) over a contiguous
 * slice of the input (SURVEY.md section 8d "CPU baseline beside it").
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <jdeflate/deflator.h>
#include <jdeflate/inflator.h>

#define API __attribute__((visibility("default")))

typedef TDeflator* (*fn_dcreate)(uintxx, intxx, const TAllocator*);
typedef void (*fn_ddestroy)(TDeflator*);
typedef eDEFLTResult (*fn_deflate)(TDeflator*, eDEFLTFlush);
typedef TInflator* (*fn_icreate)(uintxx, const TAllocator*);
typedef void (*fn_idestroy)(TInflator*);
typedef eINFLTResult (*fn_inflate)(TInflator*, uint32);
typedef uint32 (*fn_cksum)(uint32, const uint8*, uintxx);

typedef struct {
	fn_dcreate dcreate; fn_ddestroy ddestroy; fn_deflate deflate;
	fn_icreate icreate; fn_idestroy idestroy; fn_inflate inflate;
	fn_cksum crc, adler;
} refapi;

typedef struct {
	const refapi* api;
	int op, level;
	const uint8* in; size_t n;      /* slice: raw bytes (deflate) / compressed (inflate) */
	uint8* out; size_t cap;         /* per thread output */
	size_t produced;
	int status;
	pthread_barrier_t* go;
} job;

static void*
worker(void* arg)
{
	job* j = arg;
	pthread_barrier_wait(j->go);
	if (j->op == 0) {
		TDeflator* d = j->api->dcreate(0, j->level, NULL);
		deflator_setsrc(d, j->in, j->n);
		deflator_settgt(d, j->out, j->cap);
		j->status = j->api->deflate(d, DEFLT_END);
		j->produced = deflator_tgtend(d);
		j->api->ddestroy(d);
	} else if (j->op == 1) {
		TInflator* s = j->api->icreate(0, NULL);
		inflator_setsrc(s, j->in, j->n);
		inflator_settgt(s, j->out, j->cap);
		j->status = j->api->inflate(s, 1);
		j->produced = inflator_tgtend(s);
		j->api->idestroy(s);
	} else {
		uint32 v = j->op == 2 ? j->api->crc(0xffffffffu, j->in, j->n) : j->api->adler(1, j->in, j->n);
		j->produced = v;
		j->status = 0;
	}
	pthread_barrier_wait(j->go);
	return NULL;
}

static double now(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}

/*
 * op 0 deflate, 1 inflate (input is first compressed per slice, untimed),
 * 2 crc32, 3 adler32.  `data`/`n`: uncompressed sample; `threads` instances on
 * contiguous slices.  Outputs: wall seconds of the timed op and the total
 * compressed size (ops 0/1).  Returns 0 or a negative error.
 */
API int
jdcb_run(const char* libpath, int op, int level, const uint8* data, size_t n, int threads,
         double* seconds, size_t* compressed)
{
	void* h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
	refapi api;
	job* jobs;
	pthread_t* tids;
	pthread_barrier_t go;
	int t, rc = 0;
	double t0, t1;
	size_t total = 0;

	if (!h) { fprintf(stderr, "cpu_baseline: %s\n", dlerror()); return -1; }
	api.dcreate = (fn_dcreate) dlsym(h, "deflator_create");
	api.ddestroy = (fn_ddestroy) dlsym(h, "deflator_destroy");
	api.deflate = (fn_deflate) dlsym(h, "deflator_deflate");
	api.icreate = (fn_icreate) dlsym(h, "inflator_create");
	api.idestroy = (fn_idestroy) dlsym(h, "inflator_destroy");
	api.inflate = (fn_inflate) dlsym(h, "inflator_inflate");
	api.crc = (fn_cksum) dlsym(h, "zstrm_crc32update");
	api.adler = (fn_cksum) dlsym(h, "zstrm_adler32update");
	if (!api.dcreate || !api.deflate || !api.icreate || !api.inflate || !api.crc) return -2;
	if (threads < 1) threads = 1;

	jobs = calloc((size_t) threads, sizeof(job));
	tids = calloc((size_t) threads, sizeof(pthread_t));
	for (t = 0; t < threads; t++) {
		size_t b = n * (size_t) t / (size_t) threads, e = n * (size_t) (t + 1) / (size_t) threads;
		job* j = &jobs[t];
		j->api = &api; j->op = op; j->level = level; j->go = &go;
		j->in = data + b; j->n = e - b;
		if (op <= 1) {
			j->cap = j->n + j->n / 8 + 1024;
			j->out = malloc(j->cap);
			if (!j->out) return -3;
		}
	}
	if (op == 1) {
		/* untimed: compress every slice with the same implementation */
		for (t = 0; t < threads; t++) {
			job* j = &jobs[t];
			TDeflator* d = api.dcreate(0, level, NULL);
			deflator_setsrc(d, j->in, j->n);
			deflator_settgt(d, j->out, j->cap);
			if (api.deflate(d, DEFLT_END) != DEFLT_OK) rc = -4;
			j->in = j->out;                        /* compressed slice is the input */
			j->cap = j->n;                         /* original size */
			j->n = deflator_tgtend(d);
			api.ddestroy(d);
			j->out = malloc(j->cap + 64);
			total += j->n;
		}
	}
	pthread_barrier_init(&go, NULL, (unsigned) threads + 1);
	for (t = 0; t < threads; t++) pthread_create(&tids[t], NULL, worker, &jobs[t]);
	pthread_barrier_wait(&go);
	t0 = now();
	pthread_barrier_wait(&go);
	t1 = now();
	for (t = 0; t < threads; t++) pthread_join(tids[t], NULL);
	for (t = 0; t < threads; t++) {
		if (jobs[t].status != 0) rc = -5;
		if (op == 0) total += jobs[t].produced;
		if (op == 1) free((void*) jobs[t].in);
		free(jobs[t].out);
	}
	pthread_barrier_destroy(&go);
	free(jobs); free(tids);
	if (seconds) *seconds = t1 - t0;
	if (compressed) *compressed = total;
	dlclose(h);
	return rc;
}
