/*
 * inflator.c -- host side of the raw DEFLATE decoder: the public TInflator
 * state machine of the reference (src/inflator.c:156-260 create/reset/destroy,
 * :729-903 validate + inflator_inflate, :905-925 setdctnr) driving the warp
 * decoder of csrc/device/inflate.cu.
 *
 * Division of labour: this file keeps the API contract (status / error codes,
 * sticky poison, cursor updates, usage validation) and moves bytes; every bit
 * of DEFLATE decoding happens on the GPU.  Compressed input is absorbed into a
 * device-resident queue (so a block header or symbol cut by the end of a
 * source window is simply replayed when more input arrives), output is decoded
 * straight into the caller's buffer when that is device memory and through a
 * device staging buffer otherwise.
 */
#include <jdeflate/inflator.h>
#include <jdeflate/b200.h>
#include <string.h>
#include <stddef.h>
#include <stdlib.h>
#include <stdio.h>
#include "jdb_host.h"
#include "jdb_internal.h"

#define POISON        0xDEADBEEFu
#define INQ_BYTES     ((size_t) 32 << 20)     /* compressed bytes per launch (grows for large windows) */
#define INQ_MAX       ((size_t) 2048 << 20)   /* one parallel step should see all the chunks of a window: two steps of half a wave each take twice as long */
#define OUT_BYTES     ((size_t) 64 << 20)     /* staging for host targets      */

/* chunk-parallel decode of one stream (streams cut by sync markers, i.e. ours) */
#define PAR_MAXC      65536u                  /* marker candidates per step    */
#define PAR_MIN_BYTES ((size_t) 4 << 20)      /* queued input that makes a step worthwhile */
#define PAR_FIRST_BYTES ((size_t) 32 << 20)   /* ... for the first step of a caller that reads ahead (zstrm) */
#define PAR_MAXFRAG   4096u                   /* chunks decoded per step (slots of par_stride bytes each) */
#define PAR_STRIDE    ((size_t) 1 << 20)      /* first guess of the slot size: the encoder's largest default chunk */
#define PAR_STRIDE_MAX ((size_t) 16 << 20)
#define PAR_SCRATCH_MAX ((size_t) 8 << 30)    /* slots of one step: fewer chunks per step rather than more memory */

struct TINFLTPblc {
	uint32 state;
	uint32 error;
	uint32 flags;
	uint32 finalinput;
	uint32 status;
	const uint8* source;
	const uint8* sbgn;
	const uint8* send;
	uint8* target;
	uint8* tbgn;
	uint8* tend;
};

struct TINFLTPrvt {
	struct TINFLTPblc public;

	const TAllocator* allctr;
	int    device;          /* the GPU this instance lives on (default device at creation) */
	uint32 used;
	uint32 done;            /* device reported the end of the stream */

	jdb_stream stream;
	jdb_dbuf   inq;         /* unconsumed compressed bytes: [inqoff, inqoff+inqlen) */
	size_t     inqoff;
	size_t     inqlen;
	size_t     inqcap;

	/* chunk-parallel path: usable while the decoder sits at a block boundary with
	 * nothing buffered (start of the stream, or right after a parallel step) */
	uint32 par_ok;
	uint64 par_total;       /* bytes produced so far (= device total_out) */
	jdb_dbuf   par_dev;     /* device: candidate ends, items, results */
	jdb_dbuf   par_scratch; /* device: one slot per candidate chunk */
	size_t     par_stride;
	uint8*     par_host;    /* pinned mirror of the same */
	/* chunks decoded by the last step that the target has not taken yet: they stay in
	 * their slots and are handed out as the caller makes room */
	struct jdb_par_frag { uint32_t slot; uint32_t produced; }* par_frag;
	uint32 par_nfrag;
	uint32 par_cur;         /* next fragment to deliver ... */
	uint32 par_off;         /* ... and how much of it is gone already */
	uint32 par_fin;         /* the last fragment ends the stream (BFINAL marker) */
	uint32 par_starved;     /* the step stopped because the input ended inside a chunk */
	size_t par_seen;        /* queued bytes that step has seen */
	size_t readahead;       /* zstrm: the caller signals the end of the input with `final`, so a chain under way may
	                         * wait for this much input before the next step (a step takes as long as its slowest chunk) */
	size_t par_want;        /* queued bytes the next step of a chain waits for */
	size_t leftover;        /* bytes after the end of the stream that came from earlier source windows: they
	                         * sit at inq[inqoff ..) until zstrm takes them back or the next reset */
	jdb_dbuf   outbuf;
	int        narrow;      /* JDB200_INFLATE_NARROW=1 (read at create): one warp instead of a CTA for the
	                         * sequential decoder -- the round-1 shape, kept for measurements */

	jdb_inflate_state* dstate;
	/* pinned: item + result; device mirrors */
	uint8* pinned;
	uint8* dscratch;

	/* checksums of the produced bytes for zstrm (device side, running) */
	int       checks;       /* JDB_CK_* mask */
	uint32_t* dchecks;      /* device: [0] crc register, [1] adler */
	jdb_dbuf  ckwork;
};

typedef char jdb_inflator_layout_check[(sizeof(struct TInflator) == sizeof(struct TINFLTPblc)) ? 1 : -1];

#define PRVT ((struct TINFLTPrvt*) state)
#define PBLC ((struct TINFLTPblc*) state)

#define H_ITEM(P)   ((jdb_inflate_item*)   ((P)->pinned))
#define H_RESULT(P) ((jdb_inflate_result*) ((P)->pinned + 64))
#define D_ITEM(P)   ((jdb_inflate_item*)   ((P)->dscratch))
#define D_RESULT(P) ((jdb_inflate_result*) ((P)->dscratch + 64))
#define D_COUNTER(P) ((uint32_t*)          ((P)->dscratch + 128))

static void
poison(struct TINFLTPrvt* state, uint32 error)
{
	if (PBLC->error == 0) {
		PBLC->error = error;
	}
	PBLC->state = POISON;
}

TInflator*
inflator_create(uintxx flags, const TAllocator* allctr)
{
	struct TInflator* state;

	if (allctr == NULL) {
		allctr = ctb_getdefaultallocator();
	}
	/* no CUDA device, no decoder: there is no CPU path */
	if (jdb_rt_init() != JDB_OK) {
		return NULL;
	}
	state = allctr->request(sizeof(struct TINFLTPrvt), allctr->user);
	if (state == NULL) {
		return NULL;
	}
	memset(state, 0, sizeof(struct TINFLTPrvt));
	PRVT->allctr = allctr;
	PRVT->device = jdb_rt_current_device();
	{
		const char* e = getenv("JDB200_INFLATE_NARROW");
		PRVT->narrow = e != NULL && e[0] == '1';
	}

	if (jdb_stream_create(&PRVT->stream) != JDB_OK) {
		goto L_FAIL;
	}
	PRVT->dstate = jdb_dev_alloc(sizeof(jdb_inflate_state));
	PRVT->dscratch = jdb_dev_alloc(256);
	PRVT->pinned = jdb_pinned_alloc(256);
	PRVT->dchecks = jdb_dev_alloc(64);
	if (PRVT->dstate == NULL || PRVT->dscratch == NULL || PRVT->pinned == NULL || PRVT->dchecks == NULL) {
		goto L_FAIL;
	}

	inflator_reset(state);
	if (PBLC->error) {
		goto L_FAIL;
	}
	PBLC->flags = (uint32) flags;
	return state;

L_FAIL:
	inflator_destroy(state);
	return NULL;
}

void
inflator_reset(TInflator* state)
{
	CTB_ASSERT(state);

	jdb_rt_use_device(PRVT->device);
	PBLC->state = 0;
	PBLC->error = 0;
	PBLC->finalinput = 0;
	PBLC->status = 0;
	PBLC->source = PBLC->sbgn = PBLC->send = NULL;
	PBLC->target = PBLC->tbgn = PBLC->tend = NULL;

	PRVT->used = 0;
	PRVT->done = 0;
	PRVT->inqoff = 0;
	PRVT->inqlen = 0;
	PRVT->leftover = 0;
	PRVT->par_ok = 1;
	PRVT->par_total = 0;
	PRVT->par_nfrag = PRVT->par_cur = PRVT->par_off = PRVT->par_fin = PRVT->par_starved = 0;
	PRVT->par_seen = 0;
	PRVT->par_want = PAR_MIN_BYTES;
	if (PRVT->par_stride == 0) {
		PRVT->par_stride = PAR_STRIDE;
	}

	{
		uint32_t* init = (uint32_t*) (PRVT->pinned + 208);
		init[0] = 0xffffffffu;
		init[1] = 1u;
		jdb_copy_async(PRVT->dchecks, init, 8, PRVT->stream);
	}
	/* a zeroed state block is "at a block header, no history" */
	if (jdb_memset_async(PRVT->dstate, 0, offsetof(jdb_inflate_state, lit), PRVT->stream) != JDB_OK) {
		PBLC->error = INFLT_EOOM;
		PBLC->state = POISON;
	}
}

void
inflator_destroy(TInflator* state)
{
	const TAllocator* a;

	if (state == NULL) {
		return;
	}
	if (PRVT->stream) {
		jdb_stream_sync(PRVT->stream);
	}
	jdb_dbuf_release(&PRVT->inq);
	jdb_dbuf_release(&PRVT->outbuf);
	jdb_dbuf_release(&PRVT->par_dev);
	jdb_dbuf_release(&PRVT->par_scratch);
	jdb_pinned_free(PRVT->par_host);
	free(PRVT->par_frag);
	jdb_dbuf_release(&PRVT->ckwork);
	jdb_dev_free(PRVT->dchecks);
	jdb_dev_free(PRVT->dstate);
	jdb_dev_free(PRVT->dscratch);
	jdb_pinned_free(PRVT->pinned);
	jdb_stream_destroy(PRVT->stream);

	a = PRVT->allctr;
	a->dispose(state, sizeof(struct TINFLTPrvt), a->user);
}

void
inflator_setdctnr(TInflator* state, const uint8* dict, uintxx size)
{
	uint64_t v[2];
	CTB_ASSERT(state && dict && size);

	jdb_rt_use_device(PRVT->device);
	/* only before the first inflate call (src/inflator.c:905-913) */
	if (PRVT->used) {
		poison(PRVT, INFLT_EINCORRECTUSE);
		return;
	}
	if (size > JDB_INFLATE_HISTORY) {
		dict = (dict + size) - JDB_INFLATE_HISTORY;
		size = JDB_INFLATE_HISTORY;
	}
	/* the dictionary becomes the first `size` bytes of the history ring */
	v[0] = size;        /* total_out */
	v[1] = size;        /* hist_avail */
	memcpy(PRVT->pinned + 192, v, sizeof(v));
	if (jdb_copy_async(PRVT->dstate->history, dict, size, PRVT->stream) != JDB_OK ||
	    jdb_copy_async(&PRVT->dstate->total_out, PRVT->pinned + 192, 16, PRVT->stream) != JDB_OK ||
	    jdb_stream_sync(PRVT->stream) != JDB_OK) {
		poison(PRVT, INFLT_EOOM);
		return;
	}
	PRVT->used = 1;
	PRVT->par_ok = 0;       /* history in front of the stream: sequential decoder only */
}

/* internal hooks for zstrm.c (hidden visibility) */
void
jdb_inflator_set_checks(TInflator* state, int which)
{
	PRVT->checks = which;
}

void
jdb_inflator_set_readahead(TInflator* state, size_t bytes)
{
	PRVT->readahead = bytes;
	if (bytes && PRVT->par_total == 0) {
		PRVT->par_want = bytes < PAR_FIRST_BYTES ? bytes : PAR_FIRST_BYTES;
	}
}

/* zstrm: the memory of the last source window is gone (its read buffer was reused): nothing
 * can be handed back through it any more */
void
jdb_inflator_drop_window(TInflator* state)
{
	PBLC->sbgn = PBLC->source = PBLC->send;
}

/* Bytes behind the end of the stream that the public source window cannot give back because
 * they were queued from earlier windows (zstrm's read-ahead): zstrm takes them back from here. */
size_t
jdb_inflator_leftover(TInflator* state)
{
	return PRVT->done ? PRVT->leftover : 0;
}

int
jdb_inflator_take_leftover(TInflator* state, uint8* dst)
{
	jdb_rt_use_device(PRVT->device);
	if (PRVT->leftover == 0) {
		return 0;
	}
	if (jdb_copy_async(dst, PRVT->inq.ptr + PRVT->inqoff, PRVT->leftover, PRVT->stream) != JDB_OK ||
	    jdb_stream_sync(PRVT->stream) != JDB_OK) {
		return -1;
	}
	PRVT->leftover = 0;
	return 0;
}

int
jdb_inflator_get_checks(TInflator* state, uint32* crc, uint32* adler)
{
	uint32_t* h = (uint32_t*) (PRVT->pinned + 208);

	jdb_rt_use_device(PRVT->device);
	if (jdb_copy_async(h, PRVT->dchecks, 8, PRVT->stream) != JDB_OK ||
	    jdb_stream_sync(PRVT->stream) != JDB_OK) {
		return -1;
	}
	*crc = h[0];
	*adler = h[1];
	return 0;
}

/* usage rules of the reference, src/inflator.c:729-762 */
static int
validate(struct TINFLTPrvt* state)
{
	if (PBLC->source == NULL || PBLC->target == NULL) {
		PBLC->error = INFLT_EINCORRECTUSE;
		return 0;
	}
	if (PBLC->status == INFLT_SRCEXHSTD) {
		if (PBLC->source == PBLC->send && PBLC->finalinput == 0) {
			PBLC->error = INFLT_EINCORRECTUSE;
			return 0;
		}
	}
	if (PBLC->status == INFLT_TGTEXHSTD) {
		if (PBLC->target == PBLC->tend) {
			PBLC->error = INFLT_EINCORRECTUSE;
			return 0;
		}
	}
	return 1;
}

/* move as much of the source window as fits into the device queue */
static int
absorb(struct TINFLTPrvt* state)
{
	size_t avail = (size_t) (PBLC->send - PBLC->source);
	size_t room, want;

	if (avail == 0) {
		return 0;
	}
	/* Large windows get a large queue so that a parallel step sees many chunks at once; so
	 * does a chain of small windows that gathers input for its next step (par_want). */
	want = PRVT->inqcap ? PRVT->inqcap : INQ_BYTES;
	while (want < PRVT->inqlen + avail && want < INQ_MAX) {
		want *= 2;
	}
	if (want > PRVT->inqcap || PRVT->inq.ptr == NULL) {
		if (PRVT->inqlen == 0) {
			if (jdb_dbuf_reserve(&PRVT->inq, want) != 0) {
				return -1;
			}
			PRVT->inqcap = want;
			PRVT->inqoff = 0;
		}
		else {
			/* the queue holds bytes: they move to the front of a larger one */
			jdb_dbuf bigger = { NULL, 0 };
			if (jdb_dbuf_reserve(&bigger, want) == 0) {
				if (jdb_copy_async(bigger.ptr, PRVT->inq.ptr + PRVT->inqoff, PRVT->inqlen, PRVT->stream) != JDB_OK ||
				    jdb_stream_sync(PRVT->stream) != JDB_OK) {
					jdb_dbuf_release(&bigger);
					return -1;
				}
				jdb_dbuf_release(&PRVT->inq);
				PRVT->inq = bigger;
				PRVT->inqcap = want;
				PRVT->inqoff = 0;
			}
			/* else: no memory for a larger queue -- the caller works with the one there is */
		}
	}
	if (PRVT->inq.ptr == NULL) {
		return -1;
	}
	if (PRVT->inqoff && PRVT->inqoff + PRVT->inqlen + avail > PRVT->inqcap) {
		/* compact: the tail is small (a cut header or symbol) unless the target
		 * was the limiting side; bounce through the staging buffer */
		if (jdb_dbuf_reserve(&PRVT->outbuf, OUT_BYTES) != 0) {
			return -1;
		}
		if (PRVT->inqlen <= OUT_BYTES) {
			jdb_copy_async(PRVT->outbuf.ptr, PRVT->inq.ptr + PRVT->inqoff, PRVT->inqlen, PRVT->stream);
			jdb_copy_async(PRVT->inq.ptr, PRVT->outbuf.ptr, PRVT->inqlen, PRVT->stream);
			PRVT->inqoff = 0;
		}
	}
	room = PRVT->inqcap - (PRVT->inqoff + PRVT->inqlen);
	if (avail > room) {
		avail = room;
	}
	if (avail) {
		if (jdb_copy_async(PRVT->inq.ptr + PRVT->inqoff + PRVT->inqlen, PBLC->source, avail, PRVT->stream) != JDB_OK) {
			return -1;
		}
		PRVT->inqlen += avail;
		PBLC->source += avail;
	}
	return 0;
}

static int
cmp_u32(const void* a, const void* b)
{
	uint32_t x = *(const uint32_t*) a, y = *(const uint32_t*) b;
	return x < y ? -1 : x > y;
}

/* index of `v` in the sorted array a[0..n), or n */
static uint32_t
find_u32(const uint32_t* a, uint32_t n, uint32_t v)
{
	uint32_t lo = 0, hi = n;
	while (lo < hi) {
		uint32_t mid = lo + (hi - lo) / 2;
		if (a[mid] < v) lo = mid + 1; else hi = mid;
	}
	return (lo < n && a[lo] == v) ? lo : n;
}

/*
 * One chunk-parallel step over the queued input:
 *   1. list every "00 00 FF FF" in the queue (marker_scan_kernel): candidate
 *      chunk ends;
 *   2. decode from the queue start and from every candidate, one warp each, into
 *      its own slot of a scratch buffer, stopping at the first empty stored
 *      block: a real chunk reports where it ends and how many bytes it decoded
 *      to; false candidates just fail or lead nowhere;
 *   3. follow the chain start -> marker -> marker ... : the chunks of the chain are
 *      the next bytes of the output.  They stay in their slots (par_frag) and
 *      deliver_pending() hands them to the caller as the target has room.
 * Nothing is assumed about who wrote the stream: a stream without such markers
 * yields an empty chain and the sequential decoder takes over.
 * Returns 1 when it decoded something, 0 when it does not apply, 2 to be called
 * again (slot size raised), -1 on failure.
 */
static int
parallel_step(struct TINFLTPrvt* state)
{
	const size_t u32_bytes = ((size_t) PAR_MAXC + 1) * 4, rec_bytes = ((size_t) PAR_MAXC + 1) * 32;
	const size_t total_bytes = 256 + 2 * u32_bytes + 2 * rec_bytes;
	uint32_t* h_count;
	uint32_t* h_ends;
	uint32_t* h_starts;
	jdb_inflate_item* h_items;
	jdb_inflate_result* h_res;
	uint8_t* d;
	uint32_t* d_count;
	uint32_t* d_ends;
	jdb_inflate_item* d_items;
	jdb_inflate_result* d_res;
	const size_t n = PRVT->inqlen;
	const size_t pad = PRVT->inqoff & 3u;
	uint32_t nc, ns, i, k, nfrag, maxfrag;
	uint64_t total, used;
	int fin = 0;

	PRVT->par_nfrag = PRVT->par_cur = PRVT->par_off = PRVT->par_fin = PRVT->par_starved = 0;
	if (n < 5 || n + pad > 0xfffffff0u) {
		return 0;
	}
	if (PRVT->par_host == NULL) {
		PRVT->par_host = jdb_pinned_alloc(total_bytes);
		if (PRVT->par_host == NULL) {
			return -1;
		}
	}
	if (PRVT->par_frag == NULL) {
		PRVT->par_frag = malloc((size_t) PAR_MAXFRAG * sizeof(*PRVT->par_frag));
		if (PRVT->par_frag == NULL) {
			return -1;
		}
	}
	if (jdb_dbuf_reserve(&PRVT->par_dev, total_bytes) != 0) {
		return -1;
	}
	h_count = (uint32_t*) PRVT->par_host;
	h_ends = (uint32_t*) (PRVT->par_host + 256);
	h_starts = (uint32_t*) (PRVT->par_host + 256 + u32_bytes);
	h_items = (jdb_inflate_item*) (PRVT->par_host + 256 + 2 * u32_bytes);
	h_res = (jdb_inflate_result*) (PRVT->par_host + 256 + 2 * u32_bytes + rec_bytes);
	d = PRVT->par_dev.ptr;
	d_count = (uint32_t*) d;
	d_ends = (uint32_t*) (d + 256);
	d_items = (jdb_inflate_item*) (d + 256 + 2 * u32_bytes);
	d_res = (jdb_inflate_result*) (d + 256 + 2 * u32_bytes + rec_bytes);

	/* 1. candidates (the scan wants a 4-byte aligned start: begin up to 3 bytes early) */
	if (jdb_marker_scan(PRVT->inq.ptr + PRVT->inqoff - pad, n + pad, d_ends, PAR_MAXC, d_count, PRVT->stream) != JDB_OK ||
	    jdb_copy_async(h_count, d_count, 4, PRVT->stream) != JDB_OK ||
	    jdb_stream_sync(PRVT->stream) != JDB_OK) {
		return -1;
	}
	nc = h_count[0];
	if (nc == 0 || nc > PAR_MAXC) {
		return 0;
	}
	if (jdb_copy_async(h_ends, d_ends, (size_t) nc * 4, PRVT->stream) != JDB_OK ||
	    jdb_stream_sync(PRVT->stream) != JDB_OK) {
		return -1;
	}
	qsort(h_ends, nc, 4, cmp_u32);

	/* 2. decode from the start and from every candidate that has input after it, each
	 * into its own slot of a scratch buffer, stopping at the first marker it reads: one
	 * pass gives both the sizes and the bytes (a slot too small for its chunk just ends
	 * the chain there).  Offsets from here on are relative to the first queued byte.
	 * The slots of one step stay within a memory budget: fewer chunks per step then. */
	maxfrag = (uint32_t) (PAR_SCRATCH_MAX / PRVT->par_stride);
	if (maxfrag > PAR_MAXFRAG) {
		maxfrag = PAR_MAXFRAG;
	}
	ns = 0;
	h_starts[ns++] = 0;
	for (i = 0; i < nc && ns < maxfrag; i++) {
		if (h_ends[i] > pad && h_ends[i] - pad < n) {
			h_starts[ns++] = (uint32_t) (h_ends[i] - pad);
		}
	}
	if (jdb_dbuf_reserve(&PRVT->par_scratch, (size_t) ns * PRVT->par_stride) != 0) {
		return 0;                           /* no memory for the slots: sequential decoder */
	}
	for (i = 0; i < ns; i++) {
		h_items[i].src_off = PRVT->inqoff + h_starts[i];
		h_items[i].dst_off = (uint64_t) i * PRVT->par_stride;
		h_items[i].src_len = n - h_starts[i];
		h_items[i].dst_cap = PRVT->par_stride;
	}
	if (jdb_copy_async(d_items, h_items, (size_t) ns * sizeof(*h_items), PRVT->stream) != JDB_OK ||
	    jdb_inflate_chunks(PRVT->inq.ptr, PRVT->par_scratch.ptr, d_items, d_res, ns, D_COUNTER(PRVT), PRVT->stream) != JDB_OK ||
	    jdb_copy_async(h_res, d_res, (size_t) ns * sizeof(*h_res), PRVT->stream) != JDB_OK ||
	    jdb_stream_sync(PRVT->stream) != JDB_OK) {
		return -1;
	}
	if (h_res[0].status == INFLT_TGTEXHSTD && PRVT->par_stride < PAR_STRIDE_MAX) {
		/* the first chunk does not fit a slot: larger slots next time */
		PRVT->par_stride *= 4;
		return 2;
	}

	/* 3. the chain of real chunks */
	total = 0;
	used = 0;
	nfrag = 0;
	k = 0;
	while (k < ns) {
		const jdb_inflate_result r = h_res[k];
		const uint32_t start = h_starts[k];
		uint64_t next;
		if (r.status != JDB_INF_ST_MARKER || r.consumed == 0 || r.consumed > n - start) {
			/* the input ended inside this chunk: more input continues the chain */
			PRVT->par_starved = (r.status == INFLT_SRCEXHSTD);
			break;
		}
		PRVT->par_frag[nfrag].slot = k;
		PRVT->par_frag[nfrag].produced = (uint32_t) r.produced;
		nfrag++;
		total += r.produced;
		used = start + r.consumed;
		if (r.error) {
			fin = 1;                    /* the marker carried BFINAL: end of the stream */
			break;
		}
		next = start + r.consumed;
		if (next >= n) {
			PRVT->par_starved = 1;      /* the queue ends with the marker */
			break;
		}
		i = find_u32(h_starts + k + 1, ns - (k + 1), (uint32_t) next);
		if (i == ns - (k + 1)) {
			break;                      /* more chunks than slots: the next step goes on from here */
		}
		k = k + 1 + i;
	}
	if (nfrag == 0) {
		/* not one whole chunk: an unfinished first chunk of a chain that was already
		 * under way waits for more input; anything else is not ours to decode */
		PRVT->par_starved = PRVT->par_starved && PRVT->par_total != 0;
		PRVT->par_seen = n;
		return 0;
	}

	/* the sequential decoder may have to go on from here: leave its state at the
	 * block boundary after the last chunk, with the newest 32 KiB as history
	 * (gathered from the slots, newest chunk first) */
	{
		const uint64_t newtotal = PRVT->par_total + total;
		uint64_t left = total < JDB_INFLATE_HISTORY ? total : JDB_INFLATE_HISTORY;
		uint64_t end = newtotal;                             /* absolute position after the bytes still to copy */
		uint64_t* hw = (uint64_t*) (PRVT->pinned + 128);
		uint32_t* hw32 = (uint32_t*) (PRVT->pinned + 128 + 24);
		uint32_t f = nfrag;
		while (left && f) {
			const struct jdb_par_frag* fr = &PRVT->par_frag[--f];
			const uint8_t* slot = PRVT->par_scratch.ptr + (size_t) fr->slot * PRVT->par_stride;
			uint64_t take = fr->produced < left ? fr->produced : left;
			uint64_t first = end - take;                     /* absolute position of the oldest byte of this piece */
			const uint8_t* src = slot + (fr->produced - take);
			uint64_t ring = first & (JDB_INFLATE_HISTORY - 1);
			uint64_t head = JDB_INFLATE_HISTORY - ring;
			if (head > take) {
				head = take;
			}
			if (jdb_copy_async(PRVT->dstate->history + ring, src, (size_t) head, PRVT->stream) != JDB_OK ||
			    (take > head &&
			     jdb_copy_async(PRVT->dstate->history, src + head, (size_t) (take - head), PRVT->stream) != JDB_OK)) {
				return -1;
			}
			end = first;
			left -= take;
		}
		hw[0] = 0;                                           /* bitbuf */
		hw[1] = newtotal;                                    /* total_out */
		hw[2] = newtotal < JDB_INFLATE_HISTORY ? newtotal : JDB_INFLATE_HISTORY;   /* hist_avail */
		hw32[0] = 0;                                         /* bitcnt */
		hw32[1] = JDB_INF_HEADER;                            /* phase */
		hw32[2] = (uint32_t) fin;                            /* lastblock */
		hw32[3] = 0;                                         /* stored_left */
		hw32[4] = 0;                                         /* pend_len */
		hw32[5] = 0;                                         /* pend_dist */
		if (jdb_copy_async(PRVT->dstate, hw, offsetof(jdb_inflate_state, lit), PRVT->stream) != JDB_OK ||
		    jdb_stream_sync(PRVT->stream) != JDB_OK) {
			return -1;
		}
		PRVT->par_total = newtotal;
	}
	PRVT->inqoff += (size_t) used;
	PRVT->inqlen -= (size_t) used;
	PRVT->par_seen = PRVT->inqlen;
	/* the next step of this chain: four times the input of this one, up to the read-ahead */
	PRVT->par_want = (size_t) used * 4 > PAR_MIN_BYTES ? (size_t) used * 4 : PAR_MIN_BYTES;
	if (PRVT->readahead && PRVT->par_want > PRVT->readahead) {
		PRVT->par_want = PRVT->readahead;
	}
	PRVT->par_nfrag = nfrag;
	PRVT->par_fin = (uint32) fin;
	if (getenv("JDB200_TRACE")) {
		fprintf(stderr, "parallel_step: queued %zu starts %u frags %u used %llu out %llu starved %u want %zu cap %zu\n",
		        n, ns, nfrag, (unsigned long long) used, (unsigned long long) total, PRVT->par_starved, PRVT->par_want, PRVT->inqcap);
	}
	return 1;
}

/* 1 when the bytes queued since the last step hold a marker candidate, 0 when not, -1 on failure */
static int
new_marker(struct TINFLTPrvt* state)
{
	const size_t from = PRVT->par_seen > 3 ? PRVT->par_seen - 3 : 0;       /* a marker may straddle old and new bytes */
	const size_t off = PRVT->inqoff + from;
	const size_t pad = off & 3u;
	uint32_t* h_count = (uint32_t*) PRVT->par_host;
	uint8_t* d = PRVT->par_dev.ptr;

	if (PRVT->par_host == NULL || d == NULL) {
		return 1;
	}
	if (jdb_marker_scan(PRVT->inq.ptr + off - pad, PRVT->inqlen - from + pad, (uint32_t*) (d + 256), PAR_MAXC,
	                    (uint32_t*) d, PRVT->stream) != JDB_OK ||
	    jdb_copy_async(h_count, d, 4, PRVT->stream) != JDB_OK ||
	    jdb_stream_sync(PRVT->stream) != JDB_OK) {
		return -1;
	}
	PRVT->par_seen = PRVT->inqlen;
	return h_count[0] != 0;
}

/*
 * Hand decoded chunks that wait in their slots to the caller, as far as the target has
 * room (device target: device copy; host target: straight from the slot).  The running
 * checksums follow the bytes delivered.  Returns 0, or -1 on failure.
 */
static int
deliver_pending(struct TINFLTPrvt* state)
{
	int copied = 0;

	while (PRVT->par_cur < PRVT->par_nfrag && PBLC->target != PBLC->tend) {
		const struct jdb_par_frag* fr = &PRVT->par_frag[PRVT->par_cur];
		const uint8_t* src = PRVT->par_scratch.ptr + (size_t) fr->slot * PRVT->par_stride + PRVT->par_off;
		size_t left = fr->produced - PRVT->par_off;
		size_t room = (size_t) (PBLC->tend - PBLC->target);
		size_t take = left < room ? left : room;

		if (take) {
			if (PRVT->checks) {
				if (jdb_dbuf_reserve(&PRVT->ckwork, jdb_checksum_workspace_bytes()) != 0 ||
				    jdb_checksum(src, take, PRVT->checks, PRVT->dchecks, PRVT->dchecks + 1,
				                 PRVT->ckwork.ptr, PRVT->stream) != JDB_OK) {
					return -1;
				}
			}
			if (jdb_copy_async(PBLC->target, src, take, PRVT->stream) != JDB_OK) {
				return -1;
			}
			copied = 1;
			PBLC->target += take;
		}
		PRVT->par_off += (uint32) take;
		if (PRVT->par_off == fr->produced) {
			PRVT->par_cur++;
			PRVT->par_off = 0;
		}
	}
	if (copied && jdb_stream_sync(PRVT->stream) != JDB_OK) {
		return -1;
	}
	return 0;
}

eINFLTResult
inflator_inflate(TInflator* state, uint32 final)
{
	int tgt_on_device;
	CTB_ASSERT(state);

	if (PBLC->state == POISON) {
		return INFLT_ERROR;
	}
	if (PBLC->finalinput == 0 && final) {
		PBLC->finalinput = 1;
	}
	if (validate(PRVT) == 0) {
		PBLC->state = POISON;
		return INFLT_ERROR;
	}
	if (jdb_rt_use_device(PRVT->device) != JDB_OK) {
		poison(PRVT, INFLT_EBADSTATE);
		return INFLT_ERROR;
	}
	PRVT->used = 1;
	tgt_on_device = jdb_ptr_is_device(PBLC->target);

	for (;;) {
		jdb_inflate_item* item = H_ITEM(PRVT);
		jdb_inflate_result* res = H_RESULT(PRVT);
		uint8* dst = NULL;
		size_t cap = 0;
		size_t absorbed_all = 0;

		/* chunks of an earlier parallel step that the target had no room for */
		if (PRVT->par_cur < PRVT->par_nfrag) {
			if (deliver_pending(PRVT) != 0) {
				poison(PRVT, INFLT_EBADSTATE);
				return INFLT_ERROR;
			}
			if (PRVT->par_cur < PRVT->par_nfrag) {
				return (eINFLTResult) (PBLC->status = INFLT_TGTEXHSTD);
			}
			if (PRVT->par_fin) {
				res->status = INFLT_OK;
				res->error = 0;
				res->consumed = 0;
				res->produced = 0;
				goto L_DELIVER;
			}
		}

		if (absorb(PRVT) != 0) {
			poison(PRVT, INFLT_EOOM);
			return INFLT_ERROR;
		}
		absorbed_all = (PBLC->source == PBLC->send);

		/* a stream cut into chunks by sync markers (ours) is decoded chunk-parallel
		 * while the decoder sits at a block boundary */
		if (PRVT->par_ok) {
			/* with read-ahead (zstrm) the first step, too, waits for a good deal of input: a step
			 * costs the time of its slowest chunk however few chunks it holds */
			const int chain = PRVT->readahead && (PRVT->par_total == 0 || PRVT->par_starved);
			int step = PRVT->inqlen >= (chain ? PRVT->par_want : PAR_MIN_BYTES) ||
			           (PBLC->finalinput && absorbed_all && PRVT->inqlen >= ((size_t) 256 << 10));
			if (!step && chain && !PBLC->finalinput) {
				/* zstrm feeds a chain of chunks and promises `final` at the end of its input:
				 * gather more of it, a step over many chunks takes no longer than one over few */
				if (absorbed_all) {
					/* the caller may reuse its window the moment this returns: the copy into
					 * the queue has to be over */
					if (jdb_stream_sync(PRVT->stream) != JDB_OK) {
						poison(PRVT, INFLT_EBADSTATE);
						return INFLT_ERROR;
					}
					return (eINFLTResult) (PBLC->status = INFLT_SRCEXHSTD);
				}
				step = 1;               /* the queue is full and cannot grow: work with what it holds */
			}
			if (!step && PRVT->par_starved) {
				/* a streaming caller with small windows: the last step stopped because the
				 * queue ended inside a chunk (or right after one).  Nothing new: ask for
				 * input.  New bytes with a marker among them: the chain may go on.  New
				 * bytes without one: cannot tell an unfinished chunk from the unmarked end
				 * of somebody else's stream -- the sequential decoder decides. */
				if (PRVT->inqlen == PRVT->par_seen) {
					if (absorbed_all && !PBLC->finalinput) {
						if (jdb_stream_sync(PRVT->stream) != JDB_OK) {
							poison(PRVT, INFLT_EBADSTATE);
							return INFLT_ERROR;
						}
						return (eINFLTResult) (PBLC->status = INFLT_SRCEXHSTD);
					}
				}
				else {
					int found = new_marker(PRVT);
					if (found < 0) {
						poison(PRVT, INFLT_EBADSTATE);
						return INFLT_ERROR;
					}
					step = found;
				}
			}
			if (step) {
				int pr = parallel_step(PRVT);
				if (pr < 0) {
					poison(PRVT, INFLT_EBADSTATE);
					return INFLT_ERROR;
				}
				if (pr == 2) {
					continue;           /* the slots were too small for the chunks: again with larger ones */
				}
				if (pr > 0) {
					continue;           /* the chunks wait in their slots: the top of the loop delivers them */
				}
				if (PRVT->par_starved && absorbed_all && !PBLC->finalinput) {
					/* still the same unfinished chunk at the head of the queue */
					return (eINFLTResult) (PBLC->status = INFLT_SRCEXHSTD);
				}
			}
			/* no chain of chunks here: the sequential decoder owns the stream from now on */
		}
		PRVT->par_ok = 0;

		cap = (size_t) (PBLC->tend - PBLC->target);
		if (tgt_on_device) {
			dst = PBLC->target;
		}
		else {
			if (jdb_dbuf_reserve(&PRVT->outbuf, OUT_BYTES) != 0) {
				poison(PRVT, INFLT_EOOM);
				return INFLT_ERROR;
			}
			if (cap > OUT_BYTES) {
				cap = OUT_BYTES;
			}
			dst = PRVT->outbuf.ptr;
		}

		item->src_off = PRVT->inqoff;
		item->src_len = PRVT->inqlen;
		item->dst_off = 0;
		item->dst_cap = cap;
		if (PRVT->inq.ptr == NULL) {
			/* nothing was ever fed: still let the kernel report starvation */
			if (jdb_dbuf_reserve(&PRVT->inq, INQ_BYTES) != 0) {
				poison(PRVT, INFLT_EOOM);
				return INFLT_ERROR;
			}
			PRVT->inqcap = INQ_BYTES;
		}

		if (jdb_copy_async(D_ITEM(PRVT), item, sizeof(*item), PRVT->stream) != JDB_OK ||
		    (PRVT->narrow
		         ? jdb_inflate_batch(PRVT->inq.ptr, dst, D_ITEM(PRVT), D_RESULT(PRVT), PRVT->dstate, 1,
		                             JDB_FMT_RAW, (uint32_t) (PBLC->finalinput && absorbed_all),
		                             D_COUNTER(PRVT), PRVT->stream)
		         : jdb_inflate_wide(PRVT->inq.ptr, dst, D_ITEM(PRVT), D_RESULT(PRVT), PRVT->dstate, 1, JDB_FMT_RAW,
		                            (uint32_t) (PBLC->finalinput && absorbed_all), PRVT->stream)) != JDB_OK ||
		    jdb_copy_async(res, D_RESULT(PRVT), sizeof(*res), PRVT->stream) != JDB_OK ||
		    jdb_stream_sync(PRVT->stream) != JDB_OK) {
			poison(PRVT, INFLT_EBADSTATE);
			return INFLT_ERROR;
		}

L_DELIVER:
		/* account for what the device used */
		PRVT->inqoff += (size_t) res->consumed;
		PRVT->inqlen -= (size_t) res->consumed;
		if (res->produced) {
			if (PRVT->checks) {
				if (jdb_dbuf_reserve(&PRVT->ckwork, jdb_checksum_workspace_bytes()) != 0 ||
				    jdb_checksum(dst, (size_t) res->produced, PRVT->checks, PRVT->dchecks, PRVT->dchecks + 1,
				                 PRVT->ckwork.ptr, PRVT->stream) != JDB_OK) {
					poison(PRVT, INFLT_EOOM);
					return INFLT_ERROR;
				}
			}
			if (!tgt_on_device) {
				if (jdb_copy_async(PBLC->target, dst, (size_t) res->produced, PRVT->stream) != JDB_OK ||
				    jdb_stream_sync(PRVT->stream) != JDB_OK) {
					poison(PRVT, INFLT_EBADSTATE);
					return INFLT_ERROR;
				}
			}
			PBLC->target += res->produced;
		}

		switch (res->status) {
			case INFLT_OK: {
				/* bytes after the end of the stream go back to the caller
				 * (exact input accounting, DESIGN.md deviation 3) */
				size_t giveback = PRVT->inqlen;
				size_t window = (size_t) (PBLC->source - PBLC->sbgn);
				if (giveback > window) {
					giveback = window;
				}
				PBLC->source -= giveback;
				/* what came from earlier windows stays in the queue: jdb_inflator_leftover() */
				PRVT->leftover = PRVT->inqlen - giveback;
				PRVT->inqlen = 0;
				PRVT->done = 1;
				PBLC->state = POISON;
				return (eINFLTResult) (PBLC->status = INFLT_OK);
			}
			case INFLT_ERROR:
				PBLC->error = res->error ? res->error : INFLT_EBADSTATE;
				PBLC->state = POISON;
				return INFLT_ERROR;
			case INFLT_TGTEXHSTD:
				if (PBLC->target == PBLC->tend) {
					return (eINFLTResult) (PBLC->status = INFLT_TGTEXHSTD);
				}
				/* only the staging buffer was full: go again */
				continue;
			default: /* INFLT_SRCEXHSTD */
				if (!absorbed_all) {
					continue;
				}
				return (eINFLTResult) (PBLC->status = INFLT_SRCEXHSTD);
		}
	}
}
