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
#include "jdb_host.h"
#include "jdb_internal.h"

#define POISON        0xDEADBEEFu
#define INQ_BYTES     ((size_t) 32 << 20)     /* compressed bytes per launch   */
#define OUT_BYTES     ((size_t) 64 << 20)     /* staging for host targets      */

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
	uint32 used;
	uint32 done;            /* device reported the end of the stream */

	jdb_stream stream;
	jdb_dbuf   inq;         /* unconsumed compressed bytes: [inqoff, inqoff+inqlen) */
	size_t     inqoff;
	size_t     inqlen;
	jdb_dbuf   outbuf;

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
}

/* internal hooks for zstrm.c (hidden visibility) */
void
jdb_inflator_set_checks(TInflator* state, int which)
{
	PRVT->checks = which;
}

int
jdb_inflator_get_checks(TInflator* state, uint32* crc, uint32* adler)
{
	uint32_t* h = (uint32_t*) (PRVT->pinned + 208);

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
	size_t room;

	if (avail == 0) {
		return 0;
	}
	if (jdb_dbuf_reserve(&PRVT->inq, INQ_BYTES) != 0) {
		/* reserve() drops the old contents: only legal while the queue is empty */
		return -1;
	}
	if (PRVT->inqlen == 0) {
		PRVT->inqoff = 0;
	}
	if (PRVT->inqoff && PRVT->inqoff + PRVT->inqlen + avail > INQ_BYTES) {
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
	room = INQ_BYTES - (PRVT->inqoff + PRVT->inqlen);
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
	if (jdb_rt_init() != JDB_OK) {
		poison(PRVT, INFLT_EBADSTATE);
		return INFLT_ERROR;
	}
	PRVT->used = 1;
	tgt_on_device = jdb_ptr_is_device(PBLC->target);

	for (;;) {
		jdb_inflate_item* item = H_ITEM(PRVT);
		jdb_inflate_result* res = H_RESULT(PRVT);
		uint8* dst;
		size_t cap;
		size_t absorbed_all;

		if (absorb(PRVT) != 0) {
			poison(PRVT, INFLT_EOOM);
			return INFLT_ERROR;
		}
		absorbed_all = (PBLC->source == PBLC->send);

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
		}
		if (jdb_copy_async(D_ITEM(PRVT), item, sizeof(*item), PRVT->stream) != JDB_OK ||
		    jdb_inflate_batch(PRVT->inq.ptr, dst, D_ITEM(PRVT), D_RESULT(PRVT), PRVT->dstate, 1,
		                      JDB_FMT_RAW, (uint32_t) (PBLC->finalinput && absorbed_all),
		                      D_COUNTER(PRVT), PRVT->stream) != JDB_OK ||
		    jdb_copy_async(res, D_RESULT(PRVT), sizeof(*res), PRVT->stream) != JDB_OK ||
		    jdb_stream_sync(PRVT->stream) != JDB_OK) {
			poison(PRVT, INFLT_EBADSTATE);
			return INFLT_ERROR;
		}

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
