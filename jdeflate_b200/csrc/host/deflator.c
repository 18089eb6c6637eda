/*
 * deflator.c -- host side of the raw DEFLATE encoder: the public TDeflator
 * contract of the reference (src/deflator.c:375-545 create/reset/destroy,
 * :663-786 validate + deflator_deflate, :241-263 level parameters) in front of
 * the chunk-parallel GPU pipeline of csrc/device/deflate.cu.
 *
 * The reference copies input into a 64/128 KiB sliding window and parses it
 * byte by byte; here input is gathered into a batch in HBM (or used in place
 * when the caller already hands over device memory), one pipeline launch
 * compresses the whole batch, and the compressed bytes are drained into the
 * caller's target window -- across as many calls as it takes, which is what
 * keeps the DEFLT_TGTEXHSTD / DEFLT_SRCEXHSTD streaming contract intact.
 *
 * Output framing: every batch consists of independent chunks, each closed by
 * the reference's sync marker (endstream, src/deflator.c:609-654); the last
 * chunk of a DEFLT_END call carries BFINAL = 1.  A DEFLT_FLUSH therefore ends
 * exactly like the reference's (00 00 FF FF) and the stream stays valid for
 * any inflater.
 */
#include <jdeflate/deflator.h>
#include <string.h>
#include <stdlib.h>
#include "jdb_host.h"
#include "jdb_internal.h"

#define POISON 0xDEADBEEFu

/* batch geometry (overridable through the environment for experiments) */
#define DEFAULT_BATCH   ((size_t) 256 << 20)
/* measured on text: +1.0 % size at 256 KiB, +0.4 % at 512 KiB, +0.1 % at 1 MiB vs the reference, same
 * encode speed; the chunk is also the unit of the parallel decode of one stream (inflator.c), where
 * smaller is faster -- 512 KiB is the balance */
#define DEFAULT_CHUNK   ((size_t) 512 << 10)
#define DEFAULT_BLOCKSEGS 8

struct TDEFLTPblc {
	uint32 state;
	uint32 error;
	uint32 flags;
	uint32 flush;
	uint32 status;
	const uint8* source;
	const uint8* sbgn;
	const uint8* send;
	uint8* target;
	uint8* tbgn;
	uint8* tend;
};

/* one batch in flight: FREE -> (filled) -> RUNNING -> DRAINING -> FREE */
enum { SLOT_FREE = 0, SLOT_RUNNING = 1, SLOT_DRAINING = 2 };

struct jdb_slot {
	jdb_dbuf stage;         /* gathered input (host sources)                     */
	size_t   stagelen;
	jdb_dbuf work;          /* pipeline workspace, holds the compressed bytes    */
	const uint8* out;       /* compressed bytes of the batch (device)            */
	uint64_t* htotal;       /* pinned: compressed size, valid after `done`       */
	jdb_event done;
	size_t   outlen;
	size_t   outpos;
	int      state;
	int      closes;        /* carries the flush / end marker of the request     */
	int      inplace;       /* reads the caller's device memory                  */
};

struct TDEFLTPrvt {
	struct TDEFLTPblc public;

	const TAllocator* allctr;
	int32  level;
	int    device;          /* the GPU this instance lives on (default device at creation) */
	uint32 used;
	uint32 closed;          /* the marker of the current request has been written */
	uint32 closing;         /* ... has been launched                               */
	uint32 dict_region;     /* preset dictionary staged in front of the first batch (bytes, 0: none) */
	uint32 dict_pad;

	jdb_deflate_cfg cfg;
	size_t batchcap;
	/* staged (host) input ramps up: the first batch is an eighth of batchcap and every
	 * launch doubles it, so that the host-to-device copy of the first batch -- which
	 * nothing can overlap -- is short and the pipeline fills quickly */
	size_t stagecap;

	/* two batches overlap: while one runs on `stream`, the other one's input is
	 * gathered (H2D) and the previous one's output is drained (D2H) on `cstream` */
	jdb_stream stream;
	jdb_stream cstream;
	struct jdb_slot slot[2];
	int fil;                /* slot receiving input                              */
	int fifo[2];            /* launch order of the RUNNING slots                 */
	int nfifo;
	int drn;                /* slot being drained, -1 none                       */

	/* checksums of the uncompressed bytes for zstrm */
	int       checks;       /* JDB_CK_* mask */
	uint32_t* dchecks;      /* device: [0] crc register, [1] adler */
	uint32_t* hchecks;      /* pinned mirror              */
	jdb_dbuf  ckwork;
};

static int stage_reserve(struct TDEFLTPrvt* state, struct jdb_slot* sl, size_t want);

typedef char jdb_deflator_layout_check[(sizeof(struct TDeflator) == sizeof(struct TDEFLTPblc)) ? 1 : -1];

#define PRVT ((struct TDEFLTPrvt*) state)
#define PBLC ((struct TDEFLTPblc*) state)

#define set_level jdb_level_params

static size_t
env_size(const char* name, size_t dflt)
{
	const char* v = getenv(name);
	if (v && *v) {
		unsigned long long x = strtoull(v, NULL, 10);
		if (x) return (size_t) x;
	}
	return dflt;
}

static size_t
first_stagecap(const struct TDEFLTPrvt* state)
{
	size_t chunk = state->cfg.chunk_bytes;
	size_t cap = state->batchcap / 8 / chunk * chunk;
	if (cap < chunk) cap = chunk;
	if (cap > state->batchcap) cap = state->batchcap;
	return cap;
}

TDeflator*
deflator_create(uintxx flags, intxx level, const TAllocator* allctr)
{
	struct TDeflator* state;
	size_t chunk;

	if (level > 9 || level < 0) {
		return NULL;
	}
	if (allctr == NULL) {
		allctr = ctb_getdefaultallocator();
	}
	/* no CUDA device, no encoder: there is no CPU path */
	if (jdb_rt_init() != JDB_OK) {
		return NULL;
	}
	state = allctr->request(sizeof(struct TDEFLTPrvt), allctr->user);
	if (state == NULL) {
		return NULL;
	}
	memset(state, 0, sizeof(struct TDEFLTPrvt));
	PRVT->allctr = allctr;
	PRVT->level = (int32) level;
	PRVT->device = jdb_rt_current_device();

	set_level(&PRVT->cfg, (int) level);
	chunk = env_size("JDB200_CHUNK_KIB", DEFAULT_CHUNK >> 10) << 10;
	chunk = (chunk + 2 * JDB_SEG - 1) / (2 * JDB_SEG) * (2 * JDB_SEG);      /* an LZ CTA takes two segments of one chunk */
	PRVT->cfg.chunk_bytes = (uint32_t) chunk;
	PRVT->cfg.block_segs = (uint32_t) env_size("JDB200_BLOCK_SEGS", DEFAULT_BLOCKSEGS);
	if (PRVT->cfg.block_segs > 16) PRVT->cfg.block_segs = 16;
	PRVT->cfg.chain_range = (uint32_t) env_size("JDB200_CHAIN_RANGE_KIB", 0) << 10;     /* 0: chosen per batch */
	PRVT->cfg.fixedonly = (flags & DEFLT_FIXEDCODES) != 0;
	PRVT->batchcap = env_size("JDB200_BATCH_MIB", DEFAULT_BATCH >> 20) << 20;
	PRVT->batchcap = (PRVT->batchcap + chunk - 1) / chunk * chunk;

	if (jdb_stream_create(&PRVT->stream) != JDB_OK || jdb_stream_create(&PRVT->cstream) != JDB_OK) {
		goto L_FAIL;
	}
	{
		int i;
		for (i = 0; i < 2; i++) {
			PRVT->slot[i].htotal = jdb_pinned_alloc(64);
			if (PRVT->slot[i].htotal == NULL || jdb_event_create(&PRVT->slot[i].done) != JDB_OK) {
				goto L_FAIL;
			}
		}
	}
	PRVT->hchecks = jdb_pinned_alloc(64);
	PRVT->dchecks = jdb_dev_alloc(64);
	if (PRVT->hchecks == NULL || PRVT->dchecks == NULL) {
		goto L_FAIL;
	}
	deflator_reset(state);
	PBLC->flags = (uint32) flags;
	return state;

L_FAIL:
	deflator_destroy(state);
	return NULL;
}

void
deflator_reset(TDeflator* state)
{
	CTB_ASSERT(state);

	PBLC->state = 0;
	PBLC->flush = 0;
	PBLC->error = 0;
	PBLC->status = 0;
	PBLC->source = PBLC->sbgn = PBLC->send = NULL;
	PBLC->target = PBLC->tbgn = PBLC->tend = NULL;

	PRVT->used = 0;
	PRVT->closed = 0;
	PRVT->closing = 0;
	PRVT->dict_region = 0;
	PRVT->dict_pad = 0;
	if (PRVT->stream) {
		jdb_rt_use_device(PRVT->device);
		/* nothing of an abandoned request may still be running */
		jdb_stream_sync(PRVT->stream);
		jdb_stream_sync(PRVT->cstream);
	}
	{
		int i;
		for (i = 0; i < 2; i++) {
			PRVT->slot[i].stagelen = 0;
			PRVT->slot[i].outlen = PRVT->slot[i].outpos = 0;
			PRVT->slot[i].state = SLOT_FREE;
			PRVT->slot[i].closes = PRVT->slot[i].inplace = 0;
		}
	}
	PRVT->fil = 0;
	PRVT->nfifo = 0;
	PRVT->drn = -1;
	PRVT->stagecap = first_stagecap(PRVT);
	if (PRVT->hchecks) {
		PRVT->hchecks[0] = 0xffffffffu;
		PRVT->hchecks[1] = 1u;
		jdb_copy_async(PRVT->dchecks, PRVT->hchecks, 8, PRVT->stream);
		jdb_stream_sync(PRVT->stream);
	}
}

void
deflator_destroy(TDeflator* state)
{
	const TAllocator* a;

	if (state == NULL) {
		return;
	}
	if (PRVT->stream) {
		jdb_stream_sync(PRVT->stream);
	}
	if (PRVT->cstream) {
		jdb_stream_sync(PRVT->cstream);
	}
	{
		int i;
		for (i = 0; i < 2; i++) {
			jdb_dbuf_release(&PRVT->slot[i].stage);
			jdb_dbuf_release(&PRVT->slot[i].work);
			jdb_pinned_free(PRVT->slot[i].htotal);
			jdb_event_destroy(PRVT->slot[i].done);
		}
	}
	jdb_dbuf_release(&PRVT->ckwork);
	jdb_pinned_free(PRVT->hchecks);
	jdb_dev_free(PRVT->dchecks);
	jdb_stream_destroy(PRVT->stream);
	jdb_stream_destroy(PRVT->cstream);
	a = PRVT->allctr;
	a->dispose(state, sizeof(struct TDEFLTPrvt), a->user);
}

void
deflator_setdctnr(TDeflator* state, const uint8* dict, uintxx size)
{
	struct jdb_slot* sl;
	size_t blockbytes, region, pad;
	CTB_ASSERT(state && dict && size);

	if (PBLC->state == POISON) {
		return;
	}
	/* only before the first deflate call, and only once (src/deflator.c:2116-2120) */
	if (PRVT->used || PRVT->dict_region) {
		if (PBLC->error == 0) {
			PBLC->error = DEFLT_EINCORRECTUSE;
		}
		PBLC->state = POISON;
		return;
	}
	if (PRVT->level == 0) {
		return;                 /* stored blocks never reference it (the reference: no-op at level 0) */
	}
	if (jdb_rt_use_device(PRVT->device) != JDB_OK) {
		return;
	}
	/* the last 32 KiB are what a match can reach */
	if (size > 32768) {
		dict += size - 32768;
		size = 32768;
	}
	/* The dictionary is history for the FIRST chunk only (chunks are independent): it is
	 * staged in front of the first batch as whole DEFLATE-block slots -- zero padding, then
	 * the dictionary -- which the kernels use as match source but never emit. */
	blockbytes = (size_t) PRVT->cfg.block_segs * JDB_SEG;
	region = (size + blockbytes - 1) / blockbytes * blockbytes;
	pad = region - size;
	if ((size_t) PRVT->cfg.chunk_bytes < region + blockbytes || PRVT->batchcap < region + blockbytes) {
		return;                 /* chunk too small to hold dictionary + data: compress without it (still a valid stream) */
	}
	sl = &PRVT->slot[PRVT->fil];
	if (stage_reserve(PRVT, sl, region + blockbytes) != 0 ||
	    jdb_memset_async(sl->stage.ptr, 0, pad, PRVT->cstream) != JDB_OK ||
	    jdb_copy_async(sl->stage.ptr + pad, dict, size, PRVT->cstream) != JDB_OK ||
	    jdb_stream_sync(PRVT->cstream) != JDB_OK) {
		if (PBLC->error == 0) {
			PBLC->error = DEFLT_EOOM;
		}
		PBLC->state = POISON;
		return;
	}
	sl->stagelen = region;
	PRVT->dict_region = (uint32) region;
	PRVT->dict_pad = (uint32) pad;
}

/* internal hooks for zstrm.c (hidden visibility) */
void
jdb_deflator_set_checks(TDeflator* state, int which)
{
	PRVT->checks = which;
}

int
jdb_deflator_get_checks(TDeflator* state, uint32* crc, uint32* adler)
{
	jdb_rt_use_device(PRVT->device);
	if (jdb_copy_async(PRVT->hchecks, PRVT->dchecks, 8, PRVT->stream) != JDB_OK ||
	    jdb_stream_sync(PRVT->stream) != JDB_OK) {
		return -1;
	}
	*crc = PRVT->hchecks[0];
	*adler = PRVT->hchecks[1];
	return 0;
}

/* usage rules of the reference, src/deflator.c:663-688 */
static int
validate(struct TDEFLTPrvt* state)
{
	if (PBLC->source == NULL || PBLC->target == NULL) {
		PBLC->error = DEFLT_EINCORRECTUSE;
		return 0;
	}
	if (PBLC->status == DEFLT_SRCEXHSTD) {
		if (PBLC->source == PBLC->send && PBLC->flush == 0) {
			PBLC->error = DEFLT_EINCORRECTUSE;
			return 0;
		}
	}
	if (PBLC->status == DEFLT_TGTEXHSTD) {
		if (PBLC->target == PBLC->tend) {
			PBLC->error = DEFLT_EINCORRECTUSE;
			return 0;
		}
	}
	return 1;
}

/* queue one pipeline launch over in[0..n) (device memory) on the compute stream */
static int
launch_batch(struct TDEFLTPrvt* state, int k, const uint8* in, size_t n, int closes, int inplace)
{
	struct jdb_slot* sl = &PRVT->slot[k];
	size_t need;
	uint8_t* out;
	uint64_t* dtotal;

	size_t skip = 0;

	PRVT->cfg.final = (uint32_t) (closes && PBLC->flush == DEFLT_END);
	PRVT->cfg.dict_region = 0;
	PRVT->cfg.dict_pad = 0;
	if (PRVT->dict_region && !inplace && n >= PRVT->dict_region) {
		/* first batch of a stream with a preset dictionary */
		PRVT->cfg.dict_region = PRVT->dict_region;
		PRVT->cfg.dict_pad = PRVT->dict_pad;
		skip = PRVT->dict_region;
		PRVT->dict_region = 0;
	}
	need = jdb_deflate_workspace_bytes(n, &PRVT->cfg);
	if (need == 0 || jdb_dbuf_reserve(&sl->work, need) != 0) {
		return -1;
	}
	if (PRVT->checks && n > skip) {
		if (jdb_dbuf_reserve(&PRVT->ckwork, jdb_checksum_workspace_bytes()) != 0) {
			return -1;
		}
		if (jdb_checksum(in + skip, n - skip, PRVT->checks, PRVT->dchecks, PRVT->dchecks + 1,
		                 PRVT->ckwork.ptr, PRVT->stream) != JDB_OK) {
			return -1;
		}
	}
	if (jdb_deflate_run(in, n, &PRVT->cfg, sl->work.ptr, &out, &dtotal, PRVT->stream) != JDB_OK) {
		return -1;
	}
	if (jdb_copy_async(sl->htotal, dtotal, 8, PRVT->stream) != JDB_OK ||
	    jdb_event_record(sl->done, PRVT->stream) != JDB_OK) {
		return -1;
	}
	sl->out = out;
	sl->outlen = 0;
	sl->outpos = 0;
	sl->state = SLOT_RUNNING;
	sl->closes = closes;
	sl->inplace = inplace;
	sl->stagelen = 0;           /* the stage buffer itself is read by the kernels until `done` */
	PRVT->fifo[PRVT->nfifo++] = k;
	return 0;
}

/* wait for the oldest running batch; it becomes the one being drained */
static int
retire_oldest(struct TDEFLTPrvt* state)
{
	int k = PRVT->fifo[0];
	struct jdb_slot* sl = &PRVT->slot[k];

	if (jdb_event_sync(sl->done) != JDB_OK) {
		return -1;
	}
	PRVT->fifo[0] = PRVT->fifo[1];
	PRVT->nfifo--;
	sl->outlen = (size_t) sl->htotal[0];
	sl->outpos = 0;
	sl->state = SLOT_DRAINING;
	PRVT->drn = k;
	return 0;
}

/* batches that read the caller's device memory must be done before we return */
static int
settle_inplace(struct TDEFLTPrvt* state)
{
	int i;

	for (i = 0; i < PRVT->nfifo; i++) {
		struct jdb_slot* sl = &PRVT->slot[PRVT->fifo[i]];
		if (sl->inplace && jdb_event_sync(sl->done) != JDB_OK) {
			return -1;
		}
	}
	return 0;
}

/* grow the stage buffer of a slot (contents preserved) */
static int
stage_reserve(struct TDEFLTPrvt* state, struct jdb_slot* sl, size_t want)
{
	jdb_dbuf bigger;
	size_t cap;

	if (sl->stage.ptr != NULL && sl->stage.cap >= want) {
		return 0;
	}
	cap = sl->stage.cap ? sl->stage.cap * 2 : ((size_t) 1 << 20);
	while (cap < want) {
		cap *= 2;
	}
	if (cap > PRVT->batchcap) {
		cap = PRVT->batchcap;
	}
	bigger.ptr = NULL;
	bigger.cap = 0;
	if (jdb_dbuf_reserve(&bigger, cap) != 0) {
		return -1;
	}
	if (sl->stagelen) {
		if (jdb_copy_async(bigger.ptr, sl->stage.ptr, sl->stagelen, PRVT->cstream) != JDB_OK ||
		    jdb_stream_sync(PRVT->cstream) != JDB_OK) {
			jdb_dbuf_release(&bigger);
			return -1;
		}
	}
	/* the old buffer may still be read by a batch launched from it */
	if (PRVT->nfifo && jdb_stream_sync(PRVT->stream) != JDB_OK) {
		jdb_dbuf_release(&bigger);
		return -1;
	}
	jdb_dbuf_release(&sl->stage);
	sl->stage = bigger;
	return 0;
}

eDEFLTResult
deflator_deflate(TDeflator* state, eDEFLTFlush flush)
{
	CTB_ASSERT(state);

	if (PBLC->state == POISON) {
		return DEFLT_ERROR;
	}
	/* a flush request is latched; END cannot be downgraded
	 * (src/deflator.c:696-699) */
	if (flush && (PBLC->flush == 0 || PBLC->flush == DEFLT_FLUSH)) {
		PBLC->flush = (uint32) flush;
	}
	if (validate(PRVT) == 0) {
		PBLC->state = POISON;
		return DEFLT_ERROR;
	}
	if (jdb_rt_use_device(PRVT->device) != JDB_OK) {
		PBLC->error = DEFLT_EBADSTATE;
		PBLC->state = POISON;
		return DEFLT_ERROR;
	}
	PRVT->used = 1;

	for (;;) {
		struct jdb_slot* fs;
		size_t srcleft;
		int all_in;

		/* 1. drain the compressed bytes of the oldest finished batch */
		if (PRVT->drn >= 0) {
			struct jdb_slot* sl = &PRVT->slot[PRVT->drn];
			size_t room = (size_t) (PBLC->tend - PBLC->target);
			size_t n = sl->outlen - sl->outpos;
			if (n > room) {
				n = room;
			}
			if (n) {
				if (jdb_copy_async(PBLC->target, sl->out + sl->outpos, n, PRVT->cstream) != JDB_OK ||
				    jdb_stream_sync(PRVT->cstream) != JDB_OK) {
					goto L_FAIL;
				}
				PBLC->target += n;
				sl->outpos += n;
			}
			if (sl->outpos < sl->outlen) {
				if (settle_inplace(PRVT) != 0) {
					goto L_FAIL;
				}
				return (eDEFLTResult) (PBLC->status = DEFLT_TGTEXHSTD);
			}
			if (sl->closes) {
				PRVT->closed = 1;
			}
			sl->state = SLOT_FREE;
			PRVT->drn = -1;
		}

		/* 2. everything requested has been written */
		if (PRVT->closed) {
			PRVT->closing = 0;
			if (PBLC->flush == DEFLT_FLUSH) {
				/* keep going on the same instance (src/deflator.c:763-768) */
				PBLC->flush = 0;
				PRVT->closed = 0;
			}
			else {
				PBLC->state = POISON;
			}
			return (eDEFLTResult) (PBLC->status = DEFLT_OK);
		}

		fs = &PRVT->slot[PRVT->fil];
		srcleft = (size_t) (PBLC->send - PBLC->source);

		if (fs->state == SLOT_FREE && !PRVT->closing) {
			/* 3. large device-resident input is compressed in place */
			if (fs->stagelen == 0 && srcleft &&
			    (PBLC->flush || srcleft >= PRVT->batchcap) &&
			    ((uintptr_t) PBLC->source & 15) == 0 && jdb_ptr_is_device(PBLC->source)) {
				size_t n = srcleft;
				int last;
				if (n > PRVT->batchcap) {
					n = PRVT->batchcap;
				}
				last = PBLC->flush && n == srcleft;
				if (launch_batch(PRVT, PRVT->fil, PBLC->source, n, last, 1) != 0) {
					goto L_FAIL;
				}
				PBLC->source += n;
				if (last) {
					PRVT->closing = 1;
				}
				PRVT->fil ^= 1;
				continue;
			}

			/* 4. gather into the staging batch */
			if (srcleft) {
				size_t room = PRVT->stagecap > fs->stagelen ? PRVT->stagecap - fs->stagelen : 0;
				if (srcleft > room) {
					srcleft = room;
				}
				if (srcleft) {
					if (stage_reserve(PRVT, fs, fs->stagelen + srcleft) != 0) {
						goto L_FAIL;
					}
					/* the caller may reuse its buffer as soon as we return: wait for
					 * the copy (only the copy -- a running batch keeps running) */
					if (jdb_copy_async(fs->stage.ptr + fs->stagelen, PBLC->source, srcleft, PRVT->cstream) != JDB_OK ||
					    jdb_stream_sync(PRVT->cstream) != JDB_OK) {
						goto L_FAIL;
					}
					PBLC->source += srcleft;
					fs->stagelen += srcleft;
				}
			}

			/* 5. queue the pipeline when the batch is full or a flush is due */
			all_in = PBLC->source == PBLC->send;
			if (fs->stagelen >= PRVT->stagecap || (PBLC->flush && all_in)) {
				int last = PBLC->flush && all_in;
				if (fs->stage.ptr == NULL && stage_reserve(PRVT, fs, 4096) != 0) {
					goto L_FAIL;
				}
				if (launch_batch(PRVT, PRVT->fil, fs->stage.ptr, fs->stagelen, last, 0) != 0) {
					goto L_FAIL;
				}
				if (last) {
					PRVT->closing = 1;
				}
				PRVT->fil ^= 1;
				PRVT->stagecap = PRVT->stagecap * 2 < PRVT->batchcap ? PRVT->stagecap * 2 : PRVT->batchcap;
				continue;
			}
			if (all_in) {
				if (settle_inplace(PRVT) != 0) {
					goto L_FAIL;
				}
				return (eDEFLTResult) (PBLC->status = DEFLT_SRCEXHSTD);
			}
			continue;
		}

		/* 6. no slot to fill (or the request is already closing): wait for the
		 * oldest batch in flight; its output is drained next */
		if (PRVT->nfifo == 0) {
			PBLC->error = DEFLT_EBADSTATE;
			PBLC->state = POISON;
			return DEFLT_ERROR;
		}
		if (retire_oldest(PRVT) != 0) {
			goto L_FAIL;
		}
	}

L_FAIL:
	if (PBLC->error == 0) {
		PBLC->error = DEFLT_EOOM;
	}
	PBLC->state = POISON;
	return DEFLT_ERROR;
}
