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
#define DEFAULT_CHUNK   ((size_t) 256 << 10)
#define DEFAULT_BLOCKSEGS 4

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

struct TDEFLTPrvt {
	struct TDEFLTPblc public;

	const TAllocator* allctr;
	int32  level;
	uint32 used;
	uint32 closed;          /* the final marker has been produced */

	jdb_deflate_cfg cfg;
	size_t batchcap;

	jdb_stream stream;
	jdb_dbuf   stage;       /* gathered input */
	size_t     stagelen;
	jdb_dbuf   work;

	const uint8* outptr;    /* compressed bytes of the last batch (device) */
	size_t  outlen;
	size_t  outpos;

	uint64_t* htotal;       /* pinned */

	/* checksums of the uncompressed bytes for zstrm */
	int       checks;       /* JDB_CK_* mask */
	uint32_t* dchecks;      /* device: [0] crc register, [1] adler */
	uint32_t* hchecks;      /* pinned mirror */
	jdb_dbuf  ckwork;
};

typedef char jdb_deflator_layout_check[(sizeof(struct TDeflator) == sizeof(struct TDEFLTPblc)) ? 1 : -1];

#define PRVT ((struct TDEFLTPrvt*) state)
#define PBLC ((struct TDEFLTPblc*) state)

/* level -> (good, nice, chain), lazy for 6..9: src/deflator.c:241-263 */
static void
set_level(jdb_deflate_cfg* cfg, int level)
{
	static const uint16_t t[10][3] = {
		{0, 0, 0}, {8, 4, 2}, {8, 8, 8}, {8, 16, 16}, {8, 32, 32}, {8, 64, 128},
		{16, 16, 48}, {32, 64, 128}, {64, 128, 320}, {192, 256, 512}
	};
	cfg->level = (uint32_t) level;
	cfg->good = t[level][0];
	cfg->nice = t[level][1];
	cfg->chain = t[level][2];
	cfg->lazy = level >= 6;
}

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

	set_level(&PRVT->cfg, (int) level);
	chunk = env_size("JDB200_CHUNK_KIB", DEFAULT_CHUNK >> 10) << 10;
	chunk = (chunk + JDB_SEG - 1) / JDB_SEG * JDB_SEG;
	PRVT->cfg.chunk_bytes = (uint32_t) chunk;
	PRVT->cfg.block_segs = (uint32_t) env_size("JDB200_BLOCK_SEGS", DEFAULT_BLOCKSEGS);
	if (PRVT->cfg.block_segs > 16) PRVT->cfg.block_segs = 16;
	PRVT->cfg.chain_range = (uint32_t) env_size("JDB200_CHAIN_RANGE_KIB", 0) << 10;     /* 0: chosen per batch */
	PRVT->cfg.fixedonly = (flags & DEFLT_FIXEDCODES) != 0;
	PRVT->batchcap = env_size("JDB200_BATCH_MIB", DEFAULT_BATCH >> 20) << 20;
	PRVT->batchcap = (PRVT->batchcap + chunk - 1) / chunk * chunk;

	if (jdb_stream_create(&PRVT->stream) != JDB_OK) {
		goto L_FAIL;
	}
	PRVT->htotal = jdb_pinned_alloc(64);
	PRVT->hchecks = jdb_pinned_alloc(64);
	PRVT->dchecks = jdb_dev_alloc(64);
	if (PRVT->htotal == NULL || PRVT->hchecks == NULL || PRVT->dchecks == NULL) {
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
	PRVT->stagelen = 0;
	PRVT->outptr = NULL;
	PRVT->outlen = 0;
	PRVT->outpos = 0;
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
	jdb_dbuf_release(&PRVT->stage);
	jdb_dbuf_release(&PRVT->work);
	jdb_dbuf_release(&PRVT->ckwork);
	jdb_pinned_free(PRVT->htotal);
	jdb_pinned_free(PRVT->hchecks);
	jdb_dev_free(PRVT->dchecks);
	jdb_stream_destroy(PRVT->stream);
	a = PRVT->allctr;
	a->dispose(state, sizeof(struct TDEFLTPrvt), a->user);
}

void
deflator_setdctnr(TDeflator* state, const uint8* dict, uintxx size)
{
	CTB_ASSERT(state && dict && size);
	(void) dict;
	(void) size;
	/* the reference accepts a dictionary only before the first deflate call
	 * (src/deflator.c:2116-2120).  Preset dictionaries are row f2 ("next") of
	 * the scope table and not implemented yet: refuse loudly instead of
	 * silently producing a stream the peer cannot decode. */
	if (PBLC->error == 0) {
		PBLC->error = DEFLT_EINCORRECTUSE;
	}
	PBLC->state = POISON;
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

/* one pipeline launch over in[0..n) (device memory); blocks until the size is known */
static int
compress_batch(struct TDEFLTPrvt* state, const uint8* in, size_t n, int final)
{
	size_t need;
	uint8_t* out;
	uint64_t* dtotal;

	PRVT->cfg.final = (uint32_t) final;
	need = jdb_deflate_workspace_bytes(n, &PRVT->cfg);
	if (need == 0 || jdb_dbuf_reserve(&PRVT->work, need) != 0) {
		return -1;
	}
	if (PRVT->checks && n) {
		if (jdb_dbuf_reserve(&PRVT->ckwork, jdb_checksum_workspace_bytes()) != 0) {
			return -1;
		}
		if (jdb_checksum(in, n, PRVT->checks, PRVT->dchecks, PRVT->dchecks + 1,
		                 PRVT->ckwork.ptr, PRVT->stream) != JDB_OK) {
			return -1;
		}
	}
	if (jdb_deflate_run(in, n, &PRVT->cfg, PRVT->work.ptr, &out, &dtotal, PRVT->stream) != JDB_OK) {
		return -1;
	}
	if (jdb_copy_async(PRVT->htotal, dtotal, 8, PRVT->stream) != JDB_OK ||
	    jdb_stream_sync(PRVT->stream) != JDB_OK) {
		return -1;
	}
	PRVT->outptr = out;
	PRVT->outlen = (size_t) PRVT->htotal[0];
	PRVT->outpos = 0;
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
	if (jdb_rt_init() != JDB_OK) {
		PBLC->error = DEFLT_EBADSTATE;
		PBLC->state = POISON;
		return DEFLT_ERROR;
	}
	PRVT->used = 1;

	for (;;) {
		size_t srcleft;

		/* 1. drain what the last batch produced */
		if (PRVT->outpos < PRVT->outlen) {
			size_t room = (size_t) (PBLC->tend - PBLC->target);
			size_t n = PRVT->outlen - PRVT->outpos;
			if (n > room) {
				n = room;
			}
			if (n) {
				if (jdb_copy_async(PBLC->target, PRVT->outptr + PRVT->outpos, n, PRVT->stream) != JDB_OK ||
				    jdb_stream_sync(PRVT->stream) != JDB_OK) {
					goto L_FAIL;
				}
				PBLC->target += n;
				PRVT->outpos += n;
			}
			if (PRVT->outpos < PRVT->outlen) {
				return (eDEFLTResult) (PBLC->status = DEFLT_TGTEXHSTD);
			}
		}

		/* 2. everything requested has been written */
		if (PRVT->closed) {
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

		srcleft = (size_t) (PBLC->send - PBLC->source);

		/* 3. large device-resident input is compressed in place */
		if (PRVT->stagelen == 0 && srcleft &&
		    (PBLC->flush || srcleft >= PRVT->batchcap) &&
		    ((uintptr_t) PBLC->source & 15) == 0 && jdb_ptr_is_device(PBLC->source)) {
			size_t n = srcleft;
			int last;
			if (n > PRVT->batchcap) {
				n = PRVT->batchcap;
			}
			last = PBLC->flush && n == srcleft;
			if (compress_batch(PRVT, PBLC->source, n, last && PBLC->flush == DEFLT_END) != 0) {
				goto L_FAIL;
			}
			PBLC->source += n;
			if (last) {
				PRVT->closed = 1;
			}
			continue;
		}

		/* 4. gather into the staging batch */
		if (srcleft) {
			size_t room;
			if (jdb_dbuf_reserve(&PRVT->stage, PRVT->batchcap) != 0 && PRVT->stagelen == 0) {
				goto L_FAIL;
			}
			if (PRVT->stage.ptr == NULL) {
				goto L_FAIL;
			}
			room = PRVT->batchcap - PRVT->stagelen;
			if (srcleft > room) {
				srcleft = room;
			}
			if (jdb_copy_async(PRVT->stage.ptr + PRVT->stagelen, PBLC->source, srcleft, PRVT->stream) != JDB_OK) {
				goto L_FAIL;
			}
			/* the caller may reuse its buffer as soon as we return */
			if (jdb_stream_sync(PRVT->stream) != JDB_OK) {
				goto L_FAIL;
			}
			PBLC->source += srcleft;
			PRVT->stagelen += srcleft;
		}

		/* 5. run the pipeline when the batch is full or a flush is due */
		{
			int all_in = PBLC->source == PBLC->send;
			int full = PRVT->stagelen == PRVT->batchcap;
			if (full || (PBLC->flush && all_in)) {
				int last = PBLC->flush && all_in;
				if (PRVT->stage.ptr == NULL && jdb_dbuf_reserve(&PRVT->stage, 4096) != 0) {
					goto L_FAIL;
				}
				if (compress_batch(PRVT, PRVT->stage.ptr, PRVT->stagelen, last && PBLC->flush == DEFLT_END) != 0) {
					goto L_FAIL;
				}
				PRVT->stagelen = 0;
				if (last) {
					PRVT->closed = 1;
				}
				continue;
			}
			if (all_in) {
				return (eDEFLTResult) (PBLC->status = DEFLT_SRCEXHSTD);
			}
		}
	}

L_FAIL:
	if (PBLC->error == 0) {
		PBLC->error = DEFLT_EOOM;
	}
	PBLC->state = POISON;
	return DEFLT_ERROR;
}
