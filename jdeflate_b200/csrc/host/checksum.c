/*
 * checksum.c -- zstrm_crc32update / zstrm_adler32update / zstrm_crc32combine,
 * the stand-alone checksum entry points of the reference
 * (jdeflate/zstrm.h:203-223, src/zstrm.c:1316-1576) on top of the
 * chunk-parallel kernels of csrc/device/checksum.cu.
 *
 * Host buffers are staged to the GPU in pieces; device buffers are read in
 * place.  There is no CPU checksum path.
 */
#include <jdeflate/zstrm.h>
#include "jdb_host.h"

#define CK_STAGE_BYTES ((size_t) 64 << 20)

/* per-thread scratch: the helpers are stateless in the reference and may be
 * called concurrently from different threads */
static __thread struct {
	jdb_stream stream;
	jdb_dbuf   stage;
	jdb_dbuf   work;
	uint32_t*  dvals;     /* device: [0] crc, [1] adler */
	uint32_t*  hvals;     /* pinned mirror              */
	int        ready;
	int        device;    /* where stream and buffers live */
} ck;

static void
ck_prepare(void)
{
	if (ck.ready) {
		if (jdb_rt_init() != JDB_OK) {
			jdb_fatal("CUDA runtime lost");
		}
		if (ck.device == jdb_rt_current_device()) {
			return;
		}
		/* the default device moved (jdb200_set_device): this thread's scratch moves with it */
		jdb_stream_sync(ck.stream);
		jdb_stream_destroy(ck.stream);
		jdb_dbuf_release(&ck.stage);
		jdb_dbuf_release(&ck.work);
		jdb_dev_free(ck.dvals);
		jdb_pinned_free(ck.hvals);
		ck.ready = 0;
	}
	if (jdb_rt_init() != JDB_OK) {
		jdb_fatal("checksum helpers need a CUDA device");
	}
	if (jdb_stream_create(&ck.stream) != JDB_OK) {
		jdb_fatal("stream creation failed");
	}
	if (jdb_dbuf_reserve(&ck.work, jdb_checksum_workspace_bytes()) != 0) {
		jdb_fatal("out of device memory");
	}
	ck.dvals = (uint32_t*) jdb_dev_alloc(64);
	ck.hvals = (uint32_t*) jdb_pinned_alloc(64);
	if (ck.dvals == NULL || ck.hvals == NULL) {
		jdb_fatal("out of memory");
	}
	ck.device = jdb_rt_current_device();
	ck.ready = 1;
}

static uint32
ck_run(int which, uint32 value, const uint8* source, uintxx size)
{
	const uint8* p;
	uintxx left;
	int ondevice;
	int slot;

	if (size == 0) {
		return value;
	}
	ck_prepare();

	slot = (which == JDB_CK_CRC32) ? 0 : 1;
	ck.hvals[slot] = value;
	if (jdb_copy_async(ck.dvals + slot, ck.hvals + slot, 4, ck.stream) != JDB_OK) {
		jdb_fatal("copy failed");
	}

	ondevice = jdb_ptr_is_device(source);
	p = source;
	left = size;
	while (left) {
		uintxx n = left;
		const uint8* d = p;

		/* one launch covers at most 1 GiB: the kernels count vectors and combine lengths in
		 * 32 bits (the running value in device memory chains the launches) */
		if (n > ((uintxx) 1 << 30)) {
			n = (uintxx) 1 << 30;
		}
		if (!ondevice) {
			if (n > CK_STAGE_BYTES) {
				n = CK_STAGE_BYTES;
			}
			if (jdb_dbuf_reserve(&ck.stage, n) != 0) {
				jdb_fatal("out of device memory");
			}
			if (jdb_copy_async(ck.stage.ptr, p, n, ck.stream) != JDB_OK) {
				jdb_fatal("copy failed");
			}
			d = ck.stage.ptr;
		}
		if (jdb_checksum(d, n, which, ck.dvals, ck.dvals + 1, ck.work.ptr, ck.stream) != JDB_OK) {
			jdb_fatal("checksum kernel launch failed");
		}
		p += n;
		left -= n;
	}

	if (jdb_copy_async(ck.hvals + slot, ck.dvals + slot, 4, ck.stream) != JDB_OK ||
	    jdb_stream_sync(ck.stream) != JDB_OK) {
		jdb_fatal("checksum kernel failed");
	}
	return ck.hvals[slot];
}

uint32
zstrm_crc32update(uint32 chcksm, const uint8* source, uintxx size)
{
	CTB_ASSERT(source);
	return ck_run(JDB_CK_CRC32, chcksm, source, size);
}

uint32
zstrm_adler32update(uint32 chcksm, const uint8* source, uintxx size)
{
	CTB_ASSERT(source);
	return ck_run(JDB_CK_ADLER32, chcksm, source, size);
}


/* ---- combine (pure GF(2) arithmetic on two words, no data touched) ------ */

#define JDB_CRCPOLY 0xEDB88320u

static uint32
gf2mul(uint32 a, uint32 b)
{
	uint32 p;
	int i;

	p = 0;
	for (i = 0; i < 32; i++) {
		if (a & (0x80000000u >> i)) {
			p ^= b;
		}
		b = (b & 1u) ? (b >> 1) ^ JDB_CRCPOLY : (b >> 1);
	}
	return p;
}

/* crc of A||B from finalised crc(A), crc(B), len(B): crc(A) * x^(8 len) + crc(B)
 * (same result as the reference's matrix walk, src/zstrm.c:1413-1443, but the
 * length is a full 64-bit value) */
static uint32
crc_combine64(uint32 crc1, uint32 crc2, uint64 len2)
{
	uint32 xp;
	uint32 m;
	int k;

	/* x^8 */
	xp = 0x80000000u;
	for (k = 0; k < 8; k++) {
		xp = (xp & 1u) ? (xp >> 1) ^ JDB_CRCPOLY : (xp >> 1);
	}
	m = 0x80000000u;
	for (; len2; len2 >>= 1) {
		if (len2 & 1u) {
			m = gf2mul(m, xp);
		}
		xp = gf2mul(xp, xp);
	}
	return gf2mul(crc1, m) ^ crc2;
}

uint32
zstrm_crc32combine(uint32 crc1, uint32 crc2, uintxx size2)
{
	return crc_combine64(crc1, crc2, (uint64) size2);
}

uint32
crc32_ncombine(uint32 crc1, uint32 crc2, uint32 size2)
{
	return crc_combine64(crc1, crc2, (uint64) size2);
}
