/*
 * records.cu -- batched compression of many independent records, the compress
 * side mirror of the batched inflate.
 *
 * The reference can only do this as a loop of deflator_reset + deflator_deflate
 * calls (src/deflator.c:2069-2104, 2169-2282), one stream after the other.  Here a
 * group of records becomes ONE pass of the chunk-parallel deflate pipeline
 * (deflate.cu) over "ragged chunks" (deflate.cuh): every record is laid into
 * whole chunk slots of a staging area, all its chunks but the last are full, and
 * the last one closes the record's stream with BFINAL.
 *
 *   records_gather_kernel   one CTA per record: record bytes -> chunk slots
 *                           (any source alignment), chunk_len[] of its chunks
 *   (jdb_deflate_run)       chain / lz / huffman / layout / pack over the slots
 *   records_finish_kernel   one CTA per record: Adler-32 of the record and the
 *                           zlib header / trailer (JDB_FMT_ZLIB), then the
 *                           record's compressed bytes -> its target range,
 *                           result entry
 *
 * HBM layout of a group: slots (nchunks * chunk_bytes), chunk_len (4 B / chunk),
 * the pipeline workspace, and the caller's source / target / item / result arrays.
 * Algorithmic traffic per record: N read + C written; implementation: + N written
 * and read again (slots) + C read and written again (placement).
 */
#include "deflate.cuh"

#define RC_THREADS 128
#define ADLER_MOD  65521u

extern "C" const void* jdb_deflate_chunk_table(uint64_t n, const jdb_deflate_cfg* cfg, const void* work);

__global__ void __launch_bounds__(RC_THREADS)
records_gather_kernel(const uint8_t* __restrict__ src_base, const jdb_inflate_item* __restrict__ items,
                      const uint32_t* __restrict__ first_chunk, uint32_t chunk_base, uint32_t chunk_bytes,
                      uint8_t* __restrict__ slots, uint32_t* __restrict__ chunk_len)
{
	const uint32_t r = blockIdx.x;
	const uint32_t tid = threadIdx.x;
	const uint64_t len = items[r].src_len;
	const uint32_t c0 = first_chunk[r] - chunk_base;
	const uint32_t nch = len ? (uint32_t) ((len + chunk_bytes - 1) / chunk_bytes) : 1u;

	for (uint32_t k = tid; k < nch; k += RC_THREADS) {
		const uint64_t left = len - (uint64_t) k * chunk_bytes;
		uint32_t v = left < chunk_bytes ? (uint32_t) left : chunk_bytes;
		if (k == 0) v |= CHUNK_FIRST;
		if (k == nch - 1) v |= CHUNK_LAST;
		chunk_len[c0 + k] = v;
	}

	/* destination words from (possibly unaligned) source bytes: two aligned loads and a
	 * funnel shift; the second word is only touched when it holds bytes of the record */
	const uint8_t* src = src_base + items[r].src_off;
	uint32_t* dst = (uint32_t*) (slots + (uint64_t) c0 * chunk_bytes);
	const uint32_t mis = (uint32_t) ((uintptr_t) src & 3u);
	const uint32_t* sw = (const uint32_t*) (src - mis);
	const uint64_t nwords = (len + 3) / 4;
	for (uint64_t i = tid; i < nwords; i += RC_THREADS) {
		uint32_t w0 = sw[i], w1 = 0;
		if (mis && i * 4 + (4 - mis) < len) w1 = sw[i + 1];
		dst[i] = __funnelshift_r(w0, w1, mis * 8);
	}
	/* bytes behind the record inside its last 16-byte vector: the LZ stage loads vectors */
	const uint64_t padded = (nwords * 4 + 15) & ~(uint64_t) 15;
	for (uint64_t i = nwords + tid; i < padded / 4; i += RC_THREADS) dst[i] = 0;
}

__global__ void __launch_bounds__(RC_THREADS)
records_finish_kernel(const uint8_t* __restrict__ slots, uint8_t* __restrict__ out,
                      const ChunkInfo* __restrict__ chunks, const jdb_inflate_item* __restrict__ items,
                      const uint32_t* __restrict__ first_chunk, uint32_t chunk_base, uint32_t chunk_bytes,
                      uint32_t format, uint32_t level, uint8_t* __restrict__ tgt_base,
                      jdb_inflate_result* __restrict__ results)
{
	__shared__ unsigned long long red[2][RC_THREADS / 32];
	__shared__ uint32_t s_adler;
	const uint32_t r = blockIdx.x;
	const uint32_t tid = threadIdx.x;
	const uint64_t len = items[r].src_len;
	const uint32_t c0 = first_chunk[r] - chunk_base;
	const uint32_t nch = len ? (uint32_t) ((len + chunk_bytes - 1) / chunk_bytes) : 1u;
	const uint64_t o0 = chunks[c0].offset;
	const uint64_t o1 = chunks[c0 + nch - 1].offset + chunks[c0 + nch - 1].bytes;
	const uint64_t clen = o1 - o0;

	if (format == JDB_FMT_ZLIB) {
		/* Adler-32 of the record: a = 1 + sum d_i, b = len + sum (len - i) d_i */
		const uint8_t* d = slots + (uint64_t) c0 * chunk_bytes;
		const uint32_t* dw = (const uint32_t*) d;
		unsigned long long sa = 0, sb = 0;
		const uint64_t nfull = len / 4;
		for (uint64_t i = tid; i < nfull; i += RC_THREADS) {
			const uint32_t w = dw[i];
			const uint32_t s = __dp4a(w, 0x01010101u, 0u);
			const uint32_t q = __dp4a(w, 0x03020100u, 0u);
			sa += s;
			sb += (unsigned long long) (len - i * 4) * s - q;
			if ((i / RC_THREADS & 0xfffu) == 0xfffu) sb %= ADLER_MOD;
		}
		if (tid == 0)
			for (uint64_t i = nfull * 4; i < len; i++) { sa += d[i]; sb += (len - i) * d[i]; }
		sa %= ADLER_MOD;
		sb %= ADLER_MOD;
		for (int o = 16; o; o >>= 1) {
			sa += __shfl_down_sync(JDB_FULL_MASK, sa, o);
			sb += __shfl_down_sync(JDB_FULL_MASK, sb, o);
		}
		if ((tid & 31) == 0) { red[0][tid >> 5] = sa; red[1][tid >> 5] = sb; }
		__syncthreads();
		if (tid == 0) {
			unsigned long long a = 1, b = len % ADLER_MOD;
			for (int w = 0; w < RC_THREADS / 32; w++) { a += red[0][w]; b += red[1][w]; }
			a %= ADLER_MOD;
			b %= ADLER_MOD;
			const uint32_t adler = (uint32_t) ((b << 16) | a);
			s_adler = adler;
			/* RFC 1950 header: CM 8, CINFO 7, FLEVEL by level, FCHECK making it a multiple of 31 */
			const uint32_t flevel = level < 2 ? 0u : level < 6 ? 1u : level == 6 ? 2u : 3u;
			uint32_t hdr = (0x78u << 8) | (flevel << 6);
			hdr += 31u - hdr % 31u;
			out[o0] = (uint8_t) (hdr >> 8);
			out[o0 + 1] = (uint8_t) hdr;
			out[o1 - 4] = (uint8_t) (adler >> 24);
			out[o1 - 3] = (uint8_t) (adler >> 16);
			out[o1 - 2] = (uint8_t) (adler >> 8);
			out[o1 - 1] = (uint8_t) adler;
		}
		__syncthreads();
	}

	/* placement: the record's compressed bytes go to its target range if they fit */
	const bool fits = clen <= items[r].dst_cap;
	if (fits) {
		uint8_t* t = tgt_base + items[r].dst_off;
		const uint8_t* s = out + o0;
		for (uint64_t i = tid; i < clen; i += RC_THREADS) t[i] = s[i];
	}
	if (tid == 0) {
		jdb_inflate_result q;
		q.status = fits ? 0u : 2u;            /* DEFLT_OK : DEFLT_TGTEXHSTD */
		q.error = 0;
		q.zerror = 0;
		q.checksum = format == JDB_FMT_ZLIB ? s_adler : 0u;
		q.consumed = fits ? len : 0;
		q.produced = fits ? clen : 0;
		results[r] = q;
	}
}

extern "C" size_t jdb_records_workspace_bytes(uint64_t slot_bytes, const jdb_deflate_cfg* cfg)
{
	jdb_deflate_cfg c = *cfg;
	static const uint32_t nonnull = 0;
	c.chunk_len = &nonnull;                  /* ragged mode sizes the output for the wrappers */
	return jdb_deflate_workspace_bytes(slot_bytes, &c);
}

/*
 * Compress the records items[0..nrec) of one group.  first_chunk[] (device) holds
 * the first chunk slot of every record, counted from the start of the whole batch;
 * chunk_base is the slot the group starts at, nchunks the slots it owns.  slots,
 * chunk_len and work are scratch of at least nchunks * chunk_bytes (+ 16),
 * nchunks * 4 and jdb_records_workspace_bytes() bytes.
 */
extern "C" int jdb_records_deflate(const uint8_t* src_base, uint8_t* tgt_base,
                                   const jdb_inflate_item* items, jdb_inflate_result* results,
                                   const uint32_t* first_chunk, uint32_t nrec,
                                   uint32_t chunk_base, uint32_t nchunks, uint32_t format,
                                   const jdb_deflate_cfg* cfg, uint8_t* slots, uint32_t* chunk_len,
                                   void* work, jdb_stream s)
{
	if (nrec == 0) return JDB_OK;
	jdb_deflate_cfg c = *cfg;
	c.chunk_len = chunk_len;
	c.wrap_head = format == JDB_FMT_ZLIB ? 2u : 0u;
	c.wrap_tail = format == JDB_FMT_ZLIB ? 4u : 0u;
	c.final = 1;
	c.dict_region = 0;
	c.dict_pad = 0;
	const uint64_t n = (uint64_t) nchunks * c.chunk_bytes;

	JDB_LAUNCH(records_gather_kernel, dim3(nrec), dim3(RC_THREADS), 0, s,
	           src_base, items, first_chunk, chunk_base, c.chunk_bytes, slots, chunk_len);
	int r = jdb_rt_check_launch("records_gather_kernel");
	if (r != JDB_OK) return r;

	uint8_t* out = NULL;
	uint64_t* total = NULL;
	r = jdb_deflate_run(slots, n, &c, work, &out, &total, s);
	if (r != JDB_OK) return r;

	const ChunkInfo* chunks = (const ChunkInfo*) jdb_deflate_chunk_table(n, &c, work);
	JDB_LAUNCH(records_finish_kernel, dim3(nrec), dim3(RC_THREADS), 0, s,
	           (const uint8_t*) slots, out, chunks, items, first_chunk, chunk_base, c.chunk_bytes,
	           format, c.level, tgt_base, results);
	return jdb_rt_check_launch("records_finish_kernel");
}
