/*
 * deflate.cu -- orchestration of the chunk-parallel deflate pipeline.
 *
 * One call compresses one batch of input that is already resident in HBM:
 *
 *   chain_kernel   (lz.cu)       links to the previous same-hash position
 *   lz_kernel      (lz.cu)       match search + lazy/greedy parse -> tokens, histograms
 *   huffman_kernel (huffman.cu)  per block code construction + block type choice
 *   layout_kernel  (pack.cu)     bit offsets, chunk size scan, total size
 *   pack_kernel    (pack.cu)     bit packing at the final offsets
 *
 * all on one stream, no host round trip in between; the host reads the total
 * compressed size afterwards.  The batch is cut into independent chunks
 * (BASELINE north_star (1)); every chunk ends with the reference's sync marker
 * (endstream, src/deflator.c:609-654) and only the last chunk of a DEFLT_END
 * batch sets BFINAL.
 *
 * Workspace layout in HBM (n = batch bytes, rounded up to whole segments):
 *   prev      2 B / byte     hash chain links
 *   tok       4 B / byte     tokens, segment k at tok[k * 8192]
 *   seg_ntok  4 B / segment
 *   seg_hist  1280 B / segment
 *   blocks    sizeof(BlockInfo) / block
 *   chunks    16 B / chunk
 *   total     8 B
 *   out       compressed bytes, worst case n + n/64 + 64/block + 4 KiB
 */
#include "deflate.cuh"

extern "C" int jdb_huffman_blocks(const uint32_t*, const uint32_t*, uint64_t, uint32_t, uint32_t, uint32_t,
                                  uint32_t, uint32_t, uint32_t, const uint32_t*, void*, jdb_stream);
extern "C" size_t jdb_blockinfo_bytes(void);
extern "C" int jdb_pack_layout(void*, void*, uint32_t, uint32_t, uint32_t*, uint64_t, uint64_t*,
                               const uint32_t*, uint32_t, uint32_t, jdb_stream);
extern "C" int jdb_pack_blocks(const uint8_t*, const uint32_t*, const uint32_t*, const void*, const void*,
                               uint32_t, uint32_t, uint32_t, uint32_t*, const uint32_t*, uint32_t, uint32_t, jdb_stream);

static size_t align256(size_t v) { return (v + 255) & ~(size_t) 255; }

struct WorkLayout {
	size_t prev, heads, tok, seg_ntok, seg_hist, blocks, chunks, total, out, out_cap, bytes;
	uint32_t nseg, nchunks, bpc, nblocks;
};

static int plan(uint64_t n, const jdb_deflate_cfg* cfg, WorkLayout* L)
{
	/* a CTA of the LZ kernel takes two consecutive segments of one chunk */
	/* chunks are whole pairs of segments (an lz_kernel CTA), or exactly one segment */
	if (cfg->chunk_bytes == 0 || (cfg->chunk_bytes % (2 * SEG) && cfg->chunk_bytes != SEG) || cfg->block_segs == 0 || cfg->block_segs > 16) return JDB_EARG;
	if (cfg->chunk_len && (n % cfg->chunk_bytes || cfg->dict_region || cfg->chunk_bytes > CHUNK_LEN_MASK)) return JDB_EARG;
	const uint64_t nseg = (n + SEG - 1) / SEG;
	const uint64_t nchunks = n ? (n + cfg->chunk_bytes - 1) / cfg->chunk_bytes : 1;
	const uint32_t spc = cfg->chunk_bytes / SEG;
	const uint32_t bpc = (spc + cfg->block_segs - 1) / cfg->block_segs;
	const uint64_t nblocks = nchunks * bpc;
	if (nseg > 0xfffffff0ull || nblocks > 0x7fffffffull) return JDB_EARG;
	const size_t npad = (size_t) (nseg ? nseg : 1) * SEG + 64;
	size_t off = 0;
	L->prev = off;      off += align256(npad * 2);
	L->heads = off;     off += align256(jdb_lz_chain_heads_bytes(n, cfg->chunk_bytes));
	L->tok = off;       off += align256(npad * 4);
	L->seg_ntok = off;  off += align256(((size_t) nchunks * spc + 16) * 4);
	L->seg_hist = off;  off += align256(((size_t) nchunks * spc + 1) * NSYM * 4);
	L->blocks = off;    off += align256((size_t) nblocks * jdb_blockinfo_bytes());
	L->chunks = off;    off += align256((size_t) nchunks * sizeof(ChunkInfo));
	L->total = off;     off += 256;
	L->out = off;
	L->out_cap = align256((size_t) n + (size_t) n / 64 + (size_t) nblocks * 64 + 4096 +
	                      (cfg->chunk_len ? (size_t) nchunks * (cfg->wrap_head + cfg->wrap_tail) : 0));
	off += L->out_cap;
	L->bytes = off;
	L->nseg = (uint32_t) nseg;
	L->nchunks = (uint32_t) nchunks;
	L->bpc = bpc;
	L->nblocks = (uint32_t) nblocks;
	return JDB_OK;
}

extern "C" size_t jdb_deflate_workspace_bytes(uint64_t n, const jdb_deflate_cfg* cfg)
{
	WorkLayout L;
	if (plan(n, cfg, &L) != JDB_OK) return 0;
	return L.bytes;
}

/* the ChunkInfo table (compressed size and output offset of every chunk) inside `work`,
 * valid once the pipeline has run on the stream; used by records.cu */
extern "C" const void* jdb_deflate_chunk_table(uint64_t n, const jdb_deflate_cfg* cfg, const void* work)
{
	WorkLayout L;
	if (plan(n, cfg, &L) != JDB_OK) return NULL;
	return (const uint8_t*) work + L.chunks;
}

extern "C" int jdb_deflate_run(const uint8_t* in, uint64_t n, const jdb_deflate_cfg* cfg,
                               void* work, uint8_t** out, uint64_t** total_dev, jdb_stream s)
{
	WorkLayout L;
	int r = plan(n, cfg, &L);
	if (r != JDB_OK) return r;
	uint8_t* w = (uint8_t*) work;
	uint16_t* prev = (uint16_t*) (w + L.prev);
	uint16_t* heads = (uint16_t*) (w + L.heads);
	uint32_t* tok = (uint32_t*) (w + L.tok);
	uint32_t* seg_ntok = (uint32_t*) (w + L.seg_ntok);
	uint32_t* seg_hist = (uint32_t*) (w + L.seg_hist);
	void* blocks = w + L.blocks;
	void* chunks = w + L.chunks;
	uint64_t* total = (uint64_t*) (w + L.total);
	uint32_t* outw = (uint32_t*) (w + L.out);

	if (cfg->level == 0 || n == 0) {
		/* stored only: no tokens, empty histograms */
		r = jdb_memset_async(seg_ntok, 0, L.seg_hist - L.seg_ntok, s);
		if (r == JDB_OK) r = jdb_memset_async(seg_hist, 0, L.blocks - L.seg_hist, s);
		if (r != JDB_OK) return r;
	} else {
		/* segment indexing is chunk relative (a chunk owns chunk_bytes/SEG slots) only
		 * when the chunk size is a multiple of SEG, which plan() enforces, so the
		 * global segment index k*SEG is the same thing */
		/* positions per chain-building warp: a whole chunk when the batch has enough
		 * chunks to fill the GPU (3 resident warps per SM, several waves), otherwise
		 * sub-ranges that each replay 32 KiB of history first */
		uint32_t range = cfg->chain_range;
		if (range == 0) {
			/* (no range replays history any more -- chain_fix_kernel -- so shorter ranges cost
			 * nothing but their head table: six resident CTAs per SM, four waves) */
			const uint64_t want = (uint64_t) jdb_rt_sm_count() * 6 * 4;
			range = cfg->chunk_bytes;
			while (range > 65536 && (range & 1) == 0 && (range / 2) % SEG == 0 && (n + range - 1) / range < want) range /= 2;
		}
		if (range > cfg->chunk_bytes || cfg->chunk_bytes % range || range % SEG) range = cfg->chunk_bytes;
		r = jdb_lz_chain(in, n, cfg->chunk_bytes, range, cfg->chunk_len, prev, range >= 65536 ? heads : (uint16_t*) 0, s);   /* (the scratch is sized for ranges of 64 KiB and more) */
		if (r != JDB_OK) return r;
		r = jdb_lz_parse(in, n, cfg->chunk_bytes, cfg->chunk_len, prev, cfg->good, cfg->nice, cfg->chain, cfg->lazy,
		                 cfg->dict_region / SEG, cfg->dict_pad, tok, seg_ntok, seg_hist, s);
		if (r != JDB_OK) return r;
	}
	r = jdb_huffman_blocks(seg_ntok, seg_hist, n, cfg->chunk_bytes, cfg->block_segs, L.nblocks,
	                       cfg->level, cfg->fixedonly, cfg->dict_region / (cfg->block_segs * SEG), cfg->chunk_len, blocks, s);
	if (r != JDB_OK) return r;
	r = jdb_pack_layout(blocks, chunks, L.nchunks, L.bpc, outw, L.out_cap / 4, total,
	                    cfg->chunk_len, cfg->wrap_head, cfg->wrap_tail, s);
	if (r != JDB_OK) return r;
	r = jdb_pack_blocks(in, tok, seg_ntok, blocks, chunks, L.bpc, L.nchunks, cfg->final, outw,
	                    cfg->chunk_len, cfg->wrap_head, cfg->wrap_tail, s);
	if (r != JDB_OK) return r;
	*out = w + L.out;
	*total_dev = total;
	return JDB_OK;
}
