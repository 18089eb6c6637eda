/*
 * pack.cu -- output placement and bit packing of the deflate pipeline.
 *
 * Replaces the reference's serial bit emitter: putbits/tryemitbits
 * (src/deflator.c:563-607), emittrees (:1633-1722), emitlz / emitlzfast
 * (:1421-1631), the stored-block writer of compress0 (:796-926) and the
 * sync / final marker endstream (:609-654).
 *
 *  layout_kernel  (one CTA)  walks the blocks of every chunk to give each
 *                 block its starting bit (stored blocks need the running
 *                 alignment), closes every chunk with the byte aligned marker,
 *                 exclusive-scans the chunk sizes so that all chunks land
 *                 contiguously in the output, and zeroes the one 32-bit word
 *                 at every block start (the only words two CTAs share).
 *
 *  pack_kernel    one CTA per block.  Every thread turns tokens into
 *                 (bits, nbits) pairs from the block's code table, a block
 *                 prefix sum of nbits gives each token its bit offset, the
 *                 bits are OR-ed into a shared-memory staging window and the
 *                 finished 32-bit words are written out coalesced.  Only the
 *                 first and last (partial) word of a block go out with
 *                 atomicOr.  Stored blocks are byte copies.
 *
 * Algorithmic traffic: C bytes written (+ the raw bytes re-read for stored
 * blocks); implementation traffic: 4 B per token read.
 */
#include "deflate.cuh"

#define PK_THREADS  256
#define PK_ITEMS    4
#define PK_TILE     (PK_THREADS * PK_ITEMS)
#define PK_STAGE    (PK_TILE * 48 / 32 + 4)

/* ---------------------------------------------------------------------------
 * layout
 * ------------------------------------------------------------------------- */

__global__ void __launch_bounds__(1024)
layout_kernel(BlockInfo* __restrict__ blocks, ChunkInfo* __restrict__ chunks,
              uint32_t nchunks, uint32_t bpc, uint32_t* __restrict__ out_words,
              uint64_t out_cap_words, uint64_t* __restrict__ total_out,
              const uint32_t* __restrict__ chunk_len, uint32_t wrap_head, uint32_t wrap_tail)
{
	__shared__ uint64_t warp_sum[32];
	__shared__ uint64_t carry;
	const uint32_t tid = threadIdx.x;

	/* phase 1: bit offsets inside every chunk */
	for (uint32_t c = tid; c < nchunks; c += 1024) {
		const uint32_t cl = chunk_len ? chunk_len[c] : 0u;
		/* the container header of a record sits in front of its first chunk */
		uint64_t cur = (cl & CHUNK_FIRST) ? (uint64_t) wrap_head * 8 : 0;
		for (uint32_t k = 0; k < bpc; k++) {
			BlockInfo& B = blocks[(uint64_t) c * bpc + k];
			if (B.nsegs == 0) continue;
			B.bit_off = cur;
			if (B.type == BT_STORED) {
				uint32_t left = B.in_len;
				while (left) {
					uint32_t piece = left < 65535u ? left : 65535u;
					cur += 3;
					cur = (cur + 7) & ~(uint64_t) 7;
					cur += 32 + (uint64_t) piece * 8;
					left -= piece;
				}
			} else {
				cur += B.body_bits;
			}
		}
		/* closing marker: 3 header bits, pad, 00 00 FF FF */
		cur += 3;
		cur = (cur + 7) & ~(uint64_t) 7;
		cur += 32;
		if (cl & CHUNK_LAST) cur += (uint64_t) wrap_tail * 8;
		chunks[c].bytes = cur >> 3;
	}
	if (tid == 0) carry = 0;
	__syncthreads();

	/* phase 2: exclusive scan of the chunk sizes */
	for (uint32_t base = 0; base < nchunks; base += 1024) {
		const uint32_t c = base + tid;
		uint64_t v = c < nchunks ? chunks[c].bytes : 0;
		uint64_t incl = v;
		for (int o = 1; o < 32; o <<= 1) {
			uint64_t t = __shfl_up_sync(JDB_FULL_MASK, incl, o);
			if ((int) (tid & 31) >= o) incl += t;
		}
		if ((tid & 31) == 31) warp_sum[tid >> 5] = incl;
		__syncthreads();
		if (tid < 32) {
			uint64_t w = warp_sum[tid], iw = w;
			for (int o = 1; o < 32; o <<= 1) {
				uint64_t t = __shfl_up_sync(JDB_FULL_MASK, iw, o);
				if ((int) tid >= o) iw += t;
			}
			warp_sum[tid] = iw - w;
		}
		__syncthreads();
		const uint64_t excl = carry + warp_sum[tid >> 5] + incl - v;
		if (c < nchunks) chunks[c].offset = excl;
		__syncthreads();
		if (tid == 1023) carry = excl + v;
		__syncthreads();
	}
	const uint64_t total = carry;
	if (tid == 0) *total_out = total;

	/* phase 3: zero the words shared by neighbouring blocks */
	const uint64_t nblocks = (uint64_t) nchunks * bpc;
	for (uint64_t b = tid; b < nblocks; b += 1024) {
		const BlockInfo& B = blocks[b];
		if (B.nsegs == 0) continue;
		const uint64_t bit = chunks[b / bpc].offset * 8 + B.bit_off;
		out_words[bit >> 5] = 0;
	}
	/* chunks without any block (empty input) start with their marker */
	for (uint32_t c = tid; c < nchunks; c += 1024) {
		out_words[(chunks[c].offset * 8) >> 5] = 0;
		/* the marker of a chunk may straddle into the next word */
		const uint64_t endw = ((chunks[c].offset + chunks[c].bytes) * 8) >> 5;
		if (endw < out_cap_words) out_words[endw] = 0;
		/* with a container trailer behind it the marker ends in a word of its own */
		if (chunk_len && (chunk_len[c] & CHUNK_LAST) && wrap_tail)
			out_words[((chunks[c].offset + chunks[c].bytes - wrap_tail) * 8) >> 5] = 0;
	}
}

/* ---------------------------------------------------------------------------
 * pack
 * ------------------------------------------------------------------------- */

struct PkSmem {
	uint32_t code[NSYM];
	uint32_t stage[PK_STAGE];
	uint32_t warp_sum[PK_THREADS / 32];
	uint32_t seg_tok0[16];         /* first block-token index of every segment */
	uint32_t carry_word;
	uint32_t tile_total;
};

/* flush `nitems` staged (bits, nbits) per thread; returns the new cursor */
static __device__ uint64_t
emit_tile(PkSmem& S, uint32_t* __restrict__ out_words, uint64_t first_word,
          uint64_t cur, const uint64_t* bits, const uint32_t* nbits)
{
	const uint32_t tid = threadIdx.x;
	uint32_t mine = 0;
#pragma unroll
	for (int k = 0; k < PK_ITEMS; k++) mine += nbits[k];
	uint32_t incl = mine;
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(JDB_FULL_MASK, incl, o);
		if ((int) (tid & 31) >= o) incl += t;
	}
	if ((tid & 31) == 31) S.warp_sum[tid >> 5] = incl;
	const uint64_t word0 = cur >> 5;
	for (uint32_t i = tid; i < PK_STAGE; i += PK_THREADS) S.stage[i] = 0;
	__syncthreads();
	if (tid == 0) {
		uint32_t run = 0;
		for (int w = 0; w < PK_THREADS / 32; w++) { uint32_t v = S.warp_sum[w]; S.warp_sum[w] = run; run += v; }
		S.tile_total = run;
		S.stage[0] = S.carry_word;
	}
	__syncthreads();
	uint32_t off = (uint32_t) (cur & 31) + S.warp_sum[tid >> 5] + incl - mine;
#pragma unroll
	for (int k = 0; k < PK_ITEMS; k++) {
		const uint32_t nb = nbits[k];
		if (nb) {
			const uint64_t v = bits[k];
			const uint32_t w = off >> 5, sh = off & 31;
			atomicOr(&S.stage[w], (uint32_t) (v << sh));
			if (sh + nb > 32) {
				atomicOr(&S.stage[w + 1], (uint32_t) (v >> (32 - sh)));
				if (sh + nb > 64) atomicOr(&S.stage[w + 2], (uint32_t) (v >> (64 - sh)));
			}
			off += nb;
		}
	}
	__syncthreads();
	const uint64_t newcur = cur + S.tile_total;
	const uint32_t nfull = (uint32_t) ((newcur >> 5) - word0);
	for (uint32_t i = tid; i < nfull; i += PK_THREADS) {
		const uint64_t w = word0 + i;
		if (w == first_word) atomicOr(&out_words[w], S.stage[i]);
		else out_words[w] = S.stage[i];
	}
	__syncthreads();
	if (tid == 0) S.carry_word = (newcur & 31) ? S.stage[nfull] : 0;
	__syncthreads();
	return newcur;
}

__global__ void __launch_bounds__(PK_THREADS)
pack_kernel(const uint8_t* __restrict__ in, const uint32_t* __restrict__ tok,
            const uint32_t* __restrict__ seg_ntok, const BlockInfo* __restrict__ blocks,
            const ChunkInfo* __restrict__ chunks, uint32_t bpc, uint32_t nchunks,
            uint32_t final_stream, uint32_t* __restrict__ out_words,
            const uint32_t* __restrict__ chunk_len, uint32_t wrap_head, uint32_t wrap_tail)
{
	__shared__ PkSmem S;
	const uint32_t tid = threadIdx.x;
	const uint32_t b = blockIdx.x;
	const BlockInfo& B = blocks[b];
	const uint32_t chunk = b / bpc;
	uint8_t* out8 = (uint8_t*) out_words;

	/* a chunk without blocks (empty input): its first slot writes the marker */
	const uint32_t cl = chunk_len ? chunk_len[chunk] : 0u;
	const uint32_t lead = (cl & CHUNK_FIRST) ? wrap_head : 0u;
	const uint32_t trail = (cl & CHUNK_LAST) ? wrap_tail : 0u;
	const bool empty_chunk_marker = B.nsegs == 0 && (b % bpc) == 0 && chunks[chunk].bytes == 5 + lead + trail;
	if (B.nsegs == 0 && !empty_chunk_marker) return;

	const uint64_t bit0 = chunks[chunk].offset * 8 + (B.nsegs ? B.bit_off : (uint64_t) lead * 8);
	const bool final_marker = chunk_len ? (cl & CHUNK_LAST) != 0 : (final_stream && chunk == nchunks - 1);

	if (B.nsegs == 0) {
		if (tid == 0) {
			uint8_t* p = out8 + (bit0 >> 3);
			p[0] = final_marker ? 1 : 0; p[1] = 0; p[2] = 0; p[3] = 0xff; p[4] = 0xff;
		}
		return;
	}

	if (B.type == BT_STORED) {
		/* byte copies; the first header byte may be shared with the previous
		 * block (its low bits), everything after it is owned by this CTA */
		uint64_t cur = bit0;
		uint32_t left = B.in_len;
		uint64_t src = B.in_off;
		while (left) {
			const uint32_t piece = left < 65535u ? left : 65535u;
			const uint64_t hdr_end = (cur + 3 + 7) >> 3;        /* first byte after the padded header */
			if (tid == 0) {
				/* bytes fully covered by header + padding are zero */
				for (uint64_t q = (cur >> 3) + ((cur & 7) ? 1 : 0); q < hdr_end; q++) out8[q] = 0;
				out8[hdr_end + 0] = (uint8_t) piece;
				out8[hdr_end + 1] = (uint8_t) (piece >> 8);
				out8[hdr_end + 2] = (uint8_t) ~piece;
				out8[hdr_end + 3] = (uint8_t) (~piece >> 8);
			}
			uint8_t* dst = out8 + hdr_end + 4;
			for (uint32_t i = tid; i < piece; i += PK_THREADS) dst[i] = in[src + i];
			cur = (hdr_end + 4 + piece) * 8;
			src += piece;
			left -= piece;
		}
		if (B.last_in_chunk && tid == 0) {
			uint8_t* p = out8 + (cur >> 3);
			p[0] = final_marker ? 1 : 0; p[1] = 0; p[2] = 0; p[3] = 0xff; p[4] = 0xff;
		}
		return;
	}

	/* ---- Huffman coded block ---- */
	for (uint32_t i = tid; i < NSYM; i += PK_THREADS) S.code[i] = B.code[i];
	if (tid == 0) {
		uint32_t run = 0;
		for (uint32_t s = 0; s < B.nsegs && s < 16; s++) { S.seg_tok0[s] = run; run += seg_ntok[B.first_seg + s]; }
		S.carry_word = 0;
	}
	__syncthreads();

	const uint64_t first_word = bit0 >> 5;
	uint64_t cur = bit0;
	const uint32_t nhdr = (B.hdr_bits + 31) / 32;
	const uint32_t ntok = B.ntok;
	const uint32_t nitems = nhdr + ntok + 1;                   /* + end of block */

	for (uint32_t base = 0; base < nitems; base += PK_TILE) {
		uint64_t bits[PK_ITEMS];
		uint32_t nbits[PK_ITEMS];
#pragma unroll
		for (int k = 0; k < PK_ITEMS; k++) {
			const uint32_t i = base + tid * PK_ITEMS + k;
			bits[k] = 0;
			nbits[k] = 0;
			if (i >= nitems) continue;
			if (i < nhdr) {
				const uint32_t rem = B.hdr_bits - i * 32;
				bits[k] = B.hdr[i];
				nbits[k] = rem < 32 ? rem : 32;
			} else if (i == nitems - 1) {
				const uint32_t c = S.code[256];
				bits[k] = c & 0xffffu;
				nbits[k] = c >> 16;
			} else {
				uint32_t ti = i - nhdr, s = 0;
				while (s + 1 < B.nsegs && ti >= S.seg_tok0[s + 1]) s++;
				const uint32_t t = tok[(uint64_t) (B.first_seg + s) * SEG + (ti - S.seg_tok0[s])];
				if (t & TOK_MATCH) {
					const uint32_t len = ((t >> 16) & 0xffu) + 3, dist = (t & 0x7fffu) + 1;
					const uint32_t ls = len_symbol(len), ds = dist_symbol(dist);
					const uint32_t lc = S.code[257 + ls], dc = S.code[DSYM0 + ds];
					const uint32_t lxb = len_extra_bits(ls), dxb = dist_extra_bits(ds);
					uint64_t v = lc & 0xffffu;
					uint32_t nb = lc >> 16;
					v |= (uint64_t) len_extra_val(len, lxb) << nb; nb += lxb;
					v |= (uint64_t) (dc & 0xffffu) << nb; nb += dc >> 16;
					v |= (uint64_t) ((dist - 1) & ((1u << dxb) - 1u)) << nb; nb += dxb;
					bits[k] = v;
					nbits[k] = nb;
				} else {
					const uint32_t c = S.code[t & 0xffu];
					bits[k] = c & 0xffffu;
					nbits[k] = c >> 16;
				}
			}
		}
		cur = emit_tile(S, out_words, first_word, cur, bits, nbits);
	}

	/* closing marker of the chunk: BFINAL + stored type, pad, 00 00 FF FF */
	if (B.last_in_chunk) {
		uint64_t bits[PK_ITEMS];
		uint32_t nbits[PK_ITEMS];
#pragma unroll
		for (int k = 0; k < PK_ITEMS; k++) { bits[k] = 0; nbits[k] = 0; }
		if (tid == 0) {
			const uint32_t pad = (uint32_t) ((8 - ((cur + 3) & 7)) & 7);
			bits[0] = final_marker ? 1 : 0;
			nbits[0] = 3 + pad;
			bits[1] = 0xffff0000u;
			nbits[1] = 32;
		}
		cur = emit_tile(S, out_words, first_word, cur, bits, nbits);
	}
	/* last partial word */
	if (tid == 0 && (cur & 31)) atomicOr(&out_words[cur >> 5], S.carry_word);
}

/* ---- launchers -------------------------------------------------------------- */

extern "C" int jdb_pack_layout(void* blocks, void* chunks, uint32_t nchunks, uint32_t bpc,
                               uint32_t* out_words, uint64_t out_cap_words, uint64_t* total_out,
                               const uint32_t* chunk_len, uint32_t wrap_head, uint32_t wrap_tail, jdb_stream s)
{
	JDB_LAUNCH(layout_kernel, dim3(1), dim3(1024), 0, s, (BlockInfo*) blocks, (ChunkInfo*) chunks,
	           nchunks, bpc, out_words, out_cap_words, total_out, chunk_len, wrap_head, wrap_tail);
	return jdb_rt_check_launch("layout_kernel");
}

extern "C" int jdb_pack_blocks(const uint8_t* in, const uint32_t* tok, const uint32_t* seg_ntok,
                               const void* blocks, const void* chunks, uint32_t bpc, uint32_t nchunks,
                               uint32_t final_stream, uint32_t* out_words,
                               const uint32_t* chunk_len, uint32_t wrap_head, uint32_t wrap_tail, jdb_stream s)
{
	const uint64_t nblocks = (uint64_t) nchunks * bpc;
	if (nblocks == 0) return JDB_OK;
	JDB_LAUNCH(pack_kernel, dim3((unsigned) nblocks), dim3(PK_THREADS), 0, s, in, tok, seg_ntok,
	           (const BlockInfo*) blocks, (const ChunkInfo*) chunks, bpc, nchunks, final_stream, out_words,
	           chunk_len, wrap_head, wrap_tail);
	return jdb_rt_check_launch("pack_kernel");
}
