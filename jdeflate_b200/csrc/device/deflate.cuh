/* deflate.cuh -- data layout shared by the deflate pipeline kernels
 * (lz.cu, huffman.cu, pack.cu, deflate.cu). */
#ifndef JDB_DEFLATE_CUH
#define JDB_DEFLATE_CUH

#include "common.cuh"

/*
 * Units (all sizes in input bytes):
 *   chunk    independent unit of the output stream: no match crosses a chunk
 *            start and every chunk ends in a byte aligned empty stored block
 *            (reference endstream(), src/deflator.c:609-654), so chunk outputs
 *            concatenate into one valid stream.  Multiple of SEG.
 *   segment  16 KiB of positions handled by one CTA of the LZ kernel
 *            (match search + parse); also the histogram granule.
 *   block    one DEFLATE block = up to `block_segs` consecutive segments of a
 *            chunk, one Huffman code set.
 */
#define SEG            8192u
#define SEG_SHIFT      13
#define WND            32768u
#define MINLEN         4u          /* hash-4 chains find matches of >= 4 bytes */
#define MAXLEN         258u
#define NSYM           320u        /* 288 lit/len slots + 32 distance slots    */
#define DSYM0          288u

/* token: literal = byte value; match = TOK_MATCH | (len-3)<<16 | (dist-1) */
#define TOK_MATCH      0x80000000u

/* per-position search result: len<<16 | dist, bit 31 = "parser takes it" */
#define M_TAKE         0x80000000u

#define HDR_WORDS      160u        /* dynamic header: <= 17+57+316*14 = 4498 bits */

enum { BT_STORED = 0, BT_FIXED = 1, BT_DYNAMIC = 2 };

struct BlockInfo {
	uint32_t first_seg;      /* global segment index */
	uint32_t nsegs;
	uint32_t in_len;         /* input bytes covered */
	uint32_t ntok;
	uint32_t type;           /* BT_* */
	uint32_t last_in_chunk;
	uint32_t hdr_bits;       /* block header incl. the 3 type bits (fixed/dynamic) */
	uint32_t pad0;
	uint64_t body_bits;      /* header + symbols + EOB for fixed/dynamic */
	uint64_t bit_off;        /* start, in bits, relative to the chunk's first byte */
	uint64_t in_off;         /* first input byte, relative to the batch */
	uint32_t hdr[HDR_WORDS]; /* header bit string, LSB first */
	uint32_t code[NSYM];     /* bit-reversed code | length << 16 */
};

/*
 * Ragged chunks (batched compression of independent records, jdb200_deflate_batch):
 * chunk c occupies the slot [c * chunk_bytes, (c + 1) * chunk_bytes) of the batch but
 * only the first chunk_len[c] & CHUNK_LEN_MASK bytes are data.  A record is a run of
 * chunks, all full but the last; CHUNK_FIRST / CHUNK_LAST mark its ends (the last one
 * closes the record's stream with BFINAL).  chunk_len == NULL is the plain contiguous
 * batch.
 */
#define CHUNK_LEN_MASK 0x0fffffffu
#define CHUNK_FIRST    0x40000000u
#define CHUNK_LAST     0x80000000u

static __device__ __forceinline__ uint64_t
chunk_end(const uint32_t* __restrict__ chunk_len, uint64_t chunk0, uint32_t chunk_bytes, uint64_t n)
{
	if (chunk_len) return chunk0 + (chunk_len[chunk0 / chunk_bytes] & CHUNK_LEN_MASK);
	const uint64_t e = chunk0 + chunk_bytes;
	return e > n ? n : e;
}

struct ChunkInfo {
	uint64_t bytes;          /* compressed size of the chunk */
	uint64_t offset;         /* exclusive scan: position in the output */
};

/* length (3..258) -> symbol index 0..28 */
static __device__ __forceinline__ uint32_t len_symbol(uint32_t len)
{
	uint32_t v = len - 3;
	if (v < 8) return v;
	if (v == 255) return 28;
	uint32_t msb = 31 - __clz(v);
	return 4 * (msb - 1) + ((v >> (msb - 2)) & 3u);
}
static __device__ __forceinline__ uint32_t len_extra_bits(uint32_t sym)
{
	return (sym < 8 || sym == 28) ? 0 : (sym >> 2) - 1;
}
static __device__ __forceinline__ uint32_t len_extra_val(uint32_t len, uint32_t xb)
{
	return (len - 3) & ((1u << xb) - 1u);
}
/* distance (1..32768) -> symbol 0..29 */
static __device__ __forceinline__ uint32_t dist_symbol(uint32_t dist)
{
	uint32_t v = dist - 1;
	if (v < 4) return v;
	uint32_t msb = 31 - __clz(v);
	return 2 * msb + ((v >> (msb - 1)) & 1u);
}
static __device__ __forceinline__ uint32_t dist_extra_bits(uint32_t sym)
{
	return sym < 4 ? 0 : (sym >> 1) - 1;
}

#endif
