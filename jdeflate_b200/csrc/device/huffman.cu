/*
 * huffman.cu -- per-block Huffman stage of the deflate pipeline for sm_100a.
 *
 * Replaces flushblock's code construction in the reference: buildtables ->
 * setuptable -> computelengths / heapsort / katajainen / limitlengths
 * (src/deflator.c:933-1285, 1361-1390), the code length run-length coding
 * countprecodes (:1287-1354) and the header layout of emittrees (:1633-1722).
 *
 * One warp per DEFLATE block:
 *   1. sum the 320-bin histograms of the block's segments (+1 end-of-block);
 *   2. sort the used symbols by (frequency, symbol) with a warp bitonic sort
 *      (the order the reference's heapsort produces);
 *   3. in-place Moffat-Katajainen minimum-redundancy lengths, clamped to 15
 *      (7 for the precode) with the same Kraft repair as the reference, so a
 *      given histogram gets the reference's code lengths;
 *   4. canonical, bit-reversed codes;
 *   5. run-length code the lit/len + distance lengths, build the precode,
 *      assemble the dynamic header bit string;
 *   6. price the block three ways -- dynamic, fixed, stored -- and keep the
 *      cheapest (the reference never compares and never stores at level >= 1,
 *      src/deflator.c:1752-1760; DESIGN.md deviation 6).
 * Output: one BlockInfo (type, header bits, code table, size in bits).
 */
#include "deflate.cuh"

#define HF_WARPS    4
#define HF_THREADS  (HF_WARPS * 32)

struct HfSmem {
	uint32_t freq[NSYM];
	uint32_t key[512];
	uint32_t work[288];
	uint8_t  len[NSYM + 8];      /* code lengths: lit/len at 0.., distance at 288.. */
	uint16_t rle[NSYM + 8];      /* precode symbol | extra value << 8               */
	uint32_t pfreq[19];
	uint8_t  plen[19];
	uint16_t pcode[19];
	uint32_t count[16];
	uint32_t next[16];
	uint32_t hdr[HDR_WORDS];
	uint32_t red[4];
};

__constant__ uint8_t c_pre_order[19] = {
	16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15
};

static __device__ __forceinline__ uint32_t rev_bits(uint32_t code, uint32_t len)
{
	return __brev(code) >> (32 - len);
}

static __device__ __forceinline__ uint32_t fixed_len(uint32_t s)
{
	return s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
}

static __device__ __forceinline__ uint32_t fixed_code(uint32_t s)
{
	return s < 144 ? 0x30 + s : s < 256 ? 0x190 + (s - 144) : s < 280 ? s - 256 : 0xc0 + (s - 280);
}

/* extra bits carried by a lit/len slot (0..287) or a distance slot (288..) */
static __device__ __forceinline__ uint32_t slot_extra(uint32_t s)
{
	if (s < 257) return 0;
	if (s < DSYM0) return s <= 285 ? len_extra_bits(s - 257) : 0;
	return dist_extra_bits(s - DSYM0);
}

/* warp bitonic sort of key[0..N) ascending, N a power of two >= 32 */
static __device__ void
sort_keys(uint32_t* key, uint32_t N)
{
	const unsigned lane = jdb_lane();
	for (uint32_t k = 2; k <= N; k <<= 1) {
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			for (uint32_t t = lane; t < N / 2; t += 32) {
				uint32_t i = ((t / j) * (j << 1)) + (t % j);
				uint32_t l = i + j;
				uint32_t a = key[i], b = key[l];
				bool up = (i & k) == 0;
				if ((a > b) == up) { key[i] = b; key[l] = a; }
			}
			__syncwarp();
		}
	}
}

/* Moffat-Katajainen in place: a[0..n) ascending frequencies -> code lengths
 * (lane 0 only).  src/deflator.c:1032-1081 */
static __device__ void
mr_lengths(uint32_t* a, int n)
{
	if (n == 1) { a[0] = 1; return; }
	int root = 0, leaf = 0, next;
	for (next = 0; next < n - 1; next++) {
		if (leaf >= n || (root < next && a[root] < a[leaf])) { a[next] = a[root]; a[root++] = (uint32_t) next; }
		else a[next] = a[leaf++];
		if (leaf >= n || (root < next && a[root] < a[leaf])) { a[next] += a[root]; a[root++] = (uint32_t) next; }
		else a[next] += a[leaf++];
	}
	int prev = n - 2, tree = n - 2, k = n - 1, avail = 2;
	for (int depth = 1; k > 0; depth++) {
		int used = 0;
		while (tree && a[tree - 1] >= (uint32_t) prev) { tree--; used++; }
		for (int j = avail - used; j; j--) a[k--] = (uint32_t) depth;
		avail = used << 1;
		prev = tree;
	}
}

/* clamp + Kraft repair (lane 0 only).  src/deflator.c:991-1028 */
static __device__ void
limit_lengths(uint32_t* len, int n, uint32_t maxlen)
{
	int k = 0;
	for (int i = 0; i < n; i++) {
		if (len[i] > maxlen) len[i] = maxlen;
		k += 1 << (15 - len[i]);
	}
	for (int i = 0; i < n; i++) {
		while (len[i] < maxlen && k > 0x8000) { len[i]++; k -= 1 << (15 - len[i]); }
	}
	for (int i = n - 1; i >= 0; i--) {
		while (k + (1 << (15 - len[i])) <= 0x8000) { k += 1 << (15 - len[i]); len[i]--; }
	}
}

/*
 * Code lengths for the `n` symbols whose frequencies sit at freq[0..n):
 * result in len[0..n).  All lanes call; needs n <= 288.
 */
static __device__ void
make_lengths(HfSmem& S, uint32_t* freq, uint8_t* len, int n, uint32_t maxlen)
{
	const unsigned lane = jdb_lane();
	/* at least two codes (src/deflator.c:1149-1162) */
	if (lane == 0) {
		int used = 0;
		for (int i = 0; i < n; i++) used += freq[i] != 0;
		if (used == 0) { freq[0] = 1; freq[1] = 1; }
		else if (used == 1) { if (freq[0]) freq[1] = 1; else freq[0] = 1; }
	}
	__syncwarp();
	/* only the symbols that occur are sorted: compact their keys, pad to a power of two
	 * (a small block uses a fraction of the 286 symbols; sorting 512 slots three times per
	 * block was the larger part of this kernel on batches of small records) */
	uint32_t nused = 0;
	for (int i0 = 0; i0 < n; i0 += 32) {
		const int i = i0 + (int) lane;
		const bool on = i < n && freq[i] != 0;
		const unsigned b = __ballot_sync(JDB_FULL_MASK, on);
		if (on) S.key[nused + __popc(b & ((1u << lane) - 1u))] = (freq[i] << 9) | (uint32_t) i;
		nused += __popc(b);
	}
	uint32_t N = 32;
	while (N < nused) N <<= 1;
	for (uint32_t i = nused + lane; i < N; i += 32) S.key[i] = 0xffffffffu;
	for (int i = lane; i < n; i += 32) len[i] = 0;
	__syncwarp();
	sort_keys(S.key, N);
	if (lane == 0) {
		const int used = (int) nused;
		for (int i = 0; i < used; i++) S.work[i] = S.key[i] >> 9;
		mr_lengths(S.work, used);
		limit_lengths(S.work, used, maxlen);
		for (int i = 0; i < used; i++) len[S.key[i] & 511u] = (uint8_t) S.work[i];
	}
	__syncwarp();
}

/* canonical codes for len[0..n) -> out[i] = reversed code | len << 16.  All lanes call:
 * counts per length by ballots, then every symbol's code is the first code of its length
 * plus the number of earlier symbols with the same length (32 symbols per step). */
static __device__ void
assign_codes(HfSmem& S, const uint8_t* len, int n, uint32_t* out)
{
	const unsigned lane = jdb_lane();
	if (lane < 16) S.count[lane] = 0;
	__syncwarp();
	for (int i0 = 0; i0 < n; i0 += 32) {
		const int i = i0 + (int) lane;
		const uint32_t l = i < n ? len[i] : 0u;
		for (uint32_t q = 1; q <= 15; q++) {
			const unsigned b = __ballot_sync(JDB_FULL_MASK, l == q);
			if (lane == 0 && b) S.count[q] += __popc(b);
		}
	}
	__syncwarp();
	if (lane == 0) {
		uint32_t c = 0;
		S.next[0] = 0;
		for (int l = 1; l <= 15; l++) {
			c = (c + (l > 1 ? S.count[l - 1] : 0u)) << 1;
			S.next[l] = c;
		}
	}
	__syncwarp();
	for (int i0 = 0; i0 < n; i0 += 32) {
		const int i = i0 + (int) lane;
		const uint32_t l = i < n ? len[i] : 0u;
		uint32_t rank = 0, same = 0;
		for (uint32_t q = 1; q <= 15; q++) {
			const unsigned b = __ballot_sync(JDB_FULL_MASK, l == q);
			if (l == q) { rank = __popc(b & ((1u << lane) - 1u)); same = b; }
		}
		uint32_t code = 0;
		if (l) code = S.next[l] + rank;
		__syncwarp();
		/* the last lane of every length group moves the group's counter on */
		if (l && (same >> lane) <= 1u) S.next[l] = code + 1;
		__syncwarp();
		if (i < n) out[i] = l ? rev_bits(code, l) | (l << 16) : 0;
	}
}

struct HdrWriter {
	uint32_t* w;
	uint32_t bits;
	__device__ void put(uint32_t v, uint32_t n)
	{
		if (n == 0) return;
		uint32_t idx = bits >> 5, sh = bits & 31;
		w[idx] |= v << sh;
		if (sh + n > 32) w[idx + 1] |= v >> (32 - sh);
		bits += n;
	}
};

__global__ void __launch_bounds__(HF_THREADS)
huffman_kernel(const uint32_t* __restrict__ seg_ntok, const uint32_t* __restrict__ seg_hist,
               uint64_t n, uint32_t chunk_bytes, uint32_t block_segs, uint32_t nblocks,
               uint32_t level, uint32_t fixedonly, uint32_t skip_blocks,
               const uint32_t* __restrict__ chunk_len, BlockInfo* __restrict__ blocks)
{
	__shared__ HfSmem smem[HF_WARPS];
	HfSmem& S = smem[jdb_warp()];
	const unsigned lane = jdb_lane();
	const uint32_t b = blockIdx.x * HF_WARPS + jdb_warp();
	if (b >= nblocks) return;

	const uint32_t segs_per_chunk = chunk_bytes / SEG;
	const uint32_t bpc = (segs_per_chunk + block_segs - 1) / block_segs;      /* blocks per chunk */
	const uint32_t chunk = b / bpc, k = b % bpc;
	const uint64_t chunk0 = (uint64_t) chunk * chunk_bytes;
	const uint64_t chunk1 = chunk_end(chunk_len, chunk0, chunk_bytes, n);
	const uint32_t csegs = (uint32_t) ((chunk1 - chunk0 + SEG - 1) / SEG);    /* segments in this chunk */
	const uint32_t s0 = k * block_segs;
	uint32_t ns = s0 < csegs ? csegs - s0 : 0;
	if (ns > block_segs) ns = block_segs;
	if (b < skip_blocks) ns = 0;                 /* preset dictionary: history only, no block */
	const uint32_t gseg0 = chunk * segs_per_chunk + s0;
	const uint64_t in_off = chunk0 + (uint64_t) s0 * SEG;
	uint64_t in_end = in_off + (uint64_t) ns * SEG;
	if (in_end > chunk1) in_end = chunk1;
	const uint32_t in_len = ns ? (uint32_t) (in_end - in_off) : 0;
	const bool last = ns && (s0 + ns >= csegs);

	BlockInfo& B = blocks[b];
	if (ns == 0) {
		if (lane == 0) {
			B.first_seg = gseg0; B.nsegs = 0; B.in_len = 0; B.ntok = 0; B.type = BT_FIXED;
			B.last_in_chunk = 0; B.hdr_bits = 0; B.body_bits = 0; B.bit_off = 0; B.in_off = in_off;
		}
		return;
	}

	/* 1. histogram of the block */
	uint32_t ntok = 0;
	for (uint32_t i = lane; i < NSYM; i += 32) {
		uint32_t f = 0;
		for (uint32_t s = 0; s < ns; s++) f += seg_hist[(uint64_t) (gseg0 + s) * NSYM + i];
		S.freq[i] = f;
	}
	for (uint32_t s = lane; s < ns; s += 32) ntok += seg_ntok[gseg0 + s];
	for (int o = 16; o; o >>= 1) ntok += __shfl_xor_sync(JDB_FULL_MASK, ntok, o);
	for (uint32_t i = lane; i < HDR_WORDS; i += 32) S.hdr[i] = 0;
	__syncwarp();
	if (lane == 0) S.freq[256] += 1;
	__syncwarp();

	/* price of the fixed code and of the extra bits (before freq is clobbered) */
	uint64_t fixed_bits = 0, extra_bits = 0;
	for (uint32_t i = lane; i < NSYM; i += 32) {
		const uint32_t f = S.freq[i];
		extra_bits += (uint64_t) f * slot_extra(i);
		fixed_bits += (uint64_t) f * (i < DSYM0 ? fixed_len(i) : 5u);
	}
	for (int o = 16; o; o >>= 1) {
		fixed_bits += __shfl_xor_sync(JDB_FULL_MASK, fixed_bits, o);
		extra_bits += __shfl_xor_sync(JDB_FULL_MASK, extra_bits, o);
	}
	fixed_bits += extra_bits + 3;

	/* 2-4. dynamic code lengths (symbols 286/287 and 30/31 never occur) */
	make_lengths(S, S.freq, S.len, 286, 15);
	/* the distance histogram is kept for pricing: lengths go to len[288..] */
	uint64_t dyn_bits = 0;
	{
		/* body cost needs the frequencies again: re-read them from global */
		for (uint32_t i = lane; i < 286; i += 32) {
			uint32_t f = (i == 256);
			for (uint32_t s = 0; s < ns; s++) f += seg_hist[(uint64_t) (gseg0 + s) * NSYM + i];
			dyn_bits += (uint64_t) f * S.len[i];
		}
	}
	make_lengths(S, S.freq + DSYM0, S.len + DSYM0, 30, 15);
	for (uint32_t i = lane; i < 30; i += 32) {
		uint32_t f = 0;
		for (uint32_t s = 0; s < ns; s++) f += seg_hist[(uint64_t) (gseg0 + s) * NSYM + DSYM0 + i];
		dyn_bits += (uint64_t) f * S.len[DSYM0 + i];
	}
	for (int o = 16; o; o >>= 1) dyn_bits += __shfl_xor_sync(JDB_FULL_MASK, dyn_bits, o);
	dyn_bits += extra_bits;

	/* 5. header: run-length code the hlit + hdist lengths as one sequence */
	uint32_t hdr_bits = 0;
	if (lane == 0) {
		int hlit = 286, hdist = 30;
		while (hlit > 257 && S.len[hlit - 1] == 0) hlit--;
		while (hdist > 1 && S.len[DSYM0 + hdist - 1] == 0) hdist--;
		/* gather into work[] as one array */
		int total = hlit + hdist;
		for (int i = 0; i < hlit; i++) S.work[i] = S.len[i];
		/* work has 288 slots: hlit + hdist <= 316 does not fit -> use key[] */
		uint32_t* seq = S.key;
		for (int i = 0; i < hlit; i++) seq[i] = S.len[i];
		for (int i = 0; i < hdist; i++) seq[hlit + i] = S.len[DSYM0 + i];
		for (int i = 0; i < 19; i++) S.pfreq[i] = 0;
		int nr = 0;
		for (int i = 0; i < total;) {
			uint32_t v = seq[i];
			int run = 1;
			while (i + run < total && seq[i + run] == v) run++;
			i += run;
			if (v == 0) {
				while (run >= 11) { int r = run > 138 ? 138 : run; S.rle[nr++] = (uint16_t) (18 | ((r - 11) << 8)); S.pfreq[18]++; run -= r; }
				if (run >= 3) { S.rle[nr++] = (uint16_t) (17 | ((run - 3) << 8)); S.pfreq[17]++; run = 0; }
				while (run-- > 0) { S.rle[nr++] = 0; S.pfreq[0]++; }
			} else {
				S.rle[nr++] = (uint16_t) v; S.pfreq[v]++; run--;
				while (run >= 3) { int r = run > 6 ? 6 : run; S.rle[nr++] = (uint16_t) (16 | ((r - 3) << 8)); S.pfreq[16]++; run -= r; }
				while (run-- > 0) { S.rle[nr++] = (uint16_t) v; S.pfreq[v]++; }
			}
		}
		S.red[0] = (uint32_t) nr;
		S.red[1] = (uint32_t) hlit;
		S.red[2] = (uint32_t) hdist;
	}
	__syncwarp();
	{
		/* precode lengths (limit 7): tiny, reuse make_lengths on a scratch copy */
		uint32_t* pf = S.work + 256;          /* 19 slots inside work[] tail */
		if (lane < 19) pf[lane] = S.pfreq[lane];
		__syncwarp();
		uint8_t* pl = S.plen;
		make_lengths(S, pf, pl, 19, 7);
	}
	assign_codes(S, S.plen, 19, S.work);
	__syncwarp();
	if (lane < 19) S.pcode[lane] = (uint16_t) (S.work[lane] & 0xffffu);
	__syncwarp();
	if (lane == 0) {
		int hclen = 19;
		while (hclen > 4 && S.plen[c_pre_order[hclen - 1]] == 0) hclen--;
		const int nr = (int) S.red[0], hlit = (int) S.red[1], hdist = (int) S.red[2];
		HdrWriter hw;
		hw.w = S.hdr; hw.bits = 0;
		hw.put(0, 1);                      /* BFINAL: only the closing marker sets it */
		hw.put(BT_DYNAMIC, 2);
		hw.put((uint32_t) hlit - 257, 5);
		hw.put((uint32_t) hdist - 1, 5);
		hw.put((uint32_t) hclen - 4, 4);
		for (int i = 0; i < hclen; i++) hw.put(S.plen[c_pre_order[i]], 3);
		for (int i = 0; i < nr; i++) {
			uint32_t sym = S.rle[i] & 0xffu, xv = S.rle[i] >> 8;
			hw.put(S.pcode[sym], S.plen[sym]);
			if (sym == 16) hw.put(xv, 2);
			else if (sym == 17) hw.put(xv, 3);
			else if (sym == 18) hw.put(xv, 7);
		}
		S.red[3] = hw.bits;
	}
	__syncwarp();
	hdr_bits = S.red[3];
	dyn_bits += hdr_bits;

	/* 6. choose */
	const uint32_t pieces = (in_len + 65534u) / 65535u;
	const uint64_t stored_bits = (uint64_t) in_len * 8 + (uint64_t) pieces * 40 + 8;
	uint32_t type = BT_DYNAMIC;
	uint64_t best = dyn_bits;
	if (fixed_bits <= best) { type = BT_FIXED; best = fixed_bits; }
	if (fixedonly) { type = BT_FIXED; best = fixed_bits; }
	else if (stored_bits < best || level == 0) { type = BT_STORED; best = stored_bits; }

	/* 7. publish */
	if (type == BT_DYNAMIC) {
		assign_codes(S, S.len, 286, B.code);
		assign_codes(S, S.len + DSYM0, 30, B.code + DSYM0);
		for (uint32_t i = lane; i < HDR_WORDS; i += 32) B.hdr[i] = S.hdr[i];
	} else if (type == BT_FIXED) {
		for (uint32_t i = lane; i < DSYM0; i += 32) B.code[i] = rev_bits(fixed_code(i), fixed_len(i)) | (fixed_len(i) << 16);
		B.code[DSYM0 + lane] = rev_bits(lane, 5) | (5u << 16);
		if (lane == 0) B.hdr[0] = BT_FIXED << 1;     /* BFINAL 0, BTYPE 01 */
	}
	if (lane == 0) {
		B.first_seg = gseg0;
		B.nsegs = ns;
		B.in_len = in_len;
		B.ntok = ntok;
		B.type = type;
		B.last_in_chunk = last ? 1u : 0u;
		B.hdr_bits = type == BT_DYNAMIC ? hdr_bits : type == BT_FIXED ? 3u : 0u;
		B.body_bits = type == BT_STORED ? 0 : best;
		B.bit_off = 0;
		B.in_off = in_off;
	}
}

extern "C" int jdb_huffman_blocks(const uint32_t* seg_ntok, const uint32_t* seg_hist, uint64_t n,
                                  uint32_t chunk_bytes, uint32_t block_segs, uint32_t nblocks,
                                  uint32_t level, uint32_t fixedonly, uint32_t skip_blocks,
                                  const uint32_t* chunk_len, void* blocks, jdb_stream s)
{
	if (nblocks == 0) return JDB_OK;
	JDB_LAUNCH(huffman_kernel, dim3((nblocks + HF_WARPS - 1) / HF_WARPS), dim3(HF_THREADS), 0, s,
	           seg_ntok, seg_hist, n, chunk_bytes, block_segs, nblocks, level, fixedonly, skip_blocks, chunk_len, (BlockInfo*) blocks);
	return jdb_rt_check_launch("huffman_kernel");
}

extern "C" size_t jdb_blockinfo_bytes(void) { return sizeof(BlockInfo); }
