/*
 * Image build: document-major shard -> term-major CSR postings in HBM.
 *
 * The reference builds its reverse index one posting at a time
 * (roaring64_bitmap_add per (term, doc), ref src/index/dtmap.c:386-438).
 * Here the whole shard is transposed on the device: every (doc, term, count)
 * becomes a (key = term, value = packed posting) pair, a stable LSD radix
 * sort on the term bits (CUB) groups them by term while keeping ascending
 * document order, and the sorted value array IS the posting array.
 */
#ifndef NXSB_GPU_BUILD_CUH
#define NXSB_GPU_BUILD_CUH

#include "common.cuh"

/*
 * One warp per document: lanes stride over its (term, count) pairs, which
 * sit contiguously in the document-major array -> coalesced 8-byte loads
 * and stores.
 */
template <bool WIDE>
__global__ void __launch_bounds__(256)
expand_pairs_kernel(const uint2 *__restrict__ pairs,
    const unsigned long long *__restrict__ doc_off,
    const uint32_t *__restrict__ doc_len, uint32_t n_docs, uint32_t n_terms,
    uint32_t *__restrict__ keys, uint2 *__restrict__ vals)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps_per_grid = (gridDim.x * blockDim.x) >> 5;
	uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;

	for (; d < n_docs; d += warps_per_grid) {
		const unsigned long long s = doc_off[d], e = doc_off[d + 1];
		const uint32_t dl = doc_len[d];

		for (unsigned long long j = s + lane; j < e; j += 32) {
			const uint2 p = pairs[j];
			const uint32_t t = p.x - 1;	// 1-based id -> index

			/* Out-of-range ids sort past every real term and are cut. */
			keys[j] = t < n_terms ? t : n_terms;
			vals[j] = make_uint2(d, WIDE ? p.y : (p.y | (dl << 16)));
		}
	}
}

/*
 * The same straight from the bytes of an `nxsdtmap` file: the
 * (term id, count) pairs of document d are doc_off[d+1] - doc_off[d]
 * big-endian u32 pairs at raw + raw_off[d] (ref src/index/storage.h:80-84,
 * 8-byte aligned: 32-byte header, 16-byte block header, 8-byte pairs), so the
 * host never converts or copies a posting.
 */
__device__ __forceinline__ uint32_t
be32(uint32_t v)
{
	return __byte_perm(v, 0, 0x0123);
}

template <bool WIDE>
__global__ void __launch_bounds__(256)
expand_raw_kernel(const unsigned char *__restrict__ raw,
    const unsigned long long *__restrict__ raw_off,
    const unsigned long long *__restrict__ doc_off,
    const uint32_t *__restrict__ doc_len, uint32_t n_docs, uint32_t n_terms,
    uint32_t *__restrict__ keys, uint2 *__restrict__ vals)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps_per_grid = (gridDim.x * blockDim.x) >> 5;
	uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;

	for (; d < n_docs; d += warps_per_grid) {
		const unsigned long long s = doc_off[d];
		const uint32_t n = (uint32_t)(doc_off[d + 1] - s);
		const uint2 *src = reinterpret_cast<const uint2 *>(raw + raw_off[d]);
		const uint32_t dl = doc_len[d];

		for (uint32_t i = lane; i < n; i += 32) {
			const uint2 p = src[i];
			const uint32_t t = be32(p.x) - 1, c = be32(p.y);

			keys[s + i] = t < n_terms ? t : n_terms;
			vals[s + i] = make_uint2(d, WIDE ? c : (c | (dl << 16)));
		}
	}
}

/* Largest count in the raw blocks (decides packed vs wide postings). */
__global__ void __launch_bounds__(256)
raw_max_count_kernel(const unsigned char *__restrict__ raw,
    const unsigned long long *__restrict__ raw_off,
    const unsigned long long *__restrict__ doc_off, uint32_t n_docs,
    uint32_t *__restrict__ max_count)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps_per_grid = (gridDim.x * blockDim.x) >> 5;
	uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	uint32_t m = 0;

	for (; d < n_docs; d += warps_per_grid) {
		const uint32_t n = (uint32_t)(doc_off[d + 1] - doc_off[d]);
		const uint2 *src = reinterpret_cast<const uint2 *>(raw + raw_off[d]);

		for (uint32_t i = lane; i < n; i += 32)
			m = max(m, be32(src[i].y));
	}
	m = __reduce_max_sync(0xffffffffu, m);
	if (lane == 0 && m > 0xffffu)
		atomicMax(max_count, m);
}

/*
 * CSR offsets from the sorted keys: off[t] = first j with keys[j] >= t.
 * Thread j owns the (possibly empty) run of terms (keys[j-1], keys[j]].
 */
__global__ void __launch_bounds__(256)
term_offsets_kernel(const uint32_t *__restrict__ keys, unsigned long long n,
    uint32_t n_terms, unsigned long long *__restrict__ off)
{
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
	unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;

	if (n == 0) {
		for (unsigned long long t = j; t <= n_terms; t += stride)
			off[t] = 0;
		return;
	}
	for (; j < n; j += stride) {
		const uint32_t hi = keys[j];
		const uint32_t lo = j ? keys[j - 1] + 1 : 0;

		for (uint32_t t = lo; t <= hi && t <= n_terms; t++)
			off[t] = j;
		if (j == n - 1) {
			for (uint32_t t = hi + 1; t <= n_terms; t++)
				off[t] = n;
		}
	}
}

__global__ void __launch_bounds__(256)
local_df_kernel(const unsigned long long *__restrict__ off, uint32_t n_terms,
    uint32_t *__restrict__ df)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;

	if (t < n_terms)
		df[t] = (uint32_t)(off[t + 1] - off[t]);
}

/*
 * Slice boundaries of one posting list: row[j] = index (relative to the
 * list) of the first posting whose document lies in tile >= j, for
 * j in [0, ntiles]; row[ntiles] = df.  Each thread looks at one posting and
 * its predecessor and writes the boundaries that fall between them.
 */
__device__ __forceinline__ void
fill_skip_row(const uint2 *__restrict__ list, uint32_t df, uint32_t ntiles,
    uint32_t *__restrict__ row, uint32_t shift = TILE_SHIFT)
{
	if (df == 0) {
		for (uint32_t j = threadIdx.x; j <= ntiles; j += blockDim.x)
			row[j] = 0;
		return;
	}
	for (uint32_t i = threadIdx.x; i < df; i += blockDim.x) {
		const uint32_t tile = list[i].x >> shift;
		const uint32_t lo = i ? (list[i - 1].x >> shift) + 1 : 0;

		for (uint32_t j = lo; j <= tile; j++)
			row[j] = i;
		if (i == df - 1) {
			for (uint32_t j = tile + 1; j <= ntiles; j++)
				row[j] = df;
		}
	}
}

/*
 * Permanent rows: one block per long term (rows[r] = term index); tiles of
 * 2^shift documents (TILE_SHIFT for the scorers' tiles, MT_SHIFT for the
 * mini-tiles of the block scorer).
 */
__global__ void __launch_bounds__(256)
build_skip_rows_kernel(const uint2 *__restrict__ post,
    const unsigned long long *__restrict__ term_off,
    const uint32_t *__restrict__ long_terms, uint32_t ntiles,
    uint32_t *__restrict__ skip, uint32_t shift)
{
	const uint32_t t = long_terms[blockIdx.x];
	const unsigned long long s = term_off[t];

	fill_skip_row(post + s, (uint32_t)(term_off[t + 1] - s), ntiles,
	    skip + (size_t)blockIdx.x * (ntiles + 1), shift);
}

/*
 * Dense columns: block b scatters the postings of dense term cols[b] into
 * column b (pre-zeroed): col[doc] = packed word.
 */
__global__ void __launch_bounds__(256)
build_dense_columns_kernel(const uint2 *__restrict__ post,
    const unsigned long long *__restrict__ term_off,
    const uint32_t *__restrict__ cols, unsigned long long col_words,
    uint32_t *__restrict__ dense)
{
	const uint32_t t = cols[blockIdx.y];
	const unsigned long long s = term_off[t], e = term_off[t + 1];
	uint32_t *col = dense + (unsigned long long)blockIdx.y * col_words;

	for (unsigned long long i = s + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	    i < e; i += (unsigned long long)gridDim.x * blockDim.x) {
		const uint2 p = post[i];

		col[p.x] = p.y;
	}
}

#endif
