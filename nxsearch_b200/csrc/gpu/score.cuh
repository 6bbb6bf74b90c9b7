/*
 * Query scoring kernels (sm_100a).
 *
 * The reference scores a query document-at-a-time on one core: walk the
 * result bitmap, and for every document and every query token do a bitmap
 * probe, a hashmap lookup, a binary search for the count and two libm
 * log() calls (ref src/query/search.c:235-272, src/algo/ranking.c).  Here a
 * batch of queries is scored tile-at-a-time:
 *
 *   work item = (query, tile of 16384 consecutive documents).  A CTA takes
 *   the item's slice of every token's posting list (found through the skip
 *   rows), streams it with coalesced 8-byte loads, and accumulates
 *   per-document scores into a 64 KB shared-memory accumulator -- plain
 *   read-modify-write, because a posting list holds a document at most once
 *   and tokens are processed one after the other IN TOKEN-LIST ORDER, which
 *   is also the float summation order of the reference
 *   (src/core/results.c:135-137).  Boolean operators run as word-wide set
 *   operations on per-token tile bitmaps.  A threshold-pruned radix select
 *   keeps at most `limit` candidates per tile; a second kernel merges a
 *   query's candidates into its final top-k.
 *
 * Items are handed out tile-major (all queries of a tile before the next
 * tile), so the slices of popular terms are re-read from L2, not HBM.
 */
#ifndef NXSB_GPU_SCORE_CUH
#define NXSB_GPU_SCORE_CUH

#include "common.cuh"

#define TILE_THREADS	256

struct QDesc {			// device copy of nxsb_query_t
	uint32_t	tok_off, n_tokens, prog_off, n_prog;
};

struct ScoreParams {
	const uint2 *		post;
	const DTok *		toks;
	const QDesc *		queries;
	const int32_t *		prog;
	const uint32_t *	qlist;		// queries of this launch
	uint32_t		n_q;
	uint32_t		ntiles;
	uint32_t		k;		// candidates kept per tile
	/* The three below are indexed by POSITION in qlist, not query id. */
	unsigned long long *	thr;		// pruning threshold keys
	uint32_t *		cand_count;
	unsigned long long *	cand;		// [n_q][cand_cap]
	unsigned long long	cand_cap;
	uint32_t *		work_counter;
	const float *		logtab;		// [LOGTAB_N]
	const uint32_t *	doc_len;	// WIDE mode only
	float			K0, K1;		// BM25: k(1-b), k*b/adl
	int			algo;
	uint32_t		max_tokens;	// LOGIC: bitmap rows in smem
};

/*
 * Batched term lookup: term id -> posting list, skip row, idf.
 * One thread per token instance of the batch.
 */
__global__ void __launch_bounds__(256)
resolve_tokens_kernel(const uint32_t *__restrict__ term_ids, uint32_t n,
    uint32_t n_terms, const unsigned long long *__restrict__ term_off,
    const int32_t *__restrict__ skip_row, const int32_t *__restrict__ dense_col,
    const uint32_t *__restrict__ bcol, uint32_t *__restrict__ dense_used, unsigned long long col_words, const uint32_t *__restrict__ skip,
    const uint32_t *__restrict__ skip_mt, uint32_t n_mt,
    uint32_t *__restrict__ tmp_skip, const float *__restrict__ idf,
    const float *__restrict__ wmax, const float *__restrict__ kth, uint32_t kth_step,
    const uint8_t *__restrict__ mtmax, const uint32_t *__restrict__ mtbits, uint32_t mt_stride,
    uint32_t ntiles, DTok *__restrict__ out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	DTok t;

	if (i >= n)
		return;
	t.wk = 0.f;
	t.wmax = 0.f;
	t.mtmax = nullptr;
	t.mtbits = nullptr;
	const uint32_t id = term_ids[i];

	if (id == 0 || id > n_terms) {
		/* Not a term of this index: empty list, zero row. */
		t.post_off = 0;
		t.df_local = 0;
		t.idf = 0.f;
		t.dense_off = DENSE_NONE;
		t.bcol = 0xffffffffu;
		t.skip = tmp_skip + (size_t)i * (ntiles + 1);
		t.fine = t.skip;
		t.fine_shift = TILE_SHIFT;
	} else {
		const uint32_t ti = id - 1;
		const unsigned long long s = term_off[ti];
		const int32_t row = skip_row[ti];

		t.post_off = s;
		t.df_local = (uint32_t)(term_off[ti + 1] - s);
		t.idf = idf[ti];
		t.dense_off = dense_col[ti] >= 0 ? (unsigned long long)dense_col[ti] * col_words
		    : DENSE_NONE;
		t.bcol = bcol ? bcol[ti] : 0xffffffffu;
		t.wmax = wmax ? wmax[ti] : 0.f;
		t.wk = kth ? kth[(size_t)ti * 8u + kth_step] : 0.f;
		if (dense_col[ti] >= 0)
			dense_used[dense_col[ti]] = 1u;	/* dense_scores_kernel will fill it */
		t.skip = row >= 0 ? skip + (size_t)row * (ntiles + 1)
		    : tmp_skip + (size_t)i * (ntiles + 1);
		t.fine = row >= 0 && skip_mt ? skip_mt + (size_t)row * (n_mt + 1) : t.skip;
		t.fine_shift = row >= 0 && skip_mt ? MT_SHIFT : TILE_SHIFT;
		if (row >= 0 && mtmax && t.bcol == 0xffffffffu) {
			t.mtmax = mtmax + (size_t)row * mt_stride;
			t.mtbits = mtbits ? mtbits + (size_t)row * mt_stride : nullptr;
		}
	}
	out[i] = t;
}

/* Per-batch skip rows of the short lists: one block per token instance. */
__global__ void __launch_bounds__(128)
build_temp_skips_kernel(const uint2 *__restrict__ post,
    const DTok *__restrict__ toks, uint32_t *__restrict__ tmp_skip,
    uint32_t ntiles)
{
	const DTok t = toks[blockIdx.x];
	uint32_t *row = tmp_skip + (size_t)blockIdx.x * (ntiles + 1);

	if (t.skip != row)
		return;		// long term: permanent row
	fill_skip_row(post + t.post_off, t.df_local, ntiles, row);
}

/*
 * Block-wide bitonic sort (descending) of n <= SORT_CAP keys in shared
 * memory; n is padded to a power of two with zeros by the caller.
 */
__device__ __forceinline__ void
bitonic_sort_desc(unsigned long long *s, uint32_t npow2)
{
	for (uint32_t size = 2; size <= npow2; size <<= 1) {
		for (uint32_t stride = size >> 1; stride; stride >>= 1) {
			__syncthreads();
			for (uint32_t i = threadIdx.x; i < npow2 / 2; i += blockDim.x) {
				const uint32_t lo = 2 * i - (i & (stride - 1));
				const uint32_t hi = lo + stride;
				const bool desc = (lo & size) == 0;
				const unsigned long long a = s[lo], b = s[hi];

				if ((a < b) == desc) {
					s[lo] = b;
					s[hi] = a;
				}
			}
		}
	}
	__syncthreads();
}

#include "tiles.cuh"
#include "stream.cuh"
#include "bmw.cuh"

/*
 * Final per-query top-k: merge the candidates its tiles emitted.  One CTA
 * per query.  Keys are unique (they embed the document), so the k-th largest
 * key is an exact cut.
 */
__global__ void __launch_bounds__(256)
finalize_topk_kernel(const unsigned long long *__restrict__ cand,
    unsigned long long cand_cap, const uint32_t *__restrict__ cand_count,
    const uint32_t *__restrict__ qlist, uint32_t k,
    const unsigned long long *__restrict__ doc_ids,
    Rec *__restrict__ recs, uint32_t *__restrict__ counts)
{
	__shared__ unsigned long long s_keys[SORT_CAP];
	__shared__ uint32_t s_hist[256];
	__shared__ uint32_t s_want, s_n;
	__shared__ unsigned long long s_prefix;

	const uint32_t slot = blockIdx.x, tid = threadIdx.x;
	const uint32_t q = qlist[slot];
	const unsigned long long *in = cand + (unsigned long long)slot * cand_cap;
	const uint32_t n = cand_count[slot];
	uint32_t m;		// keys to sort

	if (n <= SORT_CAP) {
		for (uint32_t i = tid; i < n; i += blockDim.x)
			s_keys[i] = in[i];
		m = n;
	} else {
		/* n > SORT_CAP >= 2k: radix-select the k-th largest key. */
		if (tid == 0) {
			s_prefix = 0;
			s_want = k;
			s_n = 0;
		}
		for (int sh = 56; sh >= 0; sh -= 8) {
			if (tid < 256)
				s_hist[tid] = 0;
			__syncthreads();
			const unsigned long long prefix = s_prefix;
			for (uint32_t i = tid; i < n; i += blockDim.x) {
				const unsigned long long key = in[i];
				if (sh == 56 || (key >> (sh + 8)) == prefix)
					atomicAdd(&s_hist[(uint32_t)(key >> sh) & 255u], 1u);
			}
			__syncthreads();
			if (tid == 0) {
				uint32_t want = s_want, cum = 0;
				int b = 255;

				for (; b > 0; b--) {
					if (cum + s_hist[b] >= want)
						break;
					cum += s_hist[b];
				}
				s_want = want - cum;
				s_prefix = (prefix << 8) | (unsigned)b;
			}
			__syncthreads();
		}
		const unsigned long long kth = s_prefix;
		for (uint32_t i = tid; i < n; i += blockDim.x) {
			const unsigned long long key = in[i];
			if (key >= kth)
				s_keys[atomicAdd(&s_n, 1u)] = key;
		}
		__syncthreads();
		m = s_n;	// == k
	}

	uint32_t npow2 = 2;
	while (npow2 < m)
		npow2 <<= 1;
	for (uint32_t i = m + tid; i < npow2; i += blockDim.x)
		s_keys[i] = 0;
	bitonic_sort_desc(s_keys, npow2);

	const uint32_t cnt = m < k ? m : k;
	Rec *out = recs + (size_t)q * k;
	for (uint32_t r = tid; r < k; r += blockDim.x) {
		Rec rec;
		if (r < cnt) {
			const unsigned long long key = s_keys[r];
			rec.doc_id = doc_ids[(uint32_t)key];
			rec.score = __uint_as_float((uint32_t)(key >> 32));
			rec.valid = 1;
		} else {
			rec.doc_id = 0;
			rec.score = 0.f;
			rec.valid = 0;
		}
		out[r] = rec;
	}
	if (tid == 0)
		counts[q] = cnt;
}

/* Large-k path: records from one query's fully sorted candidate keys. */
__global__ void __launch_bounds__(256)
emit_sorted_kernel(const unsigned long long *__restrict__ keys, uint32_t n,
    uint32_t k, const unsigned long long *__restrict__ doc_ids,
    Rec *__restrict__ out, uint32_t *__restrict__ count)
{
	const uint32_t cnt = n < k ? n : k;

	for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < k;
	    r += gridDim.x * blockDim.x) {
		Rec rec;
		if (r < cnt) {
			const unsigned long long key = keys[r];
			rec.doc_id = doc_ids[(uint32_t)key];
			rec.score = __uint_as_float((uint32_t)(key >> 32));
			rec.valid = 1;
		} else {
			rec.doc_id = 0;
			rec.score = 0.f;
			rec.valid = 0;
		}
		out[r] = rec;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0)
		*count = cnt;
}

/*
 * Cross-shard merge.  in[g][q][r] are per-shard lists sorted by (score desc,
 * id desc); shard g holds a higher id range than shard g-1.  Each record
 * computes its global rank by counting, in every other shard's list, the
 * records that precede it -- a binary search per shard, no shared memory, any
 * k.  One thread per (query, shard, position).
 */
__global__ void __launch_bounds__(256)
merge_topk_kernel(const Rec *__restrict__ in, uint32_t n_shards,
    uint32_t n_queries, uint32_t k, Rec *__restrict__ out)
{
	const unsigned long long total = (unsigned long long)n_queries * n_shards * k;
	unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;

	if (idx >= total)
		return;
	const uint32_t r = idx % k;
	const uint32_t g = (idx / k) % n_shards;
	const uint32_t q = idx / ((unsigned long long)k * n_shards);
	const Rec me = in[((size_t)g * n_queries + q) * k + r];

	if (!me.valid) {
		/*
		 * Invalid slots fill the tail: the out buffer is pre-cleared by
		 * the host wrapper, nothing to write.
		 */
		return;
	}
	uint32_t rank = r;
	for (uint32_t o = 0; o < n_shards; o++) {
		if (o == g)
			continue;
		const Rec *lst = in + ((size_t)o * n_queries + q) * k;
		uint32_t lo = 0, hi = k;

		/* first index whose record does NOT precede `me` */
		while (lo < hi) {
			const uint32_t mid = (lo + hi) >> 1;
			const Rec x = lst[mid];
			const bool before = x.valid && (x.score > me.score ||
			    (x.score == me.score && o > g));

			if (before)
				lo = mid + 1;
			else
				hi = mid;
		}
		rank += lo;
	}
	if (rank < k)
		out[(size_t)q * k + rank] = me;
}

/*
 * Segmented image (incremental refresh).  The index image is a
 * base segment plus a few delta segments on the same GPU; each scores the
 * batch on its own and writes a list of k_in = limit + (removed documents a
 * segment may still hold) records per query into in[g][q][.].  Segments do
 * NOT hold ordered id ranges, so ties are ordered by the external id itself.
 *
 * drop_dead_kernel: one warp per (segment, query) list; removes, in place and
 * keeping the order, the records whose document was deleted after the
 * segment was built (dead[dead_off[g] .. dead_off[g+1]) ascending ids) and
 * pads the tail with invalid records.
 */
__global__ void __launch_bounds__(128)
drop_dead_kernel(Rec *__restrict__ recs, uint32_t n_segs, uint32_t n_queries,
    uint32_t k_in, const unsigned long long *__restrict__ dead,
    const uint32_t *__restrict__ dead_off)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t list = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);

	if (list >= n_segs * n_queries)
		return;
	const uint32_t g = list / n_queries;
	const uint32_t d0 = dead_off[g], d1 = dead_off[g + 1];

	if (d0 == d1)
		return;
	Rec *lst = recs + (size_t)list * k_in;
	uint32_t out = 0;

	for (uint32_t base = 0; base < k_in; base += 32) {
		const uint32_t r = base + lane;
		Rec me = { 0, 0.f, 0 };
		bool live = false;

		if (r < k_in) {
			me = lst[r];
			live = me.valid != 0;
		}
		if (live) {
			uint32_t lo = d0, hi = d1;

			while (lo < hi) {
				const uint32_t mid = (lo + hi) >> 1;
				if (dead[mid] < me.doc_id)
					lo = mid + 1;
				else
					hi = mid;
			}
			live = !(lo < d1 && dead[lo] == me.doc_id);
		}
		const uint32_t m = __ballot_sync(0xffffffffu, live);

		/* Writes land at or below the chunk just read. */
		__syncwarp();
		if (live)
			lst[out + __popc(m & ((1u << lane) - 1))] = me;
		out += __popc(m);
		__syncwarp();
	}
	for (uint32_t r = out + lane; r < k_in; r += 32)
		lst[r] = Rec{ 0, 0.f, 0 };
}

__device__ __forceinline__ bool
rec_precedes(const Rec &x, const Rec &me)
{
	return x.valid && (x.score > me.score ||
	    (x.score == me.score && x.doc_id > me.doc_id));
}

/*
 * merge_segments_kernel: one thread per (query, segment, position); a record's
 * global rank is the number of records, in every list, that precede it in
 * (score desc, id desc) order -- a binary search per list.  The k_out best go
 * to out[q][rank]; counts[q] = min(k_out, live records).  `out` is
 * pre-cleared by the caller.
 */
__global__ void __launch_bounds__(256)
merge_segments_kernel(const Rec *__restrict__ in, uint32_t n_segs,
    uint32_t n_queries, uint32_t k_in, uint32_t k_out, Rec *__restrict__ out,
    uint32_t *__restrict__ counts)
{
	const unsigned long long total = (unsigned long long)n_queries * n_segs * k_in;
	const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;

	if (idx >= total)
		return;
	const uint32_t r = idx % k_in;
	const uint32_t g = (idx / k_in) % n_segs;
	const uint32_t q = idx / ((unsigned long long)k_in * n_segs);

	if (r == 0 && g == 0) {
		/* Live records of the query = valid prefix lengths, summed. */
		unsigned long long n = 0;

		for (uint32_t o = 0; o < n_segs; o++) {
			const Rec *lst = in + ((size_t)o * n_queries + q) * k_in;
			uint32_t lo = 0, hi = k_in;

			while (lo < hi) {
				const uint32_t mid = (lo + hi) >> 1;
				if (lst[mid].valid)
					lo = mid + 1;
				else
					hi = mid;
			}
			n += lo;
		}
		counts[q] = n < k_out ? (uint32_t)n : k_out;
	}
	const Rec me = in[((size_t)g * n_queries + q) * k_in + r];

	if (!me.valid)
		return;
	uint32_t rank = r;
	for (uint32_t o = 0; o < n_segs && rank < k_out; o++) {
		if (o == g)
			continue;
		const Rec *lst = in + ((size_t)o * n_queries + q) * k_in;
		uint32_t lo = 0, hi = k_in;

		while (lo < hi) {
			const uint32_t mid = (lo + hi) >> 1;
			if (rec_precedes(lst[mid], me))
				lo = mid + 1;
			else
				hi = mid;
		}
		rank += lo;
	}
	if (rank < k_out)
		out[(size_t)q * k_out + rank] = me;
}

#endif
