/*
 * score_bmw_kernel: exact top-k for OR queries with block-max pruning.
 *
 * The reference scores every document a query matches and keeps the best
 * `limit` in a heap (ref src/query/search.c:235-272, src/core/results.c:
 * 128-220).  The answer only needs the documents that can still enter the
 * heap, so this kernel never touches most postings (the dynamic pruning of
 * block-max WAND, done block-at-a-time instead of with posting cursors):
 *
 *   The dense document space is cut into BLOCKS of 2^bshift documents, 8192
 *   blocks make a CHUNK.  For every "column term" -- a list long enough to
 *   average a posting per two blocks -- the image holds, per block, the
 *   offset of its first posting (boff) and the largest query-independent
 *   weight any of its postings has (bmax: BM25 tf-normalisation, or the
 *   TF-IDF tf weight; recomputed when the index statistics move).  A score is
 *   weight x idf, both roundings monotonic, so  sum_t bmax[t][b] * idf[t]
 *   bounds every score in block b from above, bit for bit (float addition is
 *   monotonic, absent terms add +0).  Shorter lists have no arrays: their
 *   exact per-block maxima are folded in from the postings themselves.
 *
 *   work item = (query, chunk), handed out chunk-major from the highest ids
 *   down, so that a query's threshold (the k-th best key of everything its
 *   earlier items found, kept in global memory) is known when an item starts.
 *   Per item:
 *     1. one pass, a thread per block: bound against the threshold; what
 *        survives is compacted into a list (and counted per bound bucket);
 *     2. the surviving blocks are scored -- only the most promising when
 *        there are many (the buckets pick them; the pass is then repeated
 *        with the better threshold): a warp scores one block, the block's
 *        slice of every token's list, in TOKEN-LIST ORDER (the reference's
 *        float summation order), into a per-warp accumulator of 2^bshift
 *        sums; what beats the threshold joins the item's candidate buffer,
 *        which is cut back to the k best whenever a round ends (that k-th
 *        key is the new threshold);
 *     3. the item's <= k keys go to the same per-(query, chunk) cells
 *        finalize_cells_kernel merges for the stream kernel, and the query's
 *        threshold becomes the exact k-th best of all its cells so far.
 *
 * LOGIC = true serves boolean queries of at most 8 tokens (ref get_expr_bitmap,
 * src/query/search.c:118-174): every token of a matching document scores, NOT
 * side included (search.c:235-272), so the bound is the same sum.  What is
 * added is set membership: a block (superblock) bounds to zero when the tokens
 * present in it cannot satisfy the program (the subset closure of the query's
 * truth table -- this is where "a AND b" skips everything without both), and
 * the block accumulator keeps a membership byte per document next to the sum,
 * tested against the truth table when candidates are picked.  No threshold
 * priming: a term's k-th weight says nothing about documents that also
 * satisfy an AND.
 *
 * Arithmetic is st_score() of stream.cuh, hence identical bits; ties still
 * fall to the higher document id because a bound is compared as the key
 * (bound, last document of the block).
 */
#ifndef NXSB_GPU_BMW_CUH
#define NXSB_GPU_BMW_CUH

#define BMW_THREADS	256
#ifndef BMW_MIN_CTAS
#define BMW_MIN_CTAS	4			/* CTAs per SM the register budget is set for */
#endif
#define BMW_WARPS	(BMW_THREADS / 32)
#define BMW_CH_BLOCKS	8192u			/* blocks per chunk */
#define BMW_CH_SB	(BMW_CH_BLOCKS / 32u)		/* superblocks (32 blocks) per chunk */
#ifndef BMW_DENSE_SB
#define BMW_DENSE_SB	32u			/* more live superblocks than this: dense block bounds */
#endif
#ifndef BMW_CAND
#define BMW_CAND	1024u			/* candidate keys per item */
#endif
#define BMW_SEL		512u			/* blocks selected per round */
#ifndef BMW_PARTIAL_MIN
#define BMW_PARTIAL_MIN	(BMW_SEL / 2)		/* more live blocks than this: only the most promising this round */
#endif
#define BMW_K_MAX	128u			/* limit served by this kernel */
#define BMW_HIST	64u
#ifndef BMW_ROUND_BLOCKS
#define BMW_ROUND_BLOCKS 256u			/* blocks of a partial round, at least (64: 1.60, 128: 1.49, 256: 1.43, 448: 1.44 ms per C2 batch) */
#endif
#ifndef BMW_ROUND_K
#define BMW_ROUND_K	2u			/* ... and this many per requested result */
#endif
#define BMW_BCOL_NONE	0xffffffffu
#define BMW_SHIFT_MIN	5
#define BMW_SHIFT_MAX	8
#define BMW_TOK_GROUP	4			/* tokens whose postings are in flight together */

static_assert((BMW_CH_BLOCKS << BMW_SHIFT_MIN) % TILE_DOCS == 0, "chunks start at tile boundaries");
static_assert(BMW_K_MAX + (1u << BMW_SHIFT_MAX) <= BMW_CAND, "a block always fits after a cut");

/*
 * -DBMW_PROF: cycles thread 0 of every CTA spends per phase, summed into
 * stats[4..]: bounds + select, short-list bounds, pick, score, cut,
 * emit + merge, item setup.
 */
#ifdef BMW_PROF
#define BPROF_DECL	long long bp_t = clock64(); unsigned long long bp_acc[8] = { 0 }
#define BPROF(i)	do { if (tid == 0) { const long long _t = clock64(); bp_acc[i] += _t - bp_t; bp_t = _t; } } while (0)
#define BPROF_FLUSH()	do { if (tid == 0 && p.stats) for (int _i = 0; _i < 8; _i++) \
	if (bp_acc[_i]) atomicAdd(p.stats + 4 + _i, bp_acc[_i]); } while (0)
#else
#define BPROF_DECL	do { } while (0)
#define BPROF(i)	do { } while (0)
#define BPROF_FLUSH()	do { } while (0)
#endif

struct BmwParams {
	const uint2 *		post;
	const DTok *		toks;
	const QDesc *		queries;
	const uint32_t *	qlist;
	uint32_t		n_q, nchunks, nblocks, row_stride, n_docs, ntiles, k;
	const uint32_t *	boff;		/* [n_bcol][nblocks + 1] */
	const float *		bmax;		/* [n_bcol][row_stride], the batch's algorithm */
	const float *		smax;		/* [n_bcol][sb_stride]: maxima per superblock of 32 blocks */
	uint32_t		sb_stride, n_mt;
	const uint32_t *	tt;		/* boolean queries: [n_q][8] truth table, then [n_q][8] its subset closure */
	unsigned long long *	thr;		/* [n_q] */
	uint32_t *		tile_count;	/* [n_q][nchunks] */
	unsigned long long *	cand;		/* [n_q][nchunks][k] */
	uint32_t *		work_counter;
	const float *		logtab;
	float			K0, K1;
	unsigned long long *	stats;		/* [4] items, blocks scored, postings scored, rounds */
};

struct BmwTok {
	unsigned long long	post_off;
	const uint32_t *	fine;		/* slice boundaries per 2^fine_shift documents */
	uint32_t		clo, chi;	/* the chunk's slice of the list */
	float			idf;
	uint32_t		col;
	uint32_t		fine_shift;
	float			best;		/* the list's largest score anywhere */
	float			prime;		/* its k-th largest score (0: fewer postings) */
	const uint8_t *		mtmax;		/* mini-tile maxima bytes, or NULL */
	float			step;		/* what one unit of those bytes is worth */
	float			bidf;		/* idf in bounds: 0 for a token no matching document can hold */
	const uint32_t *	mtbits;		/* blocks of a mini-tile that hold a posting, or NULL */
};
static_assert(sizeof(BmwTok) == 72, "token records are 72 bytes");

#define BMW_LOGIC_SHIFT	5u			/* block size boolean queries are served at */
#define BMW_LOGIC_TOKENS	8u			/* tokens of a boolean query served here (a membership byte) */

template <uint32_t BSHIFT, bool LOGIC = false>
struct BmwCfg {
	static constexpr uint32_t BS = 1u << BSHIFT;
	static constexpr uint32_t NP = BS / 32;
	static constexpr uint32_t MAXTOK = LOGIC ? BMW_LOGIC_TOKENS : NXSB_MAX_QUERY_TOKENS;
	static constexpr size_t SMEM = BMW_CH_BLOCKS * 4 + BMW_CAND * 8 + BMW_K_MAX * 8 +
	    BMW_WARPS * BS * 4 * (LOGIC ? 2 : 1) + BMW_SEL * 2 + LOGTAB_N * 4 +
	    MAXTOK * sizeof(BmwTok);
};

/* ---- image side: block offsets and block maxima of the column terms ---- */

/* row[j] = first posting of the list whose document lies in block >= j. */
__global__ void __launch_bounds__(256)
build_block_offsets_kernel(const uint2 *__restrict__ post,
    const unsigned long long *__restrict__ term_off,
    const uint32_t *__restrict__ col_terms, uint32_t nblocks, uint32_t bshift,
    uint32_t *__restrict__ boff)
{
	const uint32_t t = col_terms[blockIdx.x];
	const unsigned long long s = term_off[t];
	const uint32_t df = (uint32_t)(term_off[t + 1] - s);
	const uint2 *list = post + s;
	uint32_t *row = boff + (size_t)blockIdx.x * (nblocks + 1);

	if (df == 0) {
		for (uint32_t j = threadIdx.x; j <= nblocks; j += blockDim.x)
			row[j] = 0;
		return;
	}
	for (uint32_t i = threadIdx.x; i < df; i += blockDim.x) {
		const uint32_t b = list[i].x >> bshift;
		const uint32_t lo = i ? (list[i - 1].x >> bshift) + 1 : 0;

		for (uint32_t j = lo; j <= b; j++)
			row[j] = i;
		if (i == df - 1) {
			for (uint32_t j = b + 1; j <= nblocks; j++)
				row[j] = df;
		}
	}
}

/*
 * bmax_bm25[c][b], bmax_tfidf[c][b] = the largest weight (st_score with
 * idf = 1: the same roundings as the scorer, so weight * idf reproduces the
 * largest score exactly) of column c's postings in block b.  Block (x, c)
 * strides over the list; postings of one block are neighbours, so a thread
 * folds its run before touching memory.  Pre-zeroed by the caller; weights
 * are positive, so unsigned order is value order.
 */
__global__ void __launch_bounds__(256)
block_max_kernel(const uint2 *__restrict__ post,
    const unsigned long long *__restrict__ term_off,
    const uint32_t *__restrict__ col_terms, uint32_t nblocks, uint32_t bshift,
    const float *__restrict__ logtab, float K0, float K1,
    float *__restrict__ bmax_bm25, float *__restrict__ bmax_tfidf)
{
	__shared__ float s_logtab[LOGTAB_N];
	const uint32_t t = col_terms[blockIdx.y];
	const unsigned long long s = term_off[t], e = term_off[t + 1];
	/* Rows are padded to whole superblocks (nblocks here = the row stride). */
	uint32_t *o_bm = reinterpret_cast<uint32_t *>(bmax_bm25) + (size_t)blockIdx.y * nblocks;
	uint32_t *o_tf = reinterpret_cast<uint32_t *>(bmax_tfidf) + (size_t)blockIdx.y * nblocks;

	for (uint32_t i = threadIdx.x; i < LOGTAB_N; i += blockDim.x)
		s_logtab[i] = logtab[i];
	__syncthreads();

	StreamParams sp;
	sp.K0 = K0;
	sp.K1 = K1;
	sp.doc_len = nullptr;
	/* Each thread takes 4 consecutive postings. */
	for (unsigned long long i = s + 4ull * ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x);
	    i < e; i += 4ull * gridDim.x * blockDim.x) {
		uint2 v[4];
		float wb[4], wt[4];

#pragma unroll
		for (int r = 0; r < 4; r++)
			v[r] = i + r < e ? post[i + r] : make_uint2(0u, 0u);
		st_score<false, NXSB_ALGO_BM25, 4>(sp, s_logtab, v, 1.f, wb);
		st_score<false, NXSB_ALGO_TFIDF, 4>(sp, s_logtab, v, 1.f, wt);
		uint32_t b = v[0].x >> bshift;
		float mb = wb[0], mt = wt[0];
#pragma unroll
		for (int r = 1; r < 4; r++) {
			if (i + r >= e)
				break;
			const uint32_t br = v[r].x >> bshift;

			if (br != b) {
				atomicMax(o_bm + b, __float_as_uint(mb));
				atomicMax(o_tf + b, __float_as_uint(mt));
				b = br;
				mb = wb[r];
				mt = wt[r];
			} else {
				mb = fmaxf(mb, wb[r]);
				mt = fmaxf(mt, wt[r]);
			}
		}
		atomicMax(o_bm + b, __float_as_uint(mb));
		atomicMax(o_tf + b, __float_as_uint(mt));
	}
}

/* smax[c][s] = the largest of bmax[c][32 s .. 32 s + 31] (rows zero padded). */
__global__ void __launch_bounds__(256)
superblock_max_kernel(const float *__restrict__ bmax_bm25, const float *__restrict__ bmax_tfidf,
    uint32_t n_bcol, uint32_t row_stride, uint32_t sb_stride,
    float *__restrict__ smax_bm25, float *__restrict__ smax_tfidf)
{
	const size_t n = (size_t)n_bcol * sb_stride;

	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
	    i += (size_t)gridDim.x * blockDim.x) {
		const uint32_t c = (uint32_t)(i / sb_stride), sb = (uint32_t)(i % sb_stride);
		float mb = 0.f, mt = 0.f;

		for (uint32_t b = sb * 32u; b < min(row_stride, sb * 32u + 32u); b++) {
			mb = fmaxf(mb, bmax_bm25[(size_t)c * row_stride + b]);
			mt = fmaxf(mt, bmax_tfidf[(size_t)c * row_stride + b]);
		}
		smax_bm25[i] = mb;
		smax_tfidf[i] = mt;
	}
}

/*
 * wmax_*[t] = the largest weight of any posting of term t (a warp per term):
 * from the block maxima where the term has them, else from its postings
 * (those lists are short).
 */
__global__ void __launch_bounds__(256)
term_wmax_kernel(const uint2 *__restrict__ post,
    const unsigned long long *__restrict__ term_off, const uint32_t *__restrict__ bcol,
    uint32_t n_terms, const float *__restrict__ logtab, float K0, float K1,
    const float *__restrict__ bmax_bm25, const float *__restrict__ bmax_tfidf,
    uint32_t row_stride, float *__restrict__ wmax_bm25, float *__restrict__ wmax_tfidf)
{
	__shared__ float s_logtab[LOGTAB_N];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t nw = (gridDim.x * blockDim.x) >> 5;

	for (uint32_t i = threadIdx.x; i < LOGTAB_N; i += blockDim.x)
		s_logtab[i] = logtab[i];
	__syncthreads();

	StreamParams sp;
	sp.K0 = K0;
	sp.K1 = K1;
	sp.doc_len = nullptr;
	for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_terms; t += nw) {
		float mb = 0.f, mt = 0.f;

		if (bcol[t] != BMW_BCOL_NONE) {
			const size_t r0 = (size_t)bcol[t] * row_stride;

			for (uint32_t b = lane; b < row_stride; b += 32) {
				mb = fmaxf(mb, bmax_bm25[r0 + b]);
				mt = fmaxf(mt, bmax_tfidf[r0 + b]);
			}
		} else {
			const unsigned long long s = term_off[t], e = term_off[t + 1];

			for (unsigned long long i = s + lane; i < e; i += 32) {
				const uint2 v[1] = { post[i] };
				float wb[1], wt[1];

				st_score<false, NXSB_ALGO_BM25, 1>(sp, s_logtab, v, 1.f, wb);
				st_score<false, NXSB_ALGO_TFIDF, 1>(sp, s_logtab, v, 1.f, wt);
				mb = fmaxf(mb, wb[0]);
				mt = fmaxf(mt, wt[0]);
			}
		}
		for (int o = 16; o; o >>= 1) {
			mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
			mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
		}
		if (lane == 0) {
			wmax_bm25[t] = mb;
			wmax_tfidf[t] = mt;
		}
	}
}

/*
 * Long lists without block arrays (DF_LONG <= df < a posting per two blocks)
 * get one BYTE per mini-tile of 2^MT_SHIFT documents: the largest weight of
 * the list's postings there, as a fraction of the list's largest weight,
 * rounded UP to a multiple of 1/255 -- a bound, never below the true maximum.
 * The scorer's superblock pass reads it instead of walking the postings; a
 * word of bits beside it says which blocks of the mini-tile hold a posting at
 * all: the boolean scorer's dense block pass takes its present-token masks
 * from there.
 */
__device__ __forceinline__ float
mt_step(float wmax)
{
	return __fmul_ru(wmax, 0.003921569f);	/* 1/255 rounded up */
}

__device__ __forceinline__ float
mt_bound(uint32_t q, float step)
{
	return __fmul_ru((float)q, step);
}

/* A thread per (long list, mini-tile); rows of the column terms stay zero. */
__global__ void __launch_bounds__(256)
minitile_max_kernel(const uint2 *__restrict__ post,
    const unsigned long long *__restrict__ term_off, const uint32_t *__restrict__ long_terms,
    const uint32_t *__restrict__ bcol, const uint32_t *__restrict__ skip_mt,
    uint32_t n_long, uint32_t n_mt, uint32_t mt_stride,
    const float *__restrict__ logtab, float K0, float K1,
    const float *__restrict__ wmax_bm25, const float *__restrict__ wmax_tfidf,
    uint8_t *__restrict__ mt_bm25, uint8_t *__restrict__ mt_tfidf,
    uint32_t *__restrict__ mt_bits, uint32_t bshift)
{
	__shared__ float s_logtab[LOGTAB_N];
	const size_t n = (size_t)n_long * n_mt;

	for (uint32_t i = threadIdx.x; i < LOGTAB_N; i += blockDim.x)
		s_logtab[i] = logtab[i];
	__syncthreads();

	StreamParams sp;
	sp.K0 = K0;
	sp.K1 = K1;
	sp.doc_len = nullptr;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
	    i += (size_t)gridDim.x * blockDim.x) {
		const uint32_t r = (uint32_t)(i / n_mt), m = (uint32_t)(i % n_mt);
		const uint32_t t = long_terms[r];

		if (bcol[t] != BMW_BCOL_NONE)
			continue;
		const uint32_t *row = skip_mt + (size_t)r * (n_mt + 1);
		const uint32_t lo = row[m], hi = row[m + 1];

		if (lo == hi)
			continue;	/* pre-zeroed */
		const uint2 *list = post + term_off[t];
		float mb = 0.f, mt = 0.f;
		uint32_t bits = 0;	/* which blocks of the mini-tile hold a posting */

		for (uint32_t x = lo; x < hi; x++) {
			const uint2 v[1] = { list[x] };
			float wb[1], wt[1];

			bits |= 1u << ((v[0].x & ((1u << MT_SHIFT) - 1u)) >> bshift);

			st_score<false, NXSB_ALGO_BM25, 1>(sp, s_logtab, v, 1.f, wb);
			st_score<false, NXSB_ALGO_TFIDF, 1>(sp, s_logtab, v, 1.f, wt);
			mb = fmaxf(mb, wb[0]);
			mt = fmaxf(mt, wt[0]);
		}
		const float sb = mt_step(wmax_bm25[t]), st = mt_step(wmax_tfidf[t]);
		uint32_t qb = min(255u, (uint32_t)ceilf(__fdiv_ru(mb, sb)));
		uint32_t qt = min(255u, (uint32_t)ceilf(__fdiv_ru(mt, st)));

		/* Belt and braces: the byte's bound is at least the maximum. */
		while (qb < 255u && mt_bound(qb, sb) < mb)
			qb++;
		while (qt < 255u && mt_bound(qt, st) < mt)
			qt++;
		mt_bm25[(size_t)r * mt_stride + m] = (uint8_t)qb;
		mt_tfidf[(size_t)r * mt_stride + m] = (uint8_t)qt;
		if (mt_bits)
			mt_bits[(size_t)r * mt_stride + m] = bits;
	}
}

/* ---- threshold priming: the k-th largest weight of every term ----------- */

/*
 * A query's threshold need not start at zero.  A document's score is a sum of
 * non-negative terms, and float addition never rounds a sum below an operand,
 * so at least k documents score >= (k-th largest weight of term t) * idf(t)
 * for every term t of the query: the largest of those products is a lower
 * bound of the query's k-th best score before a single posting is read
 * ("threshold priming" of the WAND family).  The image keeps, per term and
 * algorithm, the k-th largest weight for the k of a short ladder; a query
 * with limit k uses the first step >= k (fewer documents can only score
 * higher).  Exact, duplicates counted: weights tie all the time.
 *
 * One warp per term streams the list and keeps the 128 largest weights seen
 * in a shared-memory buffer: a weight joins only if it beats the current
 * 128th, and the buffer is sorted and cut when it fills.  Lists longer than
 * BMW_KTH_PART postings are cut in parts whose 128 best are merged by a second
 * launch of the same code over the parts' lists.
 */
#define BMW_LADDER	8u
#define BMW_KTH_BUF	256u
#define BMW_KTH_KEEP	128u
#define BMW_KTH_PART	32768u

__host__ __device__ __forceinline__ uint32_t
bmw_ladder_k(uint32_t j)
{
	return (uint32_t)(0x806432140a040201ull >> (8u * j)) & 0xffu;	/* 1 2 4 10 20 50 100 128 */
}

/* The first step of the ladder that covers limit k (k <= BMW_K_MAX). */
static inline uint32_t
bmw_ladder_step(uint32_t k)
{
	uint32_t j = 0;

	while (j + 1 < BMW_LADDER && bmw_ladder_k(j) < k)
		j++;
	return j;
}

static_assert(BMW_KTH_KEEP >= BMW_K_MAX, "the ladder reaches the kernel's largest limit");

/* Descending bitonic sort of n words (a power of two >= 64) by one warp. */
__device__ __forceinline__ void
warp_sort_desc(uint32_t *s, uint32_t n, uint32_t lane)
{
	for (uint32_t size = 2; size <= n; size <<= 1) {
		for (uint32_t stride = size >> 1; stride; stride >>= 1) {
			__syncwarp();
			for (uint32_t p = lane; p < n / 2; p += 32) {
				const uint32_t i = 2 * p - (p & (stride - 1));
				const uint32_t a = s[i], b = s[i + stride];

				if ((a < b) == ((i & size) == 0)) {
					s[i] = b;
					s[i + stride] = a;
				}
			}
		}
	}
	__syncwarp();
}

struct KthBuf {
	uint32_t *	buf;	/* [BMW_KTH_BUF] float bits (weights are positive) */
	uint32_t	cnt, tau;
};

__device__ __forceinline__ void
kth_cut(KthBuf &k, uint32_t lane)
{
	for (uint32_t i = k.cnt + lane; i < BMW_KTH_BUF; i += 32)
		k.buf[i] = 0;
	warp_sort_desc(k.buf, BMW_KTH_BUF, lane);
	k.cnt = min(k.cnt, BMW_KTH_KEEP);
	k.tau = k.cnt >= BMW_KTH_KEEP ? k.buf[BMW_KTH_KEEP - 1] : 0u;
}

/* One value per lane (0 = none) joins the buffer if it can still matter. */
__device__ __forceinline__ void
kth_push(KthBuf &k, uint32_t w, uint32_t lane)
{
	bool pass = w > k.tau;
	uint32_t m = __ballot_sync(0xffffffffu, pass);

	if (!m)
		return;
	if (k.cnt + __popc(m) > BMW_KTH_BUF) {
		kth_cut(k, lane);
		pass = w > k.tau;
		m = __ballot_sync(0xffffffffu, pass);
		if (!m)
			return;
	}
	if (pass)
		k.buf[k.cnt + __popc(m & ((1u << lane) - 1u))] = w;
	k.cnt += __popc(m);
}

/* Sort what is left; the buffer then holds min(cnt, 128) weights, descending. */
__device__ __forceinline__ void
kth_finish(KthBuf &k, uint32_t lane)
{
	uint32_t n = 64;

	__syncwarp();
	while (n < k.cnt)
		n <<= 1;
	for (uint32_t i = k.cnt + lane; i < n; i += 32)
		k.buf[i] = 0;
	warp_sort_desc(k.buf, n, lane);
	k.cnt = min(k.cnt, BMW_KTH_KEEP);
}

/*
 * PARTS = false: a warp per term with at most BMW_KTH_PART postings, ladder
 * written.  PARTS = true: a warp per unit = {term, part} of a longer list,
 * the part's 128 best (zero padded) written to scratch[unit][algorithm][128].
 */
template <bool PARTS>
__global__ void __launch_bounds__(256)
term_kth_kernel(const uint2 *__restrict__ post,
    const unsigned long long *__restrict__ term_off, uint32_t n_terms,
    const uint2 *__restrict__ units, uint32_t n_units,
    const float *__restrict__ logtab, float K0, float K1,
    float *__restrict__ kth_bm25, float *__restrict__ kth_tfidf,
    uint32_t *__restrict__ scratch)
{
	__shared__ float s_logtab[LOGTAB_N];
	__shared__ uint32_t s_buf[8][2][BMW_KTH_BUF];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
	const uint32_t n_work = PARTS ? n_units : n_terms;

	for (uint32_t i = threadIdx.x; i < LOGTAB_N; i += blockDim.x)
		s_logtab[i] = logtab[i];
	__syncthreads();

	StreamParams sp;
	sp.K0 = K0;
	sp.K1 = K1;
	sp.doc_len = nullptr;
	for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_work; w += nw) {
		const uint32_t t = PARTS ? units[w].x : w;
		unsigned long long s = term_off[t], e = term_off[t + 1];

		if (PARTS) {
			s += (unsigned long long)units[w].y * BMW_KTH_PART;
			e = min(e, s + BMW_KTH_PART);
		} else if (e - s > BMW_KTH_PART || e == s) {
			continue;	/* a long list (its parts are merged), or none */
		}
		KthBuf kb = { s_buf[warp][0], 0u, 0u }, kt = { s_buf[warp][1], 0u, 0u };

		for (unsigned long long i0 = s; i0 < e; i0 += 128) {
			uint2 v[4];
			float wb[4], wt[4];

#pragma unroll
			for (int r = 0; r < 4; r++) {
				const unsigned long long i = i0 + 32u * r + lane;

				v[r] = i < e ? __ldg(post + i) : make_uint2(0u, 0u);
			}
			st_score<false, NXSB_ALGO_BM25, 4>(sp, s_logtab, v, 1.f, wb);
			st_score<false, NXSB_ALGO_TFIDF, 4>(sp, s_logtab, v, 1.f, wt);
#pragma unroll
			for (int r = 0; r < 4; r++) {
				const bool valid = v[r].y != 0u;

				kth_push(kb, valid ? __float_as_uint(wb[r]) : 0u, lane);
				kth_push(kt, valid ? __float_as_uint(wt[r]) : 0u, lane);
			}
		}
		kth_finish(kb, lane);
		kth_finish(kt, lane);
		if (PARTS) {
			uint32_t *o = scratch + (size_t)w * 2 * BMW_KTH_KEEP;

			for (uint32_t i = lane; i < BMW_KTH_KEEP; i += 32) {
				o[i] = i < kb.cnt ? kb.buf[i] : 0u;
				o[BMW_KTH_KEEP + i] = i < kt.cnt ? kt.buf[i] : 0u;
			}
		} else if (lane < BMW_LADDER) {
			const uint32_t L = bmw_ladder_k(lane);

			kth_bm25[(size_t)t * BMW_LADDER + lane] = kb.cnt >= L ? __uint_as_float(kb.buf[L - 1]) : 0.f;
			kth_tfidf[(size_t)t * BMW_LADDER + lane] = kt.cnt >= L ? __uint_as_float(kt.buf[L - 1]) : 0.f;
		}
		__syncwarp();
	}
}

/* A warp per long term: longs[i] = {term, first unit}, units of a term adjoin. */
__global__ void __launch_bounds__(256)
term_kth_merge_kernel(const uint2 *__restrict__ longs, uint32_t n_long,
    const uint2 *__restrict__ units, uint32_t n_units,
    const uint32_t *__restrict__ scratch,
    float *__restrict__ kth_bm25, float *__restrict__ kth_tfidf)
{
	__shared__ uint32_t s_buf[8][2][BMW_KTH_BUF];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t nw = (gridDim.x * blockDim.x) >> 5;

	for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_long; w += nw) {
		const uint32_t t = longs[w].x, u0 = longs[w].y;
		KthBuf kb = { s_buf[warp][0], 0u, 0u }, kt = { s_buf[warp][1], 0u, 0u };

		for (uint32_t u = u0; u < n_units && units[u].x == t; u++) {
			const uint32_t *in = scratch + (size_t)u * 2 * BMW_KTH_KEEP;

			for (uint32_t i = lane; i < BMW_KTH_KEEP; i += 32) {
				kth_push(kb, in[i], lane);
				kth_push(kt, in[BMW_KTH_KEEP + i], lane);
			}
		}
		kth_finish(kb, lane);
		kth_finish(kt, lane);
		if (lane < BMW_LADDER) {
			const uint32_t L = bmw_ladder_k(lane);

			kth_bm25[(size_t)t * BMW_LADDER + lane] = kb.cnt >= L ? __uint_as_float(kb.buf[L - 1]) : 0.f;
			kth_tfidf[(size_t)t * BMW_LADDER + lane] = kt.cnt >= L ? __uint_as_float(kt.buf[L - 1]) : 0.f;
		}
		__syncwarp();
	}
}

/* ---- the scorer --------------------------------------------------------- */

template <int ALGO, uint32_t BSHIFT, bool LOGIC>
__global__ void __launch_bounds__(BMW_THREADS, BMW_MIN_CTAS)
score_bmw_kernel(const BmwParams p)
{
	using Cfg = BmwCfg<BSHIFT, LOGIC>;
	constexpr uint32_t BS = Cfg::BS, NP = Cfg::NP;
	constexpr uint32_t FULL = 0xffffffffu;
	constexpr uint32_t SBB = 32u, SB_SHIFT = BSHIFT + 5u;
	static_assert(BMW_CH_SB == BMW_THREADS, "a thread per superblock of the chunk");

	extern __shared__ __align__(16) unsigned char smem_bmw[];
	float *ub = reinterpret_cast<float *>(smem_bmw);		/* [CH_BLOCKS] bounds; 0 = scored */
	unsigned long long *s_cand = reinterpret_cast<unsigned long long *>(ub + BMW_CH_BLOCKS);
	unsigned long long *s_top = s_cand + BMW_CAND;			/* [K_MAX] cut buffer */
	float *s_acc = reinterpret_cast<float *>(s_top + BMW_K_MAX);	/* [WARPS][BS] */
	uint32_t *s_accm = reinterpret_cast<uint32_t *>(s_acc + BMW_WARPS * BS);	/* LOGIC: [WARPS][BS] membership */
	BmwTok *s_tok = reinterpret_cast<BmwTok *>(s_accm + (LOGIC ? BMW_WARPS * BS : 0));
	float *s_logtab = reinterpret_cast<float *>(s_tok + Cfg::MAXTOK);
	uint16_t *s_sel = reinterpret_cast<uint16_t *>(s_logtab + LOGTAB_N);

	__shared__ uint32_t s_item, s_selw[2], s_next, s_ncand, s_overflow, s_cut;
	__shared__ uint32_t s_hist[BMW_HIST];
	__shared__ unsigned long long s_theta, s_kth;
	__shared__ uint32_t s_tt[16];		/* LOGIC: truth table, subset closure */
	/*
	 * Scratch of the bound passes, in buffers that are idle until the first
	 * block is scored (4 CTAs per SM must stay within the 196 KB carve-out:
	 * the next step up costs the L1 half of its capacity).
	 */
	float *s_sbs = reinterpret_cast<float *>(s_cand);		/* [CH_SB] short lists' part of a superblock's bound */
	uint32_t *s_tmp = reinterpret_cast<uint32_t *>(s_cand) + BMW_CH_SB;	/* [CH_SB] one list's superblock maxima */
	uint32_t *s_wtmp = reinterpret_cast<uint32_t *>(s_top);		/* [WARPS * 32] one list's block maxima, per warp */
	uint8_t *s_sbmask = reinterpret_cast<uint8_t *>(s_tmp + BMW_CH_SB);	/* LOGIC: [CH_SB] lists without block arrays present in a superblock */
	auto sat = [&](uint32_t m) -> bool { return (s_tt[m >> 5] >> (m & 31u)) & 1u; };		/* documents with exactly tokens m match */
	auto satisfiable = [&](uint32_t m) -> bool { return (s_tt[8 + (m >> 5)] >> (m & 31u)) & 1u; };	/* ... some subset of m */
	static_assert(BMW_CAND * 8 >= 2 * BMW_CH_SB * 4 && BMW_K_MAX * 8 >= BMW_WARPS * 32 * 4, "scratch fits");

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t n_items = p.n_q * p.nchunks;
	const uint32_t k = p.k;
	/* Blocks of a partial round: enough to hold a few times k documents. */
	const uint32_t round_blocks = min(BMW_SEL - 64u, max(BMW_ROUND_BLOCKS, BMW_ROUND_K * k));
	float *wacc = s_acc + warp * BS;
	uint32_t *wmem = s_accm + (LOGIC ? warp * BS : 0);
	unsigned long long st_blocks = 0, st_post = 0, st_rounds = 0, st_items = 0;

	for (uint32_t i = tid; i < LOGTAB_N; i += BMW_THREADS)
		s_logtab[i] = p.logtab[i];

	StreamParams sp;
	sp.K0 = p.K0;
	sp.K1 = p.K1;
	sp.doc_len = nullptr;

	/* Thread 0 takes the item after this one while this one runs. */
	uint32_t next_item = 0;
	if (tid == 0)
		next_item = atomicAdd(p.work_counter, 1u);

	BPROF_DECL;
	for (;;) {
		__syncthreads();
		BPROF(5);
		if (tid == 0) {
			s_item = next_item;
			s_ncand = 0;
			next_item = atomicAdd(p.work_counter, 1u);
		}
		__syncthreads();
		const uint32_t item = s_item;

		if (item >= n_items)
			break;
		const uint32_t chunk = p.nchunks - 1 - item / p.n_q;
		const uint32_t slot = item % p.n_q;
		const QDesc qd = p.queries[p.qlist[slot]];
		const uint32_t ntok = qd.n_tokens;
		const uint32_t cb0 = chunk * BMW_CH_BLOCKS;
		const uint32_t nb = min(BMW_CH_BLOCKS, p.nblocks - cb0);
		const uint32_t doc0 = cb0 << BSHIFT;
		const uint32_t tile0 = doc0 >> TILE_SHIFT;
		const uint32_t tile1 = min(p.ntiles,
		    (uint32_t)((((unsigned long long)(cb0 + nb) << BSHIFT) + TILE_DOCS - 1) >> TILE_SHIFT));

		if (tid < ntok) {
			const DTok t = p.toks[qd.tok_off + tid];
			BmwTok bt;

			bt.post_off = t.post_off;
			bt.clo = __ldg(t.skip + tile0);
			bt.chi = __ldg(t.skip + tile1);
			bt.idf = t.idf;
			bt.col = t.bcol;
			bt.fine = t.fine;
			bt.fine_shift = t.fine_shift;
			bt.best = __fmul_rn(t.wmax, t.idf);
			bt.prime = __fmul_rn(t.wk, t.idf);
			bt.mtmax = t.mtmax;
			bt.mtbits = t.mtbits;
			bt.step = mt_step(t.wmax);
			bt.bidf = t.idf;
			if (LOGIC) {
				/*
				 * A token only ever seen under NOT -- no satisfying
				 * membership holds it -- adds nothing to a matching
				 * document's score: it stays out of the bounds (its
				 * postings still decide membership).
				 */
				const uint32_t *tt = p.tt + (size_t)slot * 8;
				const uint32_t pat[5] = { 0xaaaaaaaau, 0xccccccccu, 0xf0f0f0f0u, 0xff00ff00u, 0xffff0000u };
				uint32_t any = 0;

				for (uint32_t w = 0; w < 8; w++)
					any |= __ldg(tt + w) & (tid < 5 ? pat[tid] : ((w >> (tid - 5)) & 1u ? 0xffffffffu : 0u));
				if (!any)
					bt.bidf = 0.f;
			}
			s_tok[tid] = bt;
		}
		__syncthreads();
		if (tid == 0)
			st_items++;

		bool any_list = false;		/* a list without block arrays */
		float best_sum = 0.f;		/* no score of the query exceeds it */
		float prime = 0.f;		/* at least k documents score this much */
		for (uint32_t j = 0; j < ntok; j++) {
			any_list |= s_tok[j].col == BMW_BCOL_NONE;
			best_sum = __fadd_ru(best_sum, s_tok[j].bidf != 0.f ? s_tok[j].best : 0.f);
			prime = fmaxf(prime, s_tok[j].prime);
		}
		/* As a key: every document scoring >= prime stays above it. */
		const unsigned long long prime_key = (!LOGIC && prime > 0.f) ? make_key(prime, 0u) - 1ull : 0ull;
		/*
		 * A block's bound is summed columns first, short lists after:
		 * another order than the token list's, which can round a few ulp
		 * lower once three or more terms are involved.  2^-16 covers 32.
		 */
		const float infl = (any_list && ntok >= 3) ? 1.0000152587890625f : 1.f;

		if (tid == 0)
			s_theta = max(prime_key, *(volatile unsigned long long *)(p.thr + slot));
		s_wtmp[tid] = 0u;
		if (LOGIC && tid < 16)
			s_tt[tid] = __ldg(p.tt + (tid < 8 ? (size_t)slot * 8 + tid : (size_t)(p.n_q + slot) * 8 + tid - 8));
		__syncthreads();

		BPROF(6);
		/*
		 * ---- superblocks first: a thread per superblock of 32 blocks ----
		 * bound = the columns' superblock maxima (smax rows) + the short
		 * lists' exact maxima, folded in from the chunk's postings at this
		 * coarse grain.  Only the superblocks that can still beat the
		 * threshold get block bounds at all.
		 */
		const uint32_t nsb = (nb + SBB - 1) / SBB;
		uint32_t sb_present = 0;	/* LOGIC: tokens with a posting in this thread's superblock */
		if (any_list) {
			s_sbs[tid] = 0.f;
			for (uint32_t j = 0; j < ntok; j++) {
				const BmwTok bt = s_tok[j];

				if (bt.col != BMW_BCOL_NONE || bt.mtmax || bt.clo == bt.chi)
					continue;
				const uint2 *list = p.post + bt.post_off;

				s_tmp[tid] = 0u;
				__syncthreads();
				for (uint32_t i = bt.clo + tid; i < bt.chi; i += BMW_THREADS) {
					const uint2 v[1] = { __ldg(list + i) };
					float sc[1];

					st_score<false, ALGO, 1>(sp, s_logtab, v, bt.idf, sc);
					atomicMax(&s_tmp[(v[0].x - doc0) >> SB_SHIFT], __float_as_uint(sc[0]));
				}
				__syncthreads();
				if (bt.bidf != 0.f)
					s_sbs[tid] = __fadd_rn(s_sbs[tid], __uint_as_float(s_tmp[tid]));
				if (LOGIC && s_tmp[tid])
					sb_present |= 1u << j;
			}
			BPROF(1);
		}
		uint32_t live_mask, n_live;
		{
			float u = 0.f;

			for (uint32_t j = 0; j < ntok; j++) {
				const BmwTok &bt = s_tok[j];

				if (bt.col == BMW_BCOL_NONE)
					continue;
				const float m = tid < nsb ? __ldg(p.smax + (size_t)bt.col * p.sb_stride +
				    chunk * BMW_CH_SB + tid) : 0.f;

				u = __fadd_rn(u, __fmul_rn(m, bt.bidf));
				if (LOGIC && m > 0.f)
					sb_present |= 1u << j;
			}
			uint32_t nc_present = LOGIC ? sb_present & 0xffu : 0u;	/* without the columns: so far only short lists */
			if (LOGIC) {
				uint32_t cols = 0;

				for (uint32_t j = 0; j < ntok; j++)
					if (s_tok[j].col != BMW_BCOL_NONE)
						cols |= 1u << j;
				nc_present &= ~cols;
			}
			if (any_list) {
				/* Long lists: the bytes of the superblock's mini-tiles. */
				for (uint32_t j = 0; j < ntok; j++) {
					const BmwTok &bt = s_tok[j];

					if (bt.col != BMW_BCOL_NONE || !bt.mtmax || tid >= nsb)
						continue;
					constexpr uint32_t MPS = 1u << (SB_SHIFT - MT_SHIFT);
					const uint8_t *q = bt.mtmax + (size_t)(chunk * BMW_CH_SB + tid) * MPS;
					uint32_t qm = 0;

#pragma unroll
					for (uint32_t m = 0; m < MPS; m++)
						qm = max(qm, (uint32_t)__ldg(q + m));
					u = __fadd_rn(u, __fmul_ru(mt_bound(qm, bt.step), bt.bidf));
					if (LOGIC && qm) {
						sb_present |= 1u << j;
						if (!bt.mtbits)
							nc_present |= 1u << j;
					}
				}
				u = __fadd_rn(u, s_sbs[tid]);
			}
			u = tid < nsb ? __fmul_ru(u, infl) : 0.f;
			if (LOGIC) {
				/* The tokens present here cannot satisfy the program: nothing to find. */
				if (!satisfiable(sb_present))
					u = 0.f;
				s_sbmask[tid] = (uint8_t)nc_present;
			}
			const bool a = u != 0.f &&
			    make_key(u, doc0 + ((tid + 1u) << SB_SHIFT) - 1u) > s_theta;
			/* Warp w owns superblocks [32 w, 32 w + 32): its live ones are a mask. */
			live_mask = __ballot_sync(FULL, a);
			n_live = (uint32_t)__syncthreads_count(a);
		}
		BPROF(0);

		/*
		 * The dense way, for chunks where most superblocks are live: every short
		 * list's exact block maxima, summed into ub[] from the chunk's postings.
		 */
		auto short_lists_dense = [&]() {
#pragma unroll 8
			for (uint32_t b = tid; b < BMW_CH_BLOCKS; b += BMW_THREADS)
				ub[b] = 0.f;
			__syncthreads();
			for (uint32_t j = 0; j < ntok; j++) {
				const BmwTok bt = s_tok[j];

				if (bt.col != BMW_BCOL_NONE || bt.bidf == 0.f)
					continue;
				const uint2 *list = p.post + bt.post_off;

				for (uint32_t i0 = bt.clo; i0 < bt.chi; i0 += BMW_THREADS) {
					const uint32_t i = i0 + tid;
					const bool valid = i < bt.chi;
					uint2 v[1] = { valid ? __ldg(list + i) : make_uint2(0xffffffffu, 0u) };
					uint32_t xp = __shfl_up_sync(FULL, v[0].x, 1);

					if (lane == 0)
						xp = (valid && i > bt.clo) ? __ldg(list + i - 1).x : 0xffffffffu;
					const uint32_t b = (v[0].x - doc0) >> BSHIFT;

					/* The head of a block's run folds the run. */
					if (!valid || (i != bt.clo && ((xp - doc0) >> BSHIFT) == b))
						continue;
					float sc[1], m;

					st_score<false, ALGO, 1>(sp, s_logtab, v, bt.idf, sc);
					m = sc[0];
					for (uint32_t i2 = i + 1; i2 < bt.chi; i2++) {
						v[0] = __ldg(list + i2);
						if (((v[0].x - doc0) >> BSHIFT) != b)
							break;
						st_score<false, ALGO, 1>(sp, s_logtab, v, bt.idf, sc);
						m = fmaxf(m, sc[0]);
					}
					atomicAdd(ub + b, m);
				}
			}
			__syncthreads();
		};

		/*
		 * A warp scores block gb: every token's slice of the block, token
		 * order, into wacc; sums beating the threshold key join s_cand.
		 * False: the candidate buffer is full, nothing was recorded.
		 */
		auto score_block = [&](uint32_t gb) -> bool {
			const uint32_t base = gb << BSHIFT;

#pragma unroll
			for (uint32_t r = 0; r < NP; r++) {
				wacc[lane + 32 * r] = 0.f;
				if (LOGIC)
					wmem[lane + 32 * r] = 0u;
			}
			/* Lane j finds token j's slice of the block. */
			uint32_t my_lo = 0, my_hi = 0;
			if (lane < ntok) {
				const BmwTok &bt = s_tok[lane];

				if (bt.col != BMW_BCOL_NONE) {
					const uint32_t *row = p.boff + (size_t)bt.col * (p.nblocks + 1) + gb;

					my_lo = __ldg(row);
					my_hi = __ldg(row + 1);
				} else {
					/*
					 * A short list: the slice of the block's mini-tile
					 * (or tile) is a handful of postings; the warp
					 * loads all of it and keeps the block's.  Longer
					 * than the warp's registers: narrow it first.
					 */
					const uint32_t idx = base >> bt.fine_shift;

					my_lo = __ldg(bt.fine + idx);
					my_hi = __ldg(bt.fine + idx + 1);
					if (my_hi - my_lo > 32 * NP) {
						const uint2 *list = p.post + bt.post_off;
						uint32_t l = my_lo, h = my_hi;

						while (l < h) {
							const uint32_t mid = (l + h) >> 1;

							if (__ldg(list + mid).x < base)
								l = mid + 1;
							else
								h = mid;
						}
						my_lo = l;
						h = min(my_hi, l + BS);
						while (l < h) {
							const uint32_t mid = (l + h) >> 1;

							if (__ldg(list + mid).x < base + BS)
								l = mid + 1;
							else
								h = mid;
						}
						my_hi = l;
					}
				}
			}
			__syncwarp();
			for (uint32_t j0 = 0; j0 < ntok; j0 += BMW_TOK_GROUP) {
				uint2 v[BMW_TOK_GROUP][NP];

				/* Every load of the group is in flight before the first sum. */
#pragma unroll
				for (uint32_t g = 0; g < BMW_TOK_GROUP; g++) {
					const uint32_t j = j0 + g;
					const uint32_t lo_j = __shfl_sync(FULL, my_lo, j & 31u);
					const uint32_t hi_j = __shfl_sync(FULL, my_hi, j & 31u);

#pragma unroll
					for (uint32_t r = 0; r < NP; r++) {
						const uint32_t i = lo_j + lane + 32 * r;

						v[g][r] = make_uint2(base, 0u);
						if (j < ntok && i < hi_j)
							v[g][r] = __ldg(p.post + s_tok[j].post_off + i);
						/* A short list's slice may reach past the block. */
						if (v[g][r].x - base >= BS)
							v[g][r] = make_uint2(base, 0u);
					}
				}
#pragma unroll
				for (uint32_t g = 0; g < BMW_TOK_GROUP; g++) {
					const uint32_t j = j0 + g;

					if (j >= ntok)
						break;
					float sc[NP];

					st_score<false, ALGO, NP>(sp, s_logtab, v[g], s_tok[j].idf, sc);
#pragma unroll
					for (uint32_t r = 0; r < NP; r++) {
						if (v[g][r].y != 0u) {
							float *a = wacc + (v[g][r].x - base);

							*a = __fadd_rn(*a, sc[r]);
							if (LOGIC)
								wmem[v[g][r].x - base] |= 1u << j;
							st_post++;
						}
					}
					__syncwarp();
				}
			}
			/* Candidates: sums that beat the current threshold key. */
			const unsigned long long th = *(volatile unsigned long long *)&s_theta;
			unsigned long long keys[NP];
			uint32_t mine = 0;

#pragma unroll
			for (uint32_t r = 0; r < NP; r++) {
				const float val = wacc[lane + 32 * r];

				keys[r] = make_key(val, base + lane + 32 * r);
				if (val == 0.f || keys[r] <= th)
					keys[r] = 0;
				/* get_expr_bitmap: the document's tokens must satisfy the program. */
				if (LOGIC && !sat(wmem[lane + 32 * r]))
					keys[r] = 0;
				mine += keys[r] != 0;
			}
			/* inclusive prefix over the lanes */
			uint32_t incl = mine;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t x = __shfl_up_sync(FULL, incl, o);

				if ((int)lane >= o)
					incl += x;
			}
			const uint32_t total = __shfl_sync(FULL, incl, 31);
			uint32_t at = 0;

			if (total) {
				if (lane == 0) {
					/*
					 * Compare-and-swap, not add-then-undo: a failed
					 * reservation must never be visible, or a later
					 * one lands past the keys that end up counted.
					 */
					uint32_t seen = *(volatile uint32_t *)&s_ncand;

					for (;;) {
						if (seen + total > BMW_CAND) {
							s_overflow = 1;
							at = 0xffffffffu;
							break;
						}
						at = atomicCAS(&s_ncand, seen, seen + total);
						if (at == seen)
							break;
						seen = at;
					}
				}
				at = __shfl_sync(FULL, at, 0);
				if (at == 0xffffffffu)
					return false;
				uint32_t w = at + incl - mine;

#pragma unroll
				for (uint32_t r = 0; r < NP; r++)
					if (keys[r])
						s_cand[w++] = keys[r];
			}
			if (lane == 0)
				st_blocks++;
			__syncwarp();
			return true;
		};

		/*
		 * Cut s_cand back to its k best (if it holds that many); the k-th
		 * key lands in s_kth (else 0).  Block-wide, ends on a barrier.
		 */
		auto cut = [&]() {
			const uint32_t nc = s_ncand;

			if (tid == 0)
				s_kth = 0;
			if (nc >= k) {
				for (uint32_t i = tid; i < nc; i += BMW_THREADS) {
					const unsigned long long key = s_cand[i];
					uint32_t rank = 0;

					for (uint32_t j = 0; j < nc; j++)
						rank += s_cand[j] > key;
					if (rank < k)
						s_top[rank] = key;
				}
				__syncthreads();
				for (uint32_t i = tid; i < k; i += BMW_THREADS)
					s_cand[i] = s_top[i];
				if (tid == 0) {
					s_ncand = k;
					s_kth = s_top[k - 1];
				}
			}
			__syncthreads();
		};

		/* Most of the chunk is live: block bounds the dense way (throughput, not latency). */
		const bool dense = n_live > BMW_DENSE_SB;
		if (dense && any_list) {
			short_lists_dense();
			BPROF(1);
		}

		/* ---- rounds: bound + select, score, cut ---- */
		bool first = true;
		for (;;) {
			if (n_live == 0)
				break;		/* nothing of this chunk can enter the top k */
			if (tid == 0) {
				const unsigned long long g = *(volatile unsigned long long *)(p.thr + slot);

				if (g > s_theta)
					s_theta = g;
				s_selw[0] = s_selw[1] = s_next = s_overflow = 0;
			}
			if (tid < BMW_HIST)
				s_hist[tid] = 0;
			__syncthreads();
			const unsigned long long theta = s_theta;
			/* Buckets between the threshold's score and the query's best. */
			const float lo = __uint_as_float((uint32_t)(theta >> 32));
			const float hi = __fmul_ru(best_sum, infl);
			const float scale = hi > lo ? (float)BMW_HIST / (hi - lo) : 0.f;
			auto bin_of = [&](float u) -> uint32_t {
				const float x = (u - lo) * scale;

				return x <= 0.f ? 0u : min(BMW_HIST - 1u, (uint32_t)x);
			};
			/* (bound, last document of the block) must beat the threshold key. */
			auto alive = [&](uint32_t b, float u) -> bool {
				return u != 0.f && make_key(u, ((cb0 + b + 1u) << BSHIFT) - 1u) > theta;
			};

			/* Live blocks are compacted into s_sel while there is room, and
			 * counted per bucket; cutbin > 0: only the buckets from there up. */
			auto select = [&](uint32_t b, float u, uint32_t cutbin, bool count, uint32_t *selw) {
				const bool a = alive(b, u);
				const uint32_t bin = a ? bin_of(u) : 0u;
				const bool take = a && bin >= cutbin;
				const uint32_t tm = __ballot_sync(FULL, take);

				if (!tm)
					return;
				uint32_t at = 0;

				if (lane == 0)
					at = atomicAdd(selw, __popc(tm));
				at = __shfl_sync(FULL, at, 0) + __popc(tm & ((1u << lane) - 1u));
				if (take && at < BMW_SEL)
					s_sel[at] = (uint16_t)b;
				if (count && take) {
					/* one atomic per distinct bucket of the warp */
					const uint32_t same = __match_any_sync(tm, bin);

					if (lane == (uint32_t)__ffs(same) - 1u)
						atomicAdd(&s_hist[bin], __popc(same));
				}
			};
			/*
			 * The bounds of a warp's live superblocks against the current
			 * threshold, four superblocks in flight; `nm` collects the
			 * superblocks that still have a live block.
			 */
			auto sweep = [&](uint32_t mask, uint32_t cutbin, bool count, uint32_t *selw, uint32_t &nm) {
				while (mask) {
					uint32_t sbi[4], bb[4];
					float u[4];

#pragma unroll
					for (int x = 0; x < 4; x++) {
						sbi[x] = mask ? (uint32_t)__ffs(mask) - 1u : 0xffffffffu;
						mask &= mask - 1u;
					}
#pragma unroll
					for (int x = 0; x < 4; x++) {
						bb[x] = (warp * 32u + (sbi[x] & 31u)) * SBB + lane;
						u[x] = sbi[x] != 0xffffffffu ? ub[bb[x]] : 0.f;
					}
#pragma unroll
					for (int x = 0; x < 4; x++) {
						if (count && __any_sync(FULL, alive(bb[x], u[x])))
							nm |= 1u << (sbi[x] & 31u);
						select(bb[x], u[x], cutbin, count, selw);
					}
				}
			};
			/*
			 * A warp per live superblock, a lane per block.  The first
			 * pass computes the block bounds: ub[b] = (columns in token
			 * order + the short lists' block maxima, from the superblock's
			 * slice of each) * slack.  Superblocks with a live block make
			 * the next pass's list.
			 */
			uint32_t next_mask = 0;

			if (first && dense) {
				/*
				 * The bounds themselves, a thread per block, eight
				 * blocks at a time so that every load of a token's row
				 * is in flight before the first sum: ub[b] = (columns
				 * in token order + what the short lists left) * slack.
				 */
				constexpr uint32_t U = 8;

				/* Block of (warp, i, lane) = 32 (32 warp + i) + lane: superblock 32 warp + i. */
				for (uint32_t i0 = 0; i0 < (BMW_CH_BLOCKS / BMW_THREADS); i0 += U) {
					float u[U];
					uint32_t pm[LOGIC ? U : 1];	/* tokens that may have a posting in the block */

#pragma unroll
					for (uint32_t x = 0; x < U; x++) {
						u[x] = 0.f;
						if (LOGIC)	/* lists without block arrays: as coarse as their superblock */
							pm[x] = s_sbmask[warp * 32u + i0 + x];
					}
					for (uint32_t j = 0; j < ntok; j++) {
						const BmwTok &bt = s_tok[j];

						if (bt.col == BMW_BCOL_NONE)
							continue;
						const float *row = p.bmax + (size_t)bt.col * p.row_stride + cb0;
						float m[U];

#pragma unroll
						for (uint32_t x = 0; x < U; x++) {
							const uint32_t b = warp * (BMW_CH_BLOCKS / BMW_WARPS) + (i0 + x) * 32u + lane;

							/* A superblock the first pass ruled out costs no loads. */
							m[x] = (b < nb && ((live_mask >> (i0 + x)) & 1u)) ? __ldg(row + b) : 0.f;
						}
#pragma unroll
						for (uint32_t x = 0; x < U; x++) {
							u[x] = __fadd_rn(u[x], __fmul_rn(m[x], bt.bidf));
							if (LOGIC && m[x] > 0.f)
								pm[x] |= 1u << j;
						}
					}
					if (LOGIC && any_list) {
						/*
						 * Long lists: which blocks hold a posting is on record
						 * (their bound still comes from the postings: a
						 * mini-tile's maximum on every present block was
						 * measured and loses, 1.55 -> 2.4 ms per C2 batch).
						 */
						for (uint32_t j = 0; j < ntok; j++) {
							const BmwTok &bt = s_tok[j];

							if (bt.col != BMW_BCOL_NONE || !bt.mtbits)
								continue;
#pragma unroll
							for (uint32_t x = 0; x < U; x++) {
								const uint32_t b = warp * (BMW_CH_BLOCKS / BMW_WARPS) + (i0 + x) * 32u + lane;
								const uint32_t doc = (cb0 + b) << BSHIFT;

								if (b < nb && ((live_mask >> (i0 + x)) & 1u) &&
								    ((__ldg(bt.mtbits + (doc >> MT_SHIFT)) >>
								    ((doc & ((1u << MT_SHIFT) - 1u)) >> BSHIFT)) & 1u))
									pm[x] |= 1u << j;
							}
						}
					}
#pragma unroll
					for (uint32_t x = 0; x < U; x++) {
						const uint32_t b = warp * (BMW_CH_BLOCKS / BMW_WARPS) + (i0 + x) * 32u + lane;

						if (any_list)
							u[x] = __fadd_rn(u[x], ub[b]);
						u[x] = (b < nb && ((live_mask >> (i0 + x)) & 1u)) ? __fmul_ru(u[x], infl) : 0.f;
						if (LOGIC && !satisfiable(pm[x]))
							u[x] = 0.f;
						ub[b] = u[x];
					}
#pragma unroll
					for (uint32_t x = 0; x < U; x++) {
						const uint32_t b = warp * (BMW_CH_BLOCKS / BMW_WARPS) + (i0 + x) * 32u + lane;

						/* The warp's 32 blocks are one superblock. */
						if (__any_sync(FULL, alive(b, u[x])))
							next_mask |= 1u << (i0 + x);
						select(b, u[x], 0u, true, &s_selw[0]);
					}
				}
			} else if (!first) {
				sweep(live_mask, 0u, true, &s_selw[0], next_mask);
			} else
			for (uint32_t lm = live_mask; lm; lm &= lm - 1u) {
				const uint32_t sbi = (uint32_t)__ffs(lm) - 1u, sb = warp * 32u + sbi;
				const uint32_t b = sb * SBB + lane;
				const bool vb = b < nb;
				float u = 0.f;
				uint32_t pm = 0;	/* LOGIC: tokens with a posting in the block */

				if (first) {
					for (uint32_t j = 0; j < ntok; j++) {
						const BmwTok &bt = s_tok[j];

						if (bt.col == BMW_BCOL_NONE)
							continue;
						const float m = vb ? __ldg(p.bmax + (size_t)bt.col * p.row_stride + cb0 + b) : 0.f;

						u = __fadd_rn(u, __fmul_rn(m, bt.bidf));
						if (LOGIC && m > 0.f)
							pm |= 1u << j;
					}
					if (any_list) {
						const uint32_t sbdoc0 = doc0 + (sb << SB_SHIFT);

						for (uint32_t j = 0; j < ntok; j++) {
							const BmwTok &bt = s_tok[j];

							if (bt.col != BMW_BCOL_NONE || bt.clo == bt.chi)
								continue;
							/* The superblock's mini-tiles, or the tile around it. */
							const uint32_t i_lo = sbdoc0 >> bt.fine_shift;
							const uint32_t i_hi = bt.fine_shift > SB_SHIFT ? i_lo + 1u
							    : min((sbdoc0 + (1u << SB_SHIFT)) >> bt.fine_shift,
							      bt.fine_shift == MT_SHIFT ? p.n_mt : p.ntiles);
							const uint32_t lo = max(bt.clo, __ldg(bt.fine + i_lo));
							const uint32_t hi = min(bt.chi, __ldg(bt.fine + i_hi));
							const uint2 *list = p.post + bt.post_off;

							for (uint32_t i0 = lo; i0 < hi; i0 += 32) {
								const uint32_t i = i0 + lane;

								if (i < hi) {
									const uint2 v[1] = { __ldg(list + i) };
									const uint32_t rel = v[0].x - sbdoc0;

									if (rel < (1u << SB_SHIFT)) {
										float sc[1];

										st_score<false, ALGO, 1>(sp, s_logtab, v, bt.idf, sc);
										atomicMax(&s_wtmp[warp * 32 + (rel >> BSHIFT)],
										    __float_as_uint(sc[0]));
									}
								}
							}
							__syncwarp();
							if (bt.bidf != 0.f)
								u = __fadd_rn(u, __uint_as_float(s_wtmp[warp * 32 + lane]));
							if (LOGIC && s_wtmp[warp * 32 + lane])
								pm |= 1u << j;
							__syncwarp();
							s_wtmp[warp * 32 + lane] = 0u;
							__syncwarp();
						}
					}
					u = vb ? __fmul_ru(u, infl) : 0.f;
					if (LOGIC && !satisfiable(pm))
						u = 0.f;
					ub[b] = u;
				}
				if (__any_sync(FULL, alive(b, u)))
					next_mask |= 1u << sbi;
				select(b, u, 0u, true, &s_selw[0]);
			}
			live_mask = next_mask;
			first = false;
			__syncthreads();
			uint32_t found = s_selw[0];

			if (found == 0)
				break;
			/* Too many: only the blocks with the highest bounds this round. */
			const bool partial = found > BMW_PARTIAL_MIN;

			if (partial) {
				if (tid == 0) {
					uint32_t cum = 0;
					int bin = BMW_HIST - 1;

					for (; bin > 0; bin--) {
						cum += s_hist[bin];
						if (cum >= round_blocks)
							break;
					}
					s_cut = (uint32_t)bin;
				}
				__syncthreads();
				const uint32_t cutbin = s_cut;

				if (cutbin > 0) {
					/* its own counter: s_selw[0] is still being read */
					uint32_t unused = 0;

					sweep(live_mask, cutbin, false, &s_selw[1], unused);
					__syncthreads();
					found = s_selw[1];
				}
			}
			const uint32_t n_round = min(found, BMW_SEL);

			if (tid == 0)
				st_rounds++;
			BPROF(partial ? 2 : 0);

			/* ---- score the selected blocks, one per warp at a time ---- */
			for (;;) {
				uint32_t si = 0;

#ifdef BMW_SERIAL
				if (warp != 0)
					break;
#endif
				if (lane == 0)
					si = *(volatile uint32_t *)&s_overflow ? n_round : atomicAdd(&s_next, 1u);
				si = __shfl_sync(FULL, si, 0);
				if (si >= n_round)
					break;
				const uint32_t b = s_sel[si];

				if (score_block(cb0 + b) && lane == 0)
					ub[b] = 0.f;		/* done */
			}
			__syncthreads();
			BPROF(3);

			/* ---- cut the candidates back to the k best ---- */
			cut();
			if (tid == 0) {
				const unsigned long long kth = s_kth;

				if (kth > s_theta) {
					s_theta = kth;
					atomicMax(p.thr + slot, kth);
				}
			}
			BPROF(4);
			if (!partial && !s_overflow)
				break;		/* every live block was scored */
			__syncthreads();
		}
		BPROF(0);

		/* ---- the item's cell ---- */
		__syncthreads();
		const uint32_t nc = s_ncand;		/* <= k */

		if (nc) {
			const unsigned long long cell0 = (unsigned long long)slot * p.nchunks;
			unsigned long long *out = p.cand + (cell0 + chunk) * k;

			for (uint32_t i = tid; i < nc; i += BMW_THREADS)
				out[i] = s_cand[i];
			__threadfence();
			__syncthreads();
			if (tid == 0)
				*(volatile uint32_t *)(p.tile_count + cell0 + chunk) = nc;
			/*
			 * The query's exact k-th best so far: this cell merged with
			 * the cells its other chunks have published (keys first, then
			 * the count, fenced).  Later items of the query start from it.
			 */
			const unsigned long long cur = s_theta;

			for (uint32_t idx = tid; idx < p.nchunks * k; idx += BMW_THREADS) {
				const uint32_t c2 = idx / k, r = idx % k;

				if (c2 == chunk)
					continue;
				const uint32_t cnt = *(volatile uint32_t *)(p.tile_count + cell0 + c2);

				if (r < cnt) {
					__threadfence();
					const unsigned long long key =
					    *(volatile unsigned long long *)(p.cand + (cell0 + c2) * k + r);

					if (key >= cur) {
						const uint32_t at = atomicAdd(&s_ncand, 1u);

						if (at < BMW_CAND)
							s_cand[at] = key;
					}
				}
			}
			__syncthreads();
			const uint32_t np = s_ncand;

			if (np > nc && np >= k && np <= BMW_CAND) {
				for (uint32_t i = tid; i < np; i += BMW_THREADS) {
					const unsigned long long key = s_cand[i];
					uint32_t rank = 0;

					for (uint32_t j = 0; j < np; j++)
						rank += s_cand[j] > key;
					if (rank == k - 1 && key > cur)
						atomicMax(p.thr + slot, key);
				}
			}
		}
	}
	BPROF_FLUSH();
	if (p.stats && lane == 0) {
		if (st_items)
			atomicAdd(p.stats + 0, st_items);
		if (st_blocks)
			atomicAdd(p.stats + 1, st_blocks);
		if (st_rounds)
			atomicAdd(p.stats + 3, st_rounds);
	}
	if (p.stats) {
		for (int o = 16; o; o >>= 1)
			st_post += __shfl_xor_sync(FULL, st_post, o);
		if (lane == 0 && st_post)
			atomicAdd(p.stats + 2, st_post);
	}
}

#endif
