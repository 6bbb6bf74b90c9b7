/*
 * score_stream_kernel: the round-1 v3 hot kernel -- a warp-specialised,
 * TMA-fed version of the tile scorer (score.cuh explains the tiling).
 *
 * The v2 kernel (tiles.cuh) was latency bound: every (query, tile) work item
 * walked a chain of dependent global loads (work counter -> query -> token ->
 * skip row -> postings) with all 256 threads waiting at barriers in between
 * (profiles/README.md: issue slots 45 % busy, DRAM idle).  Here the chain is
 * taken off the scoring threads:
 *
 *   plan_items_kernel   one thread per (tile, query): the batched
 *                       term-to-posting-list lookup.  Writes a fixed-stride
 *                       record {total, ntok, (first posting, count, idf)...}
 *                       so the scorer needs ONE load per item, at an address
 *                       that depends only on the item number.
 *   producer warp       takes items from the global work counter (two in
 *                       flight), reads their records and streams the posting
 *                       slices into a ring of shared-memory stages with 1-D
 *                       bulk TMA copies (cp.async.bulk -> UBLKCP), signalling
 *                       "full" mbarriers.  It runs ahead of the consumers
 *                       across item boundaries, so the ring stays full while
 *                       they scan / sort / emit.
 *   consumer warps      wait on a stage, score its postings (one 8-byte LDS
 *                       per posting, lanes on consecutive postings so that
 *                       dense lists hit consecutive accumulator banks) and
 *                       accumulate into the 64 KB shared accumulator in
 *                       token-list order; a named barrier separates tokens.
 *
 * Per-item epilogue:
 *   - an item that fits one stage (<= 4 postings per consumer thread; 40 %
 *     of C2 items, 1 % of its postings) never scans the tile: each thread
 *     remembers the documents it touched and collects/clears them with an
 *     atomic exchange;
 *   - otherwise ONE fused pass reads the accumulator, keeps what beats the
 *     query's threshold in a 1024-entry buffer and writes zeros back.  If the
 *     buffer overflows, the uncollected survivors stay in the accumulator,
 *     the buffer is cut to its k best, whose k-th key tightens the
 *     threshold, and the pass repeats over what is left.
 *
 * Arithmetic is identical to tiles.cuh (same tables, same operation order).
 */
#ifndef NXSB_GPU_STREAM_CUH
#define NXSB_GPU_STREAM_CUH

#ifndef ST_CWARPS
#define ST_CWARPS	8			/* consumer warps */
#endif
#define ST_NCONS	(32 * ST_CWARPS)	/* consumer threads */
#define ST_THREADS	(ST_NCONS + 32)		/* + the producer warp */
#ifndef ST_SLOTS
#define ST_SLOTS	8			/* postings per thread per stage */
#endif
#define ST_STAGE_POST	(ST_SLOTS * ST_NCONS)	/* postings per stage */
#ifndef ST_NSTAGES
#define ST_NSTAGES	2
#endif
#define ST_MAXSUB	NXSB_MAX_QUERY_TOKENS	/* slices per stage */
#ifndef ST_CAND
#define ST_CAND		1024u			/* candidate buffer */
#endif
#define ST_SLOTS_LOGIC	6			/* boolean queries: room for the  */
#define ST_CAND_LOGIC	512u			/* membership bytes (16 KB)       */
#define ST_LOGIC_TOKENS	8u			/* tokens a membership byte holds */
#define ST_K_MAX	128u			/* limit served by this kernel */
#define ST_RANK_MAX	256u			/* candidates ranked by counting */
#define ST_PUSH		1024u			/* documents noted while accumulating */
#define ST_PUSH_LOGIC	512u

/*
 * -DST_PROF: per-phase cycle counters of the consumer warps (lane 0) and the
 * producer, summed into StreamParams::prof.  Development only.
 */
#ifdef ST_PROF
#define PROF_DECL	long long prof_t = clock64(); long long prof_acc[12] = { 0 }
#define PROF(i)		do { const long long _t = clock64(); prof_acc[i] += _t - prof_t; prof_t = _t; } while (0)
#define PROF_FLUSH(base) do { if ((threadIdx.x & 31) == 0) for (int _i = 0; _i < 12; _i++) \
	if (prof_acc[_i]) atomicAdd(p.prof + (base) + _i, (unsigned long long)prof_acc[_i]); } while (0)
#else
#define PROF_DECL	do { } while (0)
#define PROF(i)		do { } while (0)
#define PROF_FLUSH(b)	do { } while (0)
#endif

#define ST_F_FIRST	1u	/* first stage of an item */
#define ST_F_LAST	2u	/* last stage of an item */
#define ST_F_CONT	4u	/* slice 0 continues the previous stage's token */
#define ST_F_END	8u	/* no more work */
#define ST_F_FULL	16u	/* one slice filling every slot of the stage */
#define ST_F_SPARSE	32u	/* whole item in this stage, <= ST_CAND postings */
#define ST_F_DENSE	64u	/* a run of a dense column: one word per document */
#define ST_F_STORE	128u	/* dense run that leads its item: store, do not add */
#define ST_F_NSUB_SHIFT	8	/* 6 bits */
#define ST_F_TOK_SHIFT	16	/* token slot of slice 0 (boolean queries) */
#define ST_SUB_TOK_SHIFT 12	/* StageSub::b0 = first slot | token slot << 12 */

/* One planned work item: header + one entry per non-empty token slice. */
struct PlanHdr {
	uint32_t	total;		/* postings in the item */
	uint32_t	ntok;		/* non-empty slices */
	/*
	 * Base columns (shared dense prefix): the query's leading dense terms
	 * are not streamed; a document's sum starts, at its first posting,
	 * from the fold of score columns base_cols[8x .. 8x+7], x < nbase.
	 */
	uint32_t	base_cols;
	uint32_t	nbase;
};
struct PlanTok {
	/*
	 * First posting (absolute index) -- or, with bit 63 set, the first word
	 * of the tile's run in the term's dense column -- | token slot << 56.
	 */
	unsigned long long g0;
	uint32_t	n;
	float		idf;
};
static_assert(sizeof(PlanHdr) == 16 && sizeof(PlanTok) == 16, "plan record");
static_assert(2 * ST_K_MAX <= ST_CAND && 2 * ST_K_MAX <= ST_CAND_LOGIC,
    "overflow rounds must make progress");

struct StageSub {		/* a slice of one token inside a stage */
	uint16_t	b0, b1;		/* buffer slots [b0 & 0xfff, b1) */
	float		idf;
};
struct StageMeta {
	/* One 16-byte load tells a consumer everything about a FULL stage. */
	uint32_t	flags;		/* ST_F_* | nsub << ST_F_NSUB_SHIFT */
	uint32_t	slot, tile_lo;
	float		idf0;		/* = sub[0].idf */
	StageSub	sub[ST_MAXSUB];
	/* First stage of an item: score half of the query's threshold key as
	 * the producer saw it (0: none yet). */
	uint32_t	ths_bits;
	uint32_t	base_cols, nbase;	/* first stage of an item: PlanHdr's */
	uint32_t	pad;
};

struct StreamParams {
	const uint2 *		post;
	const unsigned char *	plan;		/* [n_items] records */
	uint32_t		plan_stride;	/* bytes */
	uint32_t		n_q, ntiles, k;
	unsigned long long *	thr;		/* [n_q] pruning thresholds */
	uint32_t *		tile_count;	/* [n_q][ntiles] candidates emitted */
	unsigned long long *	cand;		/* [n_q][ntiles][k] */
	uint32_t *		work_counter;
	const float *		logtab;
	const uint32_t *	doc_len;	/* WIDE */
	float			K0, K1;
	const uint32_t *	tt;		/* [n_q][8] truth tables (boolean) */
	const uint32_t *	dense;		/* dense columns (common.cuh) */
	unsigned long long	col_words;	/* words per column */
	const uint32_t *	dense_max;	/* [256] largest score of a column (bits) */
	unsigned long long *	prof;		/* ST_PROF counters or NULL */
};

/*
 * Per-variant shapes.  Boolean queries keep one membership byte per document
 * (bit j = "token slot j has a posting here") next to the accumulator and pay
 * for it with a smaller ring and candidate buffer.
 */
template <bool LOGIC>
struct StCfg {
	static constexpr uint32_t SLOTS = LOGIC ? ST_SLOTS_LOGIC : ST_SLOTS;
	static constexpr uint32_t STAGE_POST = SLOTS * ST_NCONS;
	static constexpr uint32_t CAND = LOGIC ? ST_CAND_LOGIC : ST_CAND;
	static constexpr uint32_t PUSH = LOGIC ? ST_PUSH_LOGIC : ST_PUSH;
	static constexpr size_t BASE = TILE_DOCS * 4 + ST_NSTAGES * STAGE_POST * 8 +
	    CAND * 8 + ST_NSTAGES * sizeof(StageMeta) + LOGTAB_N * 4 +
	    2 * ST_NSTAGES * 8 + 64 + PUSH * 2;
	static_assert(PUSH <= CAND, "noted documents become candidates");
	static constexpr size_t SMEM = BASE + (LOGIC ? TILE_DOCS : 0);
	static_assert(BASE % 16 == 0, "membership bytes are cleared 16 at a time");
	static_assert(STAGE_POST < (1u << ST_SUB_TOK_SHIFT), "slot bits");
};

/* ---- the batched term lookup ------------------------------------------ */

__global__ void __launch_bounds__(256)
plan_items_kernel(const QDesc *__restrict__ queries,
    const uint32_t *__restrict__ qlist, const uint2 *__restrict__ qbase,
    const uint32_t *__restrict__ tu, const DTok *__restrict__ toks,
    uint32_t n_q, uint32_t ntiles, uint32_t stride,
    unsigned char *__restrict__ plan)
{
	const unsigned long long item =
	    (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;

	if (item >= (unsigned long long)n_q * ntiles)
		return;
	/* Tile-major, highest tile first (ties prefer higher ids). */
	const uint32_t tile = ntiles - 1 - (uint32_t)(item / n_q);
	const uint32_t slot = (uint32_t)(item % n_q);
	const QDesc qd = queries[qlist[slot]];
	const uint2 qb = qbase ? qbase[slot] : make_uint2(0u, 0u);
	const uint32_t nbase = qb.y & 0x0fu;	/* 0x80: trailing (the epilogue adds them) */
	unsigned char *rec = plan + item * stride;
	PlanTok *out = reinterpret_cast<PlanTok *>(rec + sizeof(PlanHdr));
	uint32_t total = 0, m = 0, present = 0;

	for (uint32_t j = 0; j < qd.n_tokens; j++) {
		const DTok &t = toks[qd.tok_off + j];
		const uint32_t lo = __ldg(t.skip + tile);
		const uint32_t hi = __ldg(t.skip + tile + 1);

		present |= hi > lo ? 1u << (j & 31u) : 0u;
		/* The query's dense terms come in through its base columns. */
		if (nbase && t.dense_off != DENSE_NONE)
			continue;
		if (hi > lo) {
			PlanTok pt;

			/* Top byte: the token's slot in the query (boolean programs). */
			pt.g0 = (t.dense_off != DENSE_NONE
			    ? (t.dense_off + (unsigned long long)tile * TILE_DOCS) | (1ull << 63)
			    : t.post_off + lo) | ((unsigned long long)j << 56);
			pt.n = hi - lo;
			pt.idf = t.idf;
			out[m++] = pt;
			total += hi - lo;
		}
	}
	/*
	 * A dense column that leads the item is STORED into the accumulator
	 * instead of added (no zero-fill, no read).  The first float addition
	 * of a document's sum is commutative, so a dense slice in second place
	 * may change places with the first without changing a single bit.
	 */
	if (m >= 2 && !(out[0].g0 >> 63) && (out[1].g0 >> 63)) {
		const PlanTok t0 = out[0];

		out[0] = out[1];
		out[1] = t0;
	}
	/* Boolean query: can the tokens present in this tile satisfy it at all? */
	if (tu && !((tu[slot * 8u + ((present & 255u) >> 5)] >> (present & 31u)) & 1u))
		total = m = 0;
	PlanHdr h;
	h.total = total;
	h.ntok = m;
	h.base_cols = qb.x;
	h.nbase = qb.y;		/* count | (boolean) token slots or (OR) prefix number */
	*reinterpret_cast<PlanHdr *>(rec) = h;
}

/*
 * Boolean queries with at most 8 tokens: the postfix program (ref
 * get_expr_bitmap, src/query/search.c:118-174) is a function of which tokens
 * contain the document, i.e. of one membership byte.  Thread m evaluates it
 * for membership m; a warp vote packs 32 answers into a table word.
 */
__global__ void __launch_bounds__(256)
truth_tables_kernel(const QDesc *__restrict__ queries,
    const uint32_t *__restrict__ qlist, const int32_t *__restrict__ prog,
    uint32_t *__restrict__ tt, uint32_t *__restrict__ tu)
{
	__shared__ unsigned char s_any[256];
	const QDesc qd = queries[qlist[blockIdx.x]];
	const uint32_t m = threadIdx.x;
	bool st[NXSB_MAX_QUERY_PROG / 2 + 2];
	int sp = 0;

	for (uint32_t c = 0; c < qd.n_prog; c++) {
		const int32_t op = prog[qd.prog_off + c];

		if (op >= 0) {
			st[sp++] = (m >> op) & 1u;
		} else if (op == NXSB_OP_EMPTY) {
			st[sp++] = false;
		} else {
			const bool b = st[--sp];
			const bool a = st[sp - 1];

			st[sp - 1] = op == NXSB_OP_AND ? (a && b) :
			    op == NXSB_OP_OR ? (a || b) : (a && !b);
		}
	}
	const bool sat = sp ? st[sp - 1] : false;
	const uint32_t w = __ballot_sync(0xffffffffu, sat);

	if ((m & 31) == 0)
		tt[blockIdx.x * 8 + (m >> 5)] = w;
	/*
	 * tu[p] = "some document whose tokens are a subset of p satisfies the
	 * query" (sum over subsets, one token at a time): a tile in which only
	 * the tokens p have postings cannot hold a match when tu[p] is 0, and
	 * plan_items_kernel leaves it out.
	 */
	s_any[m] = sat;
	for (uint32_t bit = 1; bit < 256; bit <<= 1) {
		__syncthreads();
		const unsigned char below = (m & bit) ? s_any[m ^ bit] : 0;
		__syncthreads();
		s_any[m] |= below;
	}
	const uint32_t u = __ballot_sync(0xffffffffu, s_any[m] != 0);

	if ((m & 31) == 0)
		tu[blockIdx.x * 8 + (m >> 5)] = u;
}

/* ---- PTX wrappers ------------------------------------------------------ */

__device__ __forceinline__ uint32_t
smem_addr(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void
mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}

__device__ __forceinline__ void
mbar_arrive(uint32_t bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void
mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
	    :: "r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void
mbar_wait(uint32_t bar, uint32_t parity)
{
	uint32_t ok;

	do {
		asm volatile("{\n\t.reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	} while (!ok);
}

/* 1-D bulk TMA copy global -> shared, completion on an mbarrier. */
__device__ __forceinline__ void
tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
	    " [%0], [%1], %2, [%3];"
	    :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void
cons_barrier()
{
	asm volatile("bar.sync 1, %0;" :: "n"(ST_NCONS) : "memory");
}

__device__ __forceinline__ float
rcp_fast(float x)
{
	float r;

	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

/* Bitonic sort (descending) by the consumer threads only. */
__device__ __forceinline__ void
st_sort_desc(unsigned long long *s, uint32_t npow2, uint32_t ctid)
{
	for (uint32_t size = 2; size <= npow2; size <<= 1) {
		for (uint32_t stride = size >> 1; stride; stride >>= 1) {
			cons_barrier();
			for (uint32_t i = ctid; i < npow2 / 2; i += ST_NCONS) {
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
	cons_barrier();
}

/* (float)log(tf + 1) for the counts the table does not hold (cold). */
static __device__ __noinline__ float
st_log_slow(uint32_t tf)
{
	return (float)log((double)tf + 1.0);
}

/* Plain shared-memory word access by 32-bit shared address. */
__device__ __forceinline__ float
lds_f32(uint32_t addr)
{
	float v;

	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
	return v;
}

__device__ __forceinline__ void
sts_f32(uint32_t addr, float v)
{
	asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory");
}

/*
 * Scores of the NP postings a thread holds; same arithmetic as
 * score_posting() in tiles.cuh, written branch-free so that the postings'
 * dependency chains interleave.  Invalid slots carry the word 0 (tf = 0 ->
 * weight 0).  The document length comes out of the packed word with a byte
 * permute and one subtraction (0x4B000000 | dl is the float 2^23 + dl).
 * Counts >= 256 (no table entry) are rare and patched in one cold branch.
 */
template <bool WIDE, int ALGO, int NP>
__device__ __forceinline__ void
st_score(const StreamParams &p, const float *s_logtab, const uint2 (&v)[NP],
    float idf, float (&sc)[NP])
{
	float x[NP];
	uint32_t any = 0;

#pragma unroll
	for (int r = 0; r < NP; r++) {
		x[r] = s_logtab[__byte_perm(v[r].y, 0u, 0x4440)];	/* count & 255 */
		any |= v[r].y;
	}
	if (!WIDE)
		any &= 0xffffu;
	if (any >= LOGTAB_N) {
#pragma unroll
		for (int r = 0; r < NP; r++) {
			const uint32_t tf = WIDE ? v[r].y : (v[r].y & 0xffffu);

			if (tf >= LOGTAB_N)
				x[r] = st_log_slow(tf);
		}
	}
#pragma unroll
	for (int r = 0; r < NP; r++) {
		if (ALGO == NXSB_ALGO_TFIDF) {
			sc[r] = __fmul_rn(x[r], idf);
		} else {
			const float dl = WIDE ? (float)(int)__ldg(p.doc_len + v[r].x)
			    : __fsub_rn(__uint_as_float(__byte_perm(v[r].y, 0x4b000000u,
			      0x7632)), 8388608.f);
			const float d = __fmaf_rn(p.K1, dl, p.K0);

			sc[r] = __fmul_rn(__fmul_rn(x[r], rcp_fast(__fadd_rn(x[r], d))), idf);
		}
	}
}

/*
 * Score columns of the batch: the weight of (term, document) does not depend
 * on the query, so for the dense columns a batch refers to it is computed ONCE
 * per batch here -- same st_score() arithmetic, hence the same bits -- instead
 * of once per (query, document) in the scorer, which then streams finished
 * floats.  Absent documents (word 0) score +0.0.  Block (x, c) covers a
 * grid-strided part of column c; columns no token of the batch uses are skipped.
 */
template <int ALGO>
__global__ void __launch_bounds__(256)
dense_scores_kernel(const uint32_t *__restrict__ dense,
    const uint32_t *__restrict__ dterms, const uint32_t *__restrict__ used,
    const float *__restrict__ idf, const float *__restrict__ logtab,
    float K0, float K1, unsigned long long col_words, float *__restrict__ out,
    uint32_t *__restrict__ col_max)
{
	__shared__ float s_logtab[LOGTAB_N];
	const uint32_t c = blockIdx.y;
	float mx = 0.f;

	if (!used[c])
		return;
	for (uint32_t i = threadIdx.x; i < LOGTAB_N; i += blockDim.x)
		s_logtab[i] = logtab[i];
	__syncthreads();

	StreamParams p;
	p.K0 = K0;
	p.K1 = K1;
	p.doc_len = nullptr;
	const float w = idf[dterms[c]];
	const uint4 *in4 = reinterpret_cast<const uint4 *>(dense + c * col_words);
	float4 *out4 = reinterpret_cast<float4 *>(out + c * col_words);

	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	    i < col_words / 4; i += (unsigned long long)gridDim.x * blockDim.x) {
		const uint4 x = in4[i];
		const uint2 v[4] = { make_uint2(0u, x.x), make_uint2(0u, x.y),
		    make_uint2(0u, x.z), make_uint2(0u, x.w) };
		float sc[4];

		st_score<false, ALGO, 4>(p, s_logtab, v, w, sc);
		out4[i] = make_float4(sc[0], sc[1], sc[2], sc[3]);
		mx = fmaxf(fmaxf(mx, fmaxf(sc[0], sc[1])), fmaxf(sc[2], sc[3]));
	}
	/* The column's largest score (scores are >= 0: bit order = value order);
	 * one atomic per block, they all land on the same word. */
	__shared__ float s_mx[8];
	for (int o = 16; o; o >>= 1)
		mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	if ((threadIdx.x & 31) == 0)
		s_mx[threadIdx.x >> 5] = mx;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < 8; w++)
			mx = fmaxf(mx, s_mx[w]);
		if (mx > 0.f)
			atomicMax(col_max + c, __float_as_uint(mx));
	}
}

/* val + the trailing score columns of doc, in token order (cold: epilogues only). */
static __device__ __noinline__ float
st_trail_of(const float *colf, unsigned long long col_words, uint32_t cols,
    uint32_t n, float val, uint32_t doc)
{
	for (uint32_t x = 0; x < n; x++)
		val = __fadd_rn(val, __ldg(colf + ((cols >> (8u * x)) & 0xffu) * col_words + doc));
	return val;
}

template <bool LOGIC, bool WIDE, int ALGO>
__global__ void __launch_bounds__(ST_THREADS, 2)
score_stream_kernel(const StreamParams p)
{
	using Cfg = StCfg<LOGIC>;
	constexpr uint32_t SLOTS = Cfg::SLOTS, STAGE_POST = Cfg::STAGE_POST, CAND = Cfg::CAND;
	constexpr uint32_t PUSH = Cfg::PUSH;

	extern __shared__ __align__(128) unsigned char smem_stream[];
	float *acc = reinterpret_cast<float *>(smem_stream);
	uint2 *ring = reinterpret_cast<uint2 *>(acc + TILE_DOCS);
	unsigned long long *s_cand = reinterpret_cast<unsigned long long *>(
	    ring + ST_NSTAGES * STAGE_POST);
	StageMeta *meta = reinterpret_cast<StageMeta *>(s_cand + CAND);
	float *s_logtab = reinterpret_cast<float *>(meta + ST_NSTAGES);
	unsigned long long *bars = reinterpret_cast<unsigned long long *>(
	    s_logtab + LOGTAB_N);		/* full[NSTAGES], empty[NSTAGES] */
	uint32_t *s_misc = reinterpret_cast<uint32_t *>(bars + 2 * ST_NSTAGES);
	uint32_t *s_ncand = s_misc;			/* [2], by item parity */
	unsigned long long *s_theta = reinterpret_cast<unsigned long long *>(s_misc + 2);
	uint32_t *s_tt = s_misc + 4;			/* [8] truth table (LOGIC) */
	uint32_t *s_npush = s_misc + 12;		/* [2], by item parity */
	/* tile-relative documents whose running sum reached the threshold */
	uint16_t *s_push = reinterpret_cast<uint16_t *>(s_misc + 16);
	/* bit j of memb[d]: token slot j has a posting for tile document d */
	uint8_t *memb = smem_stream + Cfg::BASE;
	static_assert(TILE_DOCS <= 65536, "s_push holds 16-bit document offsets");

	const uint32_t tid = threadIdx.x;
	const uint32_t full0 = smem_addr(bars), empty0 = smem_addr(bars + ST_NSTAGES);

	for (uint32_t i = tid; i < LOGTAB_N; i += ST_THREADS)
		s_logtab[i] = p.logtab[i];
	{
		float4 *a4 = reinterpret_cast<float4 *>(acc);
		for (uint32_t i = tid; i < TILE_DOCS / 4; i += ST_THREADS)
			a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
		if (LOGIC) {
			uint4 *m4 = reinterpret_cast<uint4 *>(memb);
			for (uint32_t i = tid; i < TILE_DOCS / 16; i += ST_THREADS)
				m4[i] = make_uint4(0u, 0u, 0u, 0u);
		}
	}
	if (tid == 0) {
		for (uint32_t s = 0; s < ST_NSTAGES; s++) {
			mbar_init(full0 + 8 * s, 1);
			mbar_init(empty0 + 8 * s, ST_CWARPS);
		}
		s_ncand[0] = s_ncand[1] = 0;
		s_npush[0] = s_npush[1] = 0;
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();

	const uint32_t n_items = p.n_q * p.ntiles;

	if (tid >= ST_NCONS) {
		/* ================= producer warp ================= */
		const uint32_t lane = tid & 31;
		uint32_t ps = 0, pph = 1;	/* stage, parity of its empty barrier */
		uint32_t it0, it1;

		{
			uint32_t a = 0, b = 0;
			if (lane == 0) {
				a = atomicAdd(p.work_counter, 1u);
				b = atomicAdd(p.work_counter, 1u);
			}
			it0 = __shfl_sync(0xffffffffu, a, 0);
			it1 = __shfl_sync(0xffffffffu, b, 0);
		}
		auto load_hdr = [&](uint32_t it) -> uint4 {
			if (it >= n_items)
				return make_uint4(0u, 0u, 0u, 0u);
			return __ldg(reinterpret_cast<const uint4 *>(
			    p.plan + (unsigned long long)it * p.plan_stride));
		};
		/* Independent of the header load: every lane a record has room for. */
		const uint32_t rec_toks = p.plan_stride / 16u - 1u;
		auto load_tok = [&](uint32_t it, uint32_t) -> uint4 {
			if (it >= n_items || lane >= rec_toks)
				return make_uint4(0u, 0u, 0u, 0u);
			return __ldg(reinterpret_cast<const uint4 *>(
			    p.plan + (unsigned long long)it * p.plan_stride + 16) + lane);
		};
		/* The query's threshold as of now: a lower bound of what the
		 * consumers will find, which is all the inline check needs. */
		auto load_thr = [&](uint32_t it) -> uint32_t {
			if (it >= n_items)
				return 0u;
			return (uint32_t)(*(volatile const unsigned long long *)
			    (p.thr + it % p.n_q) >> 32);
		};
		uint4 h0 = load_hdr(it0);
		uint4 t0 = load_tok(it0, h0.y);
		uint32_t thr0 = load_thr(it0);

		/* Open stage state. */
		bool open = false;
		uint32_t used = 0, nsub = 0, bytes = 0, flags = 0, item_total = 0;

		auto commit = [&](uint32_t extra_flags, uint32_t slot, uint32_t tile_lo) {
			if (lane == 0) {
				StageMeta &m = meta[ps];
				uint32_t f = flags | extra_flags;

				if (!(f & ST_F_DENSE) && nsub == 1 && (m.sub[0].b0 & 0xfffu) == 0 &&
				    m.sub[0].b1 == STAGE_POST)
					f |= ST_F_FULL;
				f |= (uint32_t)(m.sub[0].b0 >> ST_SUB_TOK_SHIFT) << ST_F_TOK_SHIFT;
				if ((f & (ST_F_FIRST | ST_F_LAST)) == (ST_F_FIRST | ST_F_LAST) &&
				    item_total <= CAND)
					f |= ST_F_SPARSE;
				m.flags = f | (nsub << ST_F_NSUB_SHIFT);
				m.slot = slot;
				m.tile_lo = tile_lo;
				m.idf0 = m.sub[0].idf;
				m.ths_bits = thr0;
				m.base_cols = h0.z;
				m.nbase = h0.w;
				if (bytes)
					mbar_arrive_expect_tx(full0 + 8 * ps, bytes);
				else
					mbar_arrive(full0 + 8 * ps);
			}
			__syncwarp();
			open = false;
			if (++ps == ST_NSTAGES) {
				ps = 0;
				pph ^= 1;
			}
		};
		PROF_DECL;
		auto acquire = [&]() {
			PROF(0);		/* producer: everything else */
			mbar_wait(empty0 + 8 * ps, pph);
			PROF(1);		/* producer: wait for an empty stage */
			open = true;
			used = nsub = bytes = flags = 0;
		};

		while (it0 < n_items) {
			/* Keep the next item's number and record in flight. */
			uint32_t raw2 = 0;
			if (lane == 0)
				raw2 = atomicAdd(p.work_counter, 1u);
			const uint4 h1 = load_hdr(it1);
			const uint4 t1 = load_tok(it1, h1.y);
			const uint32_t thr1 = load_thr(it1);

			if (h0.x != 0) {
				const uint32_t tile = p.ntiles - 1 - it0 / p.n_q;
				const uint32_t slot = it0 % p.n_q;
				const uint32_t tile_lo = tile << TILE_SHIFT;
				const uint32_t ntok = h0.y;
				uint32_t first = ST_F_FIRST;

				item_total = h0.x;

				for (uint32_t j = 0; j < ntok; j++) {
					const uint32_t ghi = __shfl_sync(0xffffffffu, t0.y, j);
					const uint32_t tokj = ghi >> 24;
					unsigned long long g =
					    ((unsigned long long)(ghi & 0x00ffffffu) << 32) |
					    __shfl_sync(0xffffffffu, t0.x, j);
					uint32_t n = __shfl_sync(0xffffffffu, t0.z, j);
					const uint32_t idf_bits = __shfl_sync(0xffffffffu, t0.w, j);
					bool cont = false;

					if (tokj & 0x80u) {
						/*
						 * Dense column: the tile's 16384 words go out in
						 * runs of DENSE_RUN documents, one stage each.
						 */
						constexpr uint32_t DENSE_RUN = 2 * STAGE_POST;

						if (open)
							commit(0u, slot, tile_lo);
						for (uint32_t d0 = 0; d0 < TILE_DOCS; d0 += DENSE_RUN) {
							const uint32_t nd = min(DENSE_RUN, TILE_DOCS - d0);
							const bool last = d0 + nd == TILE_DOCS && j + 1 == ntok;

							acquire();
							flags = first | (d0 ? ST_F_CONT : 0u) | ST_F_DENSE |
							    (j == 0 ? ST_F_STORE : 0u);
							first = 0;
							if (lane == 0) {
								StageSub &sb = meta[ps].sub[0];

								sb.b0 = (uint16_t)((d0 / 4u) | ((tokj & 15u) << ST_SUB_TOK_SHIFT));
								sb.b1 = (uint16_t)nd;
								sb.idf = __uint_as_float(idf_bits);
								tma_load_1d(smem_addr(ring + ps * STAGE_POST),
								    p.dense + g + d0, nd * 4u, full0 + 8 * ps);
							}
							nsub = 1;
							bytes = nd * 4u;
							commit(last ? ST_F_LAST : 0u, slot, tile_lo);
						}
						continue;
					}

					while (n > 0) {
						if (!open) {
							acquire();
							flags = first | (cont ? ST_F_CONT : 0u);
							first = 0;
						}
						const uint32_t head = (uint32_t)g & 1u;
						const uint32_t avail = STAGE_POST - used;
						const uint32_t take = min(n, avail - head);
						const uint32_t cnt = (head + take + 1u) & ~1u;

						if (lane == 0) {
							StageSub &sb = meta[ps].sub[nsub];

							sb.b0 = (uint16_t)((used + head) | (tokj << ST_SUB_TOK_SHIFT));
							sb.b1 = (uint16_t)(used + head + take);
							sb.idf = __uint_as_float(idf_bits);
							tma_load_1d(smem_addr(ring + ps * STAGE_POST + used),
							    p.post + (g - head), cnt * 8u, full0 + 8 * ps);
						}
						nsub++;
						bytes += cnt * 8u;
						used += cnt;
						g += take;
						n -= take;
						cont = true;
						if (n > 0 || used + 2 > STAGE_POST || nsub == ST_MAXSUB) {
							const bool last = (n == 0 && j + 1 == ntok);

							commit(last ? ST_F_LAST : 0u, slot, tile_lo);
						}
					}
				}
				if (open)
					commit(ST_F_LAST, slot, tile_lo);
			}
			it0 = it1;
			h0 = h1;
			t0 = t1;
			thr0 = thr1;
			it1 = __shfl_sync(0xffffffffu, raw2, 0);
		}
		/* Tell the consumers there is nothing more. */
		acquire();
		flags = ST_F_END;
		commit(0u, 0u, 0u);
		PROF_FLUSH(12);
		return;
	}

	/* ================= consumer warps ================= */
	const uint32_t ctid = tid;
	uint32_t cs = 0, cph = 0, par = 0;
	unsigned long long theta_pref = 0;
	uint32_t tt_pref = 0;
	/*
	 * Inline selection.  Scores are positive, so a document's running sum
	 * only grows: one that ends above the query's threshold reaches it at
	 * some addition, and the thread doing that addition notes the document
	 * in s_push.  The epilogue then visits the noted documents instead of
	 * scanning the tile.  ths = +inf switches the noting off (no threshold
	 * yet, or a sparse item); `dirty` says the accumulator holds leftovers
	 * of such an item and must be zero-filled before the next addition.
	 */
	float ths = __uint_as_float(0x7f800000u);
	bool dirty = false;
	/*
	 * Base columns of the current item (PlanHdr): a document's sum starts,
	 * at its first posting (accumulator still 0 -- scores are positive),
	 * from the fold of the query's leading dense score columns, which is
	 * what streaming those columns first would have left there.
	 */
	constexpr bool BASE = !WIDE;
	uint32_t nbm = 0, base_cols = 0, base_slots = 0;	/* nbm: count | 0x80 if trailing */
#define NLEAD	((nbm & 0x80u) ? 0u : nbm)
#define NTRAIL	((nbm & 0x80u) ? (nbm & 0x0fu) : 0u)
	/*
	 * OR queries whose dense terms come LAST: the accumulator holds the sum
	 * of the other terms, and the epilogue adds the columns to whatever it
	 * looks at (trail_of), in token order.  Until then a sum can still grow
	 * by at most trail_ub (the columns' largest scores, rounded up), so the
	 * inline threshold is lowered by that much.
	 */
	/* The bound is recomputed where it is needed (cold paths): one register less. */
	auto trail_bound = [&]() -> float {
		float ub = 0.f;

		for (uint32_t x = 0; x < NTRAIL; x++)
			ub = __fadd_ru(ub, __uint_as_float(
			    __ldg(p.dense_max + ((base_cols >> (8u * x)) & 0xffu))));
		return ub;
	};
	auto trail_of = [&](float val, uint32_t doc) -> float {
		return st_trail_of(reinterpret_cast<const float *>(p.dense), p.col_words,
		    base_cols, NTRAIL, val, doc);
	};
	const float *colf = reinterpret_cast<const float *>(p.dense);
	/*
	 * Boolean queries also need the membership bits of the base terms: a
	 * document is in a dense term's list iff its column value is not 0.
	 * base_slots holds the terms' token slots, 3 bits each.
	 */
	auto base_of = [&](uint32_t doc, uint32_t &bits) -> float {
		float b = __ldg(colf + (base_cols & 0xffu) * p.col_words + doc);

		if (LOGIC)
			bits = b != 0.f ? 1u << (base_slots & 7u) : 0u;
		for (uint32_t x = 1; x < NLEAD; x++) {
			const float v = __ldg(colf + ((base_cols >> (8u * x)) & 0xffu) * p.col_words + doc);

			b = __fadd_rn(b, v);
			if (LOGIC)
				bits |= v != 0.f ? 1u << ((base_slots >> (3u * x)) & 7u) : 0u;
		}
		return b;
	};
	auto note = [&](uint32_t rel) {
		const uint32_t at = atomicAdd(s_npush + par, 1u);

		if (at < PUSH)
			s_push[at] = (uint16_t)rel;
	};

	PROF_DECL;
	for (;;) {
		PROF(0);			/* other */
		mbar_wait(full0 + 8 * cs, cph);
		PROF(1);			/* wait for a full stage */
		const StageMeta &m = meta[cs];
		const uint4 hdr = *reinterpret_cast<const uint4 *>(&m);
		const uint32_t flags = hdr.x;

		if (flags & ST_F_END) {
			PROF_FLUSH(0);
			break;
		}
		const uint32_t slot = hdr.y, tile_lo = hdr.z;
		const uint2 *buf = ring + cs * STAGE_POST;
		/* Shared address of acc[doc - tile_lo] = accb + 4 * doc. */
		const uint32_t accb = smem_addr(acc) - 4u * tile_lo;
		const uint32_t my_empty = empty0 + 8 * cs;

		if (flags & ST_F_FIRST) {
			const uint32_t tb = m.ths_bits;

			if (BASE) {
				nbm = m.nbase & (LOGIC ? 0x0fu : 0x8fu);
				base_slots = m.nbase >> 8;	/* boolean queries (OR: prefix number, unused here) */
				base_cols = m.base_cols;
			}

			if (ctid == 0)
				theta_pref = *(volatile unsigned long long *)(p.thr + slot);
			if (LOGIC && ctid < 8)
				tt_pref = __ldg(p.tt + slot * 8u + ctid);
			ths = (tb != 0u && !(flags & ST_F_SPARSE)) ? __uint_as_float(tb)
			    : __uint_as_float(0x7f800000u);
			if (BASE && NTRAIL && ths != __uint_as_float(0x7f800000u))
				ths = __fsub_rd(ths, trail_bound());
			if (dirty && !(flags & ST_F_STORE)) {
				/* Nobody reads the accumulator past the previous
				 * item's last epilogue barrier. */
				float4 *z4 = reinterpret_cast<float4 *>(acc);

#pragma unroll 4
				for (uint32_t i = ctid; i < TILE_DOCS / 4; i += ST_NCONS)
					z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
				if (LOGIC) {
					uint4 *m4 = reinterpret_cast<uint4 *>(memb);

					for (uint32_t i = ctid; i < TILE_DOCS / 16; i += ST_NCONS)
						m4[i] = make_uint4(0u, 0u, 0u, 0u);
				}
				cons_barrier();
				PROF(11);		/* lazy zero-fill */
			}
			dirty = false;
		}

		if (!WIDE && (flags & ST_F_DENSE)) {
			/*
			 * A run of a dense column: word i belongs to document
			 * tile_lo + 4 * doff4 + i, so the accumulator is updated
			 * with 16-byte vectors at consecutive addresses.  Absent
			 * documents hold 0 and add +0.0.
			 */
			constexpr int DG = SLOTS / 2;	/* 4-document groups per thread */
			const StageSub sb = m.sub[0];
			const uint32_t doff4 = sb.b0 & 0xfffu, nd4 = (uint32_t)sb.b1 / 4u;
			const float4 *wb = reinterpret_cast<const float4 *>(buf);
			float4 *a4 = reinterpret_cast<float4 *>(acc) + doff4;
			const bool store = (flags & ST_F_STORE) != 0;
			float4 w[DG], a[DG];

			/* The words are the finished scores (dense_scores_kernel). */
			if (!(flags & (ST_F_FIRST | ST_F_CONT)))
				cons_barrier();
#pragma unroll
			for (int g = 0; g < DG; g++) {
				const uint32_t idx = ctid + g * ST_NCONS;

				w[g] = make_float4(0.f, 0.f, 0.f, 0.f);
				if (idx < nd4)
					w[g] = wb[idx];
			}
#pragma unroll
			for (int g = 0; g < DG; g++) {
				const uint32_t idx = ctid + g * ST_NCONS;

				a[g] = make_float4(0.f, 0.f, 0.f, 0.f);
				if (!store && idx < nd4)
					a[g] = a4[idx];
			}
#pragma unroll
			for (int g = 0; g < DG; g++) {
				const uint32_t idx = ctid + g * ST_NCONS;

				if (idx < nd4) {
					/* store: 0 + w == w, bit for bit */
					a[g].x = __fadd_rn(a[g].x, w[g].x);
					a[g].y = __fadd_rn(a[g].y, w[g].y);
					a[g].z = __fadd_rn(a[g].z, w[g].z);
					a[g].w = __fadd_rn(a[g].w, w[g].w);
					a4[idx] = a[g];
					if (fmaxf(fmaxf(a[g].x, a[g].y), fmaxf(a[g].z, a[g].w)) >= ths) {
						const uint32_t rel = 4u * (doff4 + idx);

						if (a[g].x >= ths)
							note(rel);
						if (a[g].y >= ths)
							note(rel + 1u);
						if (a[g].z >= ths)
							note(rel + 2u);
						if (a[g].w >= ths)
							note(rel + 3u);
					}
					if (LOGIC) {
						const uint32_t bit = 1u << ((sb.b0 >> ST_SUB_TOK_SHIFT) & 7u);
						uint32_t *m32 = reinterpret_cast<uint32_t *>(memb) + doff4 + idx;
						const uint32_t bits = (w[g].x != 0.f ? bit : 0u) |
						    (w[g].y != 0.f ? bit << 8 : 0u) |
						    (w[g].z != 0.f ? bit << 16 : 0u) |
						    (w[g].w != 0.f ? bit << 24 : 0u);

						*m32 = store ? bits : (*m32 | bits);
					}
				}
			}
			PROF(10);		/* dense run */
		} else if (flags & ST_F_FULL) {
			/* The common case: every slot valid, one token. */
			uint2 v[SLOTS];
			float sc[SLOTS], a[SLOTS];

			if (!(flags & (ST_F_FIRST | ST_F_CONT)))
				cons_barrier();
			PROF(2);		/* token barrier */
#pragma unroll
			for (int r = 0; r < SLOTS; r++)
				v[r] = buf[ctid + r * ST_NCONS];
			float bs[SLOTS];
			uint32_t bb[SLOTS];
			if (BASE && NLEAD) {
				/* In flight while the postings are scored. */
#pragma unroll
				for (int r = 0; r < SLOTS; r++)
					bs[r] = base_of(v[r].x, bb[r]);
			}
			st_score<WIDE, ALGO, SLOTS>(p, s_logtab, v, __uint_as_float(hdr.w), sc);
			/* Documents of one list are distinct: batch the updates. */
#pragma unroll
			for (int r = 0; r < SLOTS; r++)
				a[r] = lds_f32(accb + 4u * v[r].x);
			if (BASE && NLEAD) {
#pragma unroll
				for (int r = 0; r < SLOTS; r++) {
					/* Later touches keep the bits the first one set. */
					if (LOGIC)
						bb[r] = a[r] == 0.f ? bb[r] : 0u;
					a[r] = a[r] == 0.f ? bs[r] : a[r];
				}
			}
			float top = 0.f;
#pragma unroll
			for (int r = 0; r < SLOTS; r++) {
				a[r] = __fadd_rn(a[r], sc[r]);
				sts_f32(accb + 4u * v[r].x, a[r]);
				top = fmaxf(top, a[r]);
			}
			if (top >= ths) {
#pragma unroll
				for (int r = 0; r < SLOTS; r++)
					if (a[r] >= ths)
						note(v[r].x - tile_lo);
			}
			if (LOGIC) {
				const uint8_t bit = (uint8_t)(1u << ((flags >> ST_F_TOK_SHIFT) & 7u));
				uint8_t mb[SLOTS];

#pragma unroll
				for (int r = 0; r < SLOTS; r++)
					mb[r] = memb[v[r].x - tile_lo];
#pragma unroll
				for (int r = 0; r < SLOTS; r++)
					memb[v[r].x - tile_lo] = mb[r] | bit |
					    (uint8_t)((BASE && NLEAD) ? bb[r] : 0u);
			}
			PROF(3);		/* full stage */
		} else {
			const uint32_t nsub = (flags >> ST_F_NSUB_SHIFT) & 0x3fu;

			for (uint32_t s = 0; s < nsub; s++) {
				const StageSub sb = m.sub[s];
				const uint32_t b0 = sb.b0 & 0xfffu, b1 = sb.b1, nb = b1 - b0;
				const uint8_t bit = (uint8_t)(1u << ((sb.b0 >> ST_SUB_TOK_SHIFT) & 7u));
				const float idf = sb.idf;

				/* A new token: the previous token's updates must have landed. */
				PROF(4);	/* partial stage */
				if (s != 0 || !(flags & (ST_F_FIRST | ST_F_CONT)))
					cons_barrier();
				PROF(2);
				/* Only the slot rows the slice touches, two per round. */
				for (uint32_t base = b0 & ~(ST_NCONS - 1u); base < b1;
				    base += 2 * ST_NCONS) {
					uint2 v[2];
					bool ok[2];
					float sc[2], a[2];

#pragma unroll
					for (int r = 0; r < 2; r++) {
						const uint32_t i = base + ctid + r * ST_NCONS;

						ok[r] = (i - b0) < nb;
						/* WIDE gathers doc_len[doc]: keep the dummy in range. */
						v[r] = make_uint2(tile_lo, 0u);
						if (ok[r])
							v[r] = buf[i];
					}
					float bs[2];
					uint32_t bb[2] = { 0u, 0u };
					if (BASE && NLEAD) {
#pragma unroll
						for (int r = 0; r < 2; r++)
							if (ok[r])
								bs[r] = base_of(v[r].x, bb[r]);
					}
					st_score<WIDE, ALGO, 2>(p, s_logtab, v, idf, sc);
#pragma unroll
					for (int r = 0; r < 2; r++)
						if (ok[r])
							a[r] = lds_f32(accb + 4u * v[r].x);
					if (BASE && NLEAD) {
#pragma unroll
						for (int r = 0; r < 2; r++)
							if (ok[r]) {
								if (LOGIC)
									bb[r] = a[r] == 0.f ? bb[r] : 0u;
								a[r] = a[r] == 0.f ? bs[r] : a[r];
							}
					}
#pragma unroll
					for (int r = 0; r < 2; r++)
						if (ok[r]) {
							a[r] = __fadd_rn(a[r], sc[r]);
							sts_f32(accb + 4u * v[r].x, a[r]);
							if (a[r] >= ths)
								note(v[r].x - tile_lo);
						}
					if (LOGIC) {
#pragma unroll
						for (int r = 0; r < 2; r++)
							if (ok[r])
								memb[v[r].x - tile_lo] |= bit | (uint8_t)bb[r];
					}
				}
			}
		}
		/*
		 * Stage consumed.  A sparse item's epilogue walks the staged
		 * postings once more, so it releases the stage afterwards.
		 */
		if (!(flags & ST_F_SPARSE)) {
			__syncwarp();
			if ((ctid & 31) == 0)
				mbar_arrive(my_empty);
		}
		if (++cs == ST_NSTAGES) {
			cs = 0;
			cph ^= 1;
		}
		PROF(4);
		if (!(flags & ST_F_LAST))
			continue;

		/* ---------------- item epilogue: top-k of the tile ---------------- */
		if (ctid == 0) {
			*s_theta = theta_pref;
			/* The next item's counter: last read before the previous
			 * item's final barrier, next bumped after this item's. */
			s_npush[par ^ 1u] = 0;
		}
		if (LOGIC && ctid < 8)
			s_tt[ctid] = tt_pref;
		cons_barrier();
		PROF(5);			/* barrier before the epilogue */
		/* Does a document with this membership byte satisfy the query? */
		auto in_set = [&](uint32_t mb) -> bool {
			return !LOGIC || ((s_tt[mb >> 5] >> (mb & 31u)) & 1u);
		};
		unsigned long long thr_key = *s_theta;
		const uint32_t k = p.k;
		uint32_t *ncand = s_ncand + par;	/* zero on entry */
		const uint32_t npush = s_npush[par];
		uint32_t total;

		par ^= 1u;
		if (ths != __uint_as_float(0x7f800000u) && npush <= PUSH) {
			/*
			 * Every document that can beat the threshold was noted
			 * (possibly more than once): its first visitor takes the
			 * final sum.  The rest of the accumulator is left as it is
			 * and zero-filled when the next item needs it clean.
			 */
			for (uint32_t i = ctid; i < npush; i += ST_NCONS) {
				const uint32_t rel = s_push[i];
				float val = atomicExch(acc + rel, 0.f);

				if (val != 0.f && in_set(LOGIC ? memb[rel] : 0u)) {
					if (BASE && NTRAIL)
						val = trail_of(val, tile_lo + rel);
					const unsigned long long key = make_key(val, tile_lo + rel);

					if (key > thr_key)
						s_cand[atomicAdd(ncand, 1u)] = key;	/* < PUSH <= CAND */
				}
			}
			dirty = true;
			PROF(7);
			cons_barrier();
			PROF(8);
			total = *(volatile uint32_t *)ncand;
		} else if (flags & ST_F_SPARSE) {
			/*
			 * Sparse item: visit only the documents its postings name;
			 * the first visitor of a document takes its sum and clears it.
			 */
			const float ths = thr_key ? __uint_as_float((uint32_t)(thr_key >> 32))
			    : __uint_as_float(1u);
			const uint32_t nsub = (flags >> ST_F_NSUB_SHIFT) & 0x3fu;

			for (uint32_t s = 0; s < nsub; s++) {
				const uint32_t b1 = m.sub[s].b1;

				for (uint32_t i = (m.sub[s].b0 & 0xfffu) + ctid; i < b1; i += ST_NCONS) {
					const uint32_t doc = buf[i].x;
					float val = atomicExch(acc + (doc - tile_lo), 0.f);
					bool member = true;

					if (LOGIC && val != 0.f) {
						/* Only a document's first visitor gets here. */
						member = in_set(memb[doc - tile_lo]);
						memb[doc - tile_lo] = 0;
					}
					if (BASE && NTRAIL && val != 0.f)
						val = trail_of(val, doc);
					if (val >= ths && member) {
						const unsigned long long key = make_key(val, doc);

						if (key > thr_key) {
							/* at < item total <= CAND */
							s_cand[atomicAdd(ncand, 1u)] = key;
						}
					}
				}
			}
			__syncwarp();
			if ((ctid & 31) == 0)
				mbar_arrive(my_empty);
			PROF(6);		/* sparse collect */
			cons_barrier();
			PROF(8);		/* barrier after collect / scan */
			total = *(volatile uint32_t *)ncand;
		} else {
			for (;;) {
				const float ths = thr_key ? __uint_as_float((uint32_t)(thr_key >> 32))
				    : __uint_as_float(1u);
				float4 *a4 = reinterpret_cast<float4 *>(acc);
				/* Trailing columns: what a raw sum must reach to be looked at. */
				const bool trl = BASE && NTRAIL;
				const float ths_raw = trl ? __fsub_rd(ths, trail_bound()) : ths;

				uint32_t *memb32 = reinterpret_cast<uint32_t *>(memb);

				/* false: the buffer is full, the document stays for the next round */
				auto visit = [&](float &val, uint32_t i, uint32_t mb) -> bool {
					if ((trl ? (val != 0.f && val >= ths_raw) : val >= ths) && in_set(mb)) {
						const unsigned long long key = make_key(
						    trl ? trail_of(val, tile_lo + i) : val, tile_lo + i);

						if (key > thr_key) {
							const uint32_t at = atomicAdd(ncand, 1u);

							if (at < CAND)
								s_cand[at] = key;
							else
								return false;
						}
					}
					val = 0.f;
					return true;
				};
				/*
				 * Batches of eight independent 16-byte loads; one
				 * compare decides for all 32 values of a batch (the
				 * common case: nothing beats the threshold).
				 */
				constexpr int NB = 8;
				constexpr uint32_t N4 = TILE_DOCS / 4;
				constexpr bool EXACT = N4 % (NB * ST_NCONS) == 0;
#pragma unroll 1
				for (uint32_t base = ctid; base < N4; base += NB * ST_NCONS) {
					float4 q[NB];
					float mx = 0.f;

#pragma unroll
					for (int j = 0; j < NB; j++) {
						q[j] = make_float4(0.f, 0.f, 0.f, 0.f);
						if (EXACT || base + j * ST_NCONS < N4)
							q[j] = a4[base + j * ST_NCONS];
					}
#pragma unroll
					for (int j = 0; j < NB; j++)
						mx = fmaxf(fmaxf(mx, fmaxf(q[j].x, q[j].y)),
						    fmaxf(q[j].z, q[j].w));
					if (mx >= ths_raw) {
#pragma unroll
						for (int j = 0; j < NB; j++) {
							const uint32_t i4 = base + j * ST_NCONS;

							if (EXACT || i4 < N4) {
								/* four membership bytes ride along */
								const uint32_t mb4 = LOGIC ? memb32[i4] : 0u;
								uint32_t keep = 0;

								if (!visit(q[j].x, 4 * i4 + 0, mb4 & 0xffu))
									keep |= 0x000000ffu;
								if (!visit(q[j].y, 4 * i4 + 1, (mb4 >> 8) & 0xffu))
									keep |= 0x0000ff00u;
								if (!visit(q[j].z, 4 * i4 + 2, (mb4 >> 16) & 0xffu))
									keep |= 0x00ff0000u;
								if (!visit(q[j].w, 4 * i4 + 3, mb4 >> 24))
									keep |= 0xff000000u;
								a4[i4] = q[j];
								if (LOGIC)
									memb32[i4] = mb4 & keep;
							}
						}
					} else {
#pragma unroll
						for (int j = 0; j < NB; j++)
							if (EXACT || base + j * ST_NCONS < N4) {
								a4[base + j * ST_NCONS] =
								    make_float4(0.f, 0.f, 0.f, 0.f);
								if (LOGIC)
									memb32[base + j * ST_NCONS] = 0u;
							}
					}
				}
				PROF(7);	/* scan + zero */
				cons_barrier();
				PROF(8);
				total = *(volatile uint32_t *)ncand;
				if (total <= CAND)
					break;
				/* Overflow: keep the k best, tighten, rescan what is left. */
				st_sort_desc(s_cand, CAND, ctid);
				thr_key = s_cand[k - 1];
				cons_barrier();
				if (ctid == 0)
					*ncand = k;
				cons_barrier();
			}
		}
		/* The other parity's counter was last read an item ago. */
		if (ctid == 0)
			s_ncand[par] = 0;

		if (total != 0) {
			const uint32_t tile = tile_lo >> TILE_SHIFT;
			const unsigned long long cell = (unsigned long long)slot * p.ntiles + tile;
			unsigned long long *out = p.cand + cell * k;

			if (total <= ST_RANK_MAX) {
				/*
				 * Keys are unique: a key's rank is the number of larger
				 * ones.  No barrier: s_cand is next written after the
				 * next item's first epilogue barrier.
				 */
				if (ctid < total) {
					const unsigned long long key = s_cand[ctid];
					uint32_t rank = 0;

					for (uint32_t j = 0; j < total; j++)
						rank += s_cand[j] > key;
					if (rank < k)
						out[rank] = key;
					if (rank == k - 1)
						atomicMax(p.thr + slot, key);
				}
			} else {
				uint32_t npow2 = 512;

				while (npow2 < total)
					npow2 <<= 1;
				for (uint32_t i = total + ctid; i < npow2; i += ST_NCONS)
					s_cand[i] = 0;
				st_sort_desc(s_cand, npow2, ctid);
				for (uint32_t i = ctid; i < k; i += ST_NCONS)
					out[i] = s_cand[i];
				if (ctid == 0)
					atomicMax(p.thr + slot, s_cand[k - 1]);
			}
			if (ctid == 0)
				p.tile_count[cell] = total < k ? total : k;
		}
		PROF(9);			/* rank + emit */
	}
}

#undef NLEAD
#undef NTRAIL

/*
 * Final per-query top-k from the per-tile candidate cells the stream kernel
 * wrote: cand[slot][tile][0 .. tile_count[slot][tile]).  One CTA per query;
 * same selection as finalize_topk_kernel (keys are unique).
 */
/*
 * Shared dense prefixes (engine.cu fill_batch): a query with index >= n_real
 * is a virtual one -- the leading dense terms of some queries on their own --
 * and leaves its top-k as KEYS in prefix_keys[q - n_real][k].  A real query
 * with base columns (qbase[slot].y != 0) scored only the documents its other
 * terms name; the rest of its top-k are the prefix's best documents that none
 * of those terms names (a binary search per list), so the two are merged.  The
 * prefix's k best suffice: a document of that list that IS named scores at
 * least its prefix sum in the query's own candidates, so the list always
 * accounts for k documents at or above anything the prefix ranks lower.
 */
struct FinalizeShared {
	const QDesc *		queries;
	const DTok *		toks;
	const uint2 *		post;
	const uint2 *		qbase;		/* per slot, or NULL */
	unsigned long long *	prefix_keys;
	uint32_t *		prefix_cnt;
	uint32_t		n_real;
};

__global__ void __launch_bounds__(256)
finalize_cells_kernel(const unsigned long long *__restrict__ cand,
    const uint32_t *__restrict__ tile_count, uint32_t ntiles,
    const uint32_t *__restrict__ qlist, uint32_t k,
    const unsigned long long *__restrict__ doc_ids,
    Rec *__restrict__ recs, uint32_t *__restrict__ counts,
    const FinalizeShared fs)
{
	__shared__ unsigned long long s_keys[SORT_CAP];
	__shared__ uint32_t s_hist[256];
	__shared__ uint32_t s_want, s_n, s_total;
	__shared__ unsigned long long s_prefix;

	const uint32_t slot = blockIdx.x, tid = threadIdx.x;
	const uint32_t q = qlist[slot];
	const unsigned long long *in = cand + (unsigned long long)slot * ntiles * k;
	const uint32_t *cnt = tile_count + (size_t)slot * ntiles;
	const uint32_t cells = ntiles * k;

	if (tid == 0) {
		s_total = 0;
		s_n = 0;
	}
	__syncthreads();
	{
		uint32_t mine = 0;

		for (uint32_t t = tid; t < ntiles; t += blockDim.x)
			mine += cnt[t];
		if (mine)
			atomicAdd(&s_total, mine);
	}
	__syncthreads();
	const uint32_t n = s_total;
	uint32_t m;		// keys to sort

	if (n <= SORT_CAP) {
		for (uint32_t i = tid; i < cells; i += blockDim.x)
			if (i % k < cnt[i / k])
				s_keys[atomicAdd(&s_n, 1u)] = in[i];
		__syncthreads();
		m = n;
	} else {
		/* n > SORT_CAP >= 2k: radix-select the k-th largest key. */
		if (tid == 0) {
			s_prefix = 0;
			s_want = k;
		}
		for (int sh = 56; sh >= 0; sh -= 8) {
			if (tid < 256)
				s_hist[tid] = 0;
			__syncthreads();
			const unsigned long long prefix = s_prefix;
			for (uint32_t i = tid; i < cells; i += blockDim.x) {
				if (i % k >= cnt[i / k])
					continue;
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
		for (uint32_t i = tid; i < cells; i += blockDim.x) {
			if (i % k >= cnt[i / k])
				continue;
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

	if (q >= fs.n_real) {
		/* A virtual query: its keys are what the real ones merge with. */
		const uint32_t vn = m < k ? m : k;

		for (uint32_t r = tid; r < vn; r += blockDim.x)
			fs.prefix_keys[(size_t)(q - fs.n_real) * k + r] = s_keys[r];
		if (tid == 0)
			fs.prefix_cnt[q - fs.n_real] = vn;
		return;
	}
	const uint2 qb = fs.qbase ? fs.qbase[slot] : make_uint2(0u, 0u);
	if (qb.y & 0xffu) {
		const uint32_t pn = qb.y >> 8;
		const uint32_t pc = fs.prefix_cnt[pn];
		const QDesc qd = fs.queries[q];
		const uint32_t own = m < k ? m : k;	/* own keys beyond k cannot win */

		__syncthreads();
		if (tid == 0)
			s_n = own;
		__syncthreads();
		for (uint32_t r = tid; r < pc; r += blockDim.x) {
			const unsigned long long key = fs.prefix_keys[(size_t)pn * k + r];
			const uint32_t doc = (uint32_t)key;
			bool named = false;

			for (uint32_t j = 0; j < qd.n_tokens && !named; j++) {
				const DTok t = fs.toks[qd.tok_off + j];

				if (t.dense_off != DENSE_NONE)
					continue;
				const uint2 *lst = fs.post + t.post_off;
				uint32_t lo = 0, hi = t.df_local;

				while (lo < hi) {
					const uint32_t mid = (lo + hi) >> 1;
					if (lst[mid].x < doc)
						lo = mid + 1;
					else
						hi = mid;
				}
				named = lo < t.df_local && lst[lo].x == doc;
			}
			if (!named)
				s_keys[atomicAdd(&s_n, 1u)] = key;	/* own + pc <= 2k <= SORT_CAP */
		}
		__syncthreads();
		m = s_n;
		npow2 = 2;
		while (npow2 < m)
			npow2 <<= 1;
		for (uint32_t i = m + tid; i < npow2; i += blockDim.x)
			s_keys[i] = 0;
		bitonic_sort_desc(s_keys, npow2);
	}

	const uint32_t cn = m < k ? m : k;
	Rec *out = recs + (size_t)q * k;
	for (uint32_t r = tid; r < k; r += blockDim.x) {
		Rec rec;
		if (r < cn) {
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
		counts[q] = cn;
}

#endif
