/*
 * nxsearch-b200 GPU engine: host-side orchestration behind the C ABI of
 * include/nxsb200_gpu.h.  One engine = one CUDA device = one document shard.
 *
 * There is deliberately no CPU path in this file: every entry point either
 * runs on the device or fails with an error message.
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "build.cuh"
#include "score.cuh"
#include "fuzzy.cuh"

#define CAND_ARENA_BYTES	(2ull << 30)	// candidate keys per sub-batch
#define MAX_HANDLES		64
#define PIPE_DEPTH		4	// searches in flight (search_begin/end)
#define N_LANES			2	// streams a search may run on (engine.cu "Lanes")
#define EV_PER_RUN		8
#define EV_RING			256

static thread_local char g_last_error[512];

/*
 * Device / pinned allocations and frees since the library was loaded
 * (nxsb_alloc_events): cudaFree is a device-wide synchronisation, so none of
 * them belongs on the steady-state search path.  Every allocation in this
 * file goes through the counted wrappers below.
 */
static std::atomic<uint64_t> g_alloc_events{0};

static cudaError_t
counted_malloc(void **p, size_t bytes)
{
	g_alloc_events++;
	return cudaMalloc(p, bytes);
}

static void
counted_free(void *p)
{
	if (p) {
		g_alloc_events++;
		cudaFree(p);
	}
}

extern "C" uint64_t
nxsb_alloc_events(void)
{
	return g_alloc_events.load();
}

/* Grow-only device / pinned-host buffers: no allocation in steady state. */
struct DevBuf {
	void *	p = nullptr;
	size_t	cap = 0;

	cudaError_t ensure(size_t bytes)
	{
		if (bytes <= cap)
			return cudaSuccess;
		release();
		const size_t want = bytes + bytes / 2 + 256;
		cudaError_t rc = counted_malloc(&p, want);
		if (rc == cudaSuccess)
			cap = want;
		else
			p = nullptr;
		return rc;
	}
	void release()
	{
		counted_free(p);
		p = nullptr;
		cap = 0;
	}
};

struct PinnedBuf {
	void *	p = nullptr;
	size_t	cap = 0;

	cudaError_t ensure(size_t bytes)
	{
		if (bytes <= cap)
			return cudaSuccess;
		release();
		const size_t want = bytes + bytes / 2 + 256;
		g_alloc_events++;
		cudaError_t rc = cudaMallocHost(&p, want);
		if (rc == cudaSuccess)
			cap = want;
		else
			p = nullptr;
		return rc;
	}
	void release()
	{
		if (p) {
			g_alloc_events++;
			cudaFreeHost(p);
		}
		p = nullptr;
		cap = 0;
	}
};

static inline size_t
align16(size_t n)
{
	return (n + 15) & ~(size_t)15;
}

struct Batch {
	bool		used = false;
	int		algo = 0;
	uint32_t	limit = 0, n_q = 0, n_tok = 0, n_prog = 0;
	uint32_t	max_tokens = 0;		// over the boolean queries (bitmap rows)
	uint32_t	max_tokens_all = 0;	// over every query (plan record stride)
	uint64_t	bytes = 0;		// algorithmic bytes
	bool		bmw = false;		// OR queries go through score_bmw_kernel
	uint32_t	n_logic_pruned = 0;	// ... and so do the first this many of q_logic
	std::vector<uint32_t> q_or, q_logic;	// host lists
	/*
	 * Shared dense prefixes (stream.cuh "base columns"): the distinct ordered
	 * sets of dense terms that lead the OR queries of the batch become
	 * n_virtual extra queries at the head of q_or (query index >= n_q,
	 * tokens appended after the batch's); q_base[i] = {columns, count |
	 * prefix number << 8} of q_or[i], count 0 = the query streams its own.
	 */
	uint32_t	n_virtual = 0, n_tok_all = 0;
	std::vector<uint2> q_base;
	uint2 *		d_qbase = nullptr;
	/* Boolean queries: {columns, count | token slots << 8}, no virtual queries. */
	std::vector<uint2> q_base_logic;
	uint2 *		d_qbase_logic = nullptr;
	unsigned long long *d_prefix_keys = nullptr;	// [n_virtual][limit] top-k keys
	uint32_t *	d_prefix_cnt = nullptr;		// [n_virtual]
	/* One H2D copy: [queries | tokens | prog | qlist_or | qlist_logic]. */
	DevBuf		desc;
	PinnedBuf	h_desc;
	/* [toks | tmp_skip | thr | cand_count | work]; the tail is re-zeroed per launch. */
	DevBuf		scratch;
	/* One D2H copy: [recs | counts]. */
	DevBuf		results;
	PinnedBuf	h_results;
	/* Views into the buffers above. */
	QDesc *		d_queries = nullptr;
	uint32_t *	d_tokens = nullptr;
	int32_t *	d_prog = nullptr;
	uint32_t *	d_qlist_or = nullptr, *d_qlist_logic = nullptr;
	DTok *		d_toks = nullptr;
	uint32_t *	d_tmp_skip = nullptr;
	unsigned long long *d_thr = nullptr;
	uint32_t *	d_cand_count = nullptr;
	uint32_t *	d_work = nullptr;
	size_t		zero_bytes = 0;		// thr .. work
	Rec *		d_recs = nullptr;
	uint32_t *	d_counts = nullptr;
	size_t		results_bytes = 0;
};

struct nxsb_engine {
	int		device = 0;
	cudaStream_t	own_stream = nullptr, stream = nullptr;
	char		err[512] = { 0 };
	uint64_t	launches = 0;
	int		n_sms = 148;

	/* shard image */
	uint32_t	n_docs = 0, n_terms = 0, ntiles = 0, n_long = 0;
	uint64_t	n_post = 0;
	bool		wide = false, loaded = false;
	uint2 *		d_post = nullptr;
	unsigned long long *d_term_off = nullptr;
	uint32_t *	d_df_local = nullptr;
	float *		d_idf_bm25 = nullptr, *d_idf_tfidf = nullptr;
	int32_t *	d_skip_row = nullptr;
	uint32_t *	d_dense = nullptr;		// [n_dense][ntiles * TILE_DOCS] words
	int32_t *	d_dense_col = nullptr;		// [V] column of a term or -1
	uint32_t	n_dense = 0;
	float *		d_dense_sc = nullptr;		// per-batch score columns, same shape
	uint32_t *	d_dense_terms = nullptr;	// [n_dense] term index of a column
	uint32_t *	d_dense_used = nullptr;		// [256] column referenced by the batch
	float		dense_min = 0.6f;		// df_local / N at which a list gets a column
	uint32_t *	d_skip = nullptr;
	uint32_t *	d_skip_mt = nullptr;		// [n_long][n_mt + 1] mini-tile rows
	uint32_t	n_mt = 0;
	unsigned long long *d_doc_ids = nullptr;
	uint32_t *	d_doc_len = nullptr;
	float *		d_logtab = nullptr;
	std::vector<uint32_t> h_df_local, h_df;
	std::vector<int32_t> h_dense_col;		// [V] as d_dense_col
	bool		share_dense = true;		// NXSB_SHARE_DENSE=0: development switch
	/*
	 * Block arrays of the column terms (bmw.cuh): lists averaging at least
	 * two postings per block of 2^bshift documents.
	 */
	bool		bmw_enabled = true;		// NXSB_BMW=0: every query streams
	uint32_t	logic_pruned_max_pos = 3;	// boolean queries with more positive tokens stream (NXSB_BMW_LOGIC_POS)
	uint32_t	bshift = 5, nblocks = 0, row_stride = 0, nchunks = 0, n_bcol = 0;
	uint32_t *	d_bcol = nullptr;		// [V] row of a term or BMW_BCOL_NONE
	uint32_t *	d_bcol_terms = nullptr;		// [n_bcol] term index of a row
	uint32_t *	d_boff = nullptr;		// [n_bcol][nblocks + 1]
	float *		d_bmax_bm25 = nullptr, *d_bmax_tfidf = nullptr;	// [n_bcol][row_stride]
	float *		d_smax_bm25 = nullptr, *d_smax_tfidf = nullptr;	// [n_bcol][sb_stride]
	uint32_t	sb_stride = 0;
	/* Long lists without block arrays: a byte per mini-tile (bmw.cuh). */
	uint8_t *	d_mt_bm25 = nullptr, *d_mt_tfidf = nullptr;	// [n_long][mt_stride]
	uint32_t *	d_mt_bits = nullptr;				// [n_long][mt_stride] blocks with a posting
	uint32_t *	d_long_terms = nullptr;				// [n_long] term index of a row
	uint32_t	mt_stride = 0;
	float *		d_wmax_bm25 = nullptr, *d_wmax_tfidf = nullptr;	// [V] largest weight of a term
	/* Threshold priming: [V][BMW_LADDER] k-th largest weight of a term. */
	float *		d_kth_bm25 = nullptr, *d_kth_tfidf = nullptr;
	uint2 *		d_kth_units = nullptr, *d_kth_longs = nullptr;	// parts of the long lists
	uint32_t *	d_kth_scratch = nullptr;
	uint32_t	n_kth_units = 0, n_kth_longs = 0;
	bool		prime_enabled = true;		// NXSB_PRIME=0: thresholds start at zero
	/* The weight tables (bmax, wmax, kth) were computed with these constants. */
	bool		tables_valid = false;
	float		tables_K0 = 0, tables_K1 = 0;
	unsigned long long *d_bmw_stats = nullptr;	// [4] counters of the BMW launches
	uint64_t	token_count = 0;
	uint32_t	doc_count = 0;
	float		K0 = 0, K1 = 0;
	bool		stats_valid = false;	// N > 0 and adl >= 1

	/* candidate arena (shared by all batches) */
	unsigned long long *d_cand = nullptr;
	size_t		cand_bytes = 0;
	unsigned long long *d_sort_tmp = nullptr;	// large-k path
	size_t		sort_tmp_bytes = 0;
	void *		d_cub_tmp = nullptr;
	size_t		cub_tmp_bytes = 0;
	unsigned char *	d_plan = nullptr;		// planned work items (stream kernel)
	size_t		plan_bytes = 0;
	unsigned long long *d_prof = nullptr;		// -DST_PROF phase counters
	uint32_t *	d_tile_cnt = nullptr;		// candidates per (query, tile)
	uint32_t *	d_tt = nullptr;			// truth tables of boolean queries
	size_t		tt_bytes = 0;
	size_t		tile_cnt_bytes = 0;
	bool		force_v2 = false;		// NXSB_KERNEL=v2: A/B against tiles.cuh

	/*
	 * Lanes: the scorer is latency bound (half the issue slots idle), and a
	 * batch ends on a tail of few busy CTAs plus a handful of small kernels
	 * and copies, so two batches in flight on two streams of the same GPU
	 * overlap well (measured: +22 % at 10M documents).  The image is shared
	 * and read-only; what a batch writes outside its own Batch -- the
	 * arenas above and the per-batch score columns -- exists once per lane.
	 * The members above ARE the current lane's; use_lane() swaps them.
	 * Searches alternate lanes by slot; everything that changes the image
	 * waits for both (sync_lanes).  An external stream or delta segments
	 * keep everything on lane 0.
	 */
	struct Lane {
		cudaStream_t	stream = nullptr;
		unsigned long long *d_cand = nullptr;
		size_t		cand_bytes = 0;
		unsigned long long *d_sort_tmp = nullptr;
		size_t		sort_tmp_bytes = 0;
		void *		d_cub_tmp = nullptr;
		size_t		cub_tmp_bytes = 0;
		unsigned char *	d_plan = nullptr;
		size_t		plan_bytes = 0;
		uint32_t *	d_tile_cnt = nullptr;
		size_t		tile_cnt_bytes = 0;
		uint32_t *	d_tt = nullptr;
		size_t		tt_bytes = 0;
		float *		d_dense_sc = nullptr;
		uint32_t *	d_dense_used = nullptr;
	};
	Lane		lanes[N_LANES];
	int		cur_lane = 0;
	bool		lanes_enabled = true;		// NXSB_LANES=1: one stream, as before
	bool		external_stream = false;

	/* cudaFuncSetAttribute / occupancy results, once per kernel variant. */
	struct KernFit { size_t smem; int per_sm; };
	std::map<const void *, KernFit> kfit;

	Batch		batches[MAX_HANDLES];
	Batch		oneshot;	// reused by nxsb_engine_search
	/* nxsb_engine_search_begin/end: pooled slots, one "done" event each */
	Batch		pipe[PIPE_DEPTH];
	cudaEvent_t	pipe_done[PIPE_DEPTH] = { nullptr };
	bool		pipe_busy[PIPE_DEPTH] = { false };

	/*
	 * A replicated engine (nxsb_engine_create_replicated): this object only
	 * dispatches; every replica is a complete engine on its own device and
	 * takes a contiguous share of each batch's queries.
	 */
	std::vector<nxsb_engine *> replicas;
	struct RepSplit {
		bool		busy = false;
		uint32_t	limit = 0;
		std::vector<uint32_t> q0;	// [replicas + 1] query ranges
		std::vector<int> h;		// child handles, -1 = no queries
	};
	RepSplit	rep_pipe[PIPE_DEPTH];

	/*
	 * A sharded engine (nxsb_engine_create_sharded): the same dispatcher,
	 * but every child holds a contiguous range of the documents (split in
	 * load_shard, scored with whole-index statistics), every child scores
	 * the whole batch, and the children's top-k lists meet on the first
	 * device -- peer copies over NVLink, one merge kernel -- before a
	 * single copy to the host.
	 */
	bool		sharded = false;
	cudaStream_t	shard_stream = nullptr;		// merge + D2H, first device
	std::vector<uint64_t> shard_last_id;		// [children] largest base document id
	struct SegAt { uint32_t child, local; };
	std::vector<SegAt> seg_at;			// delta segment g lives at seg_at[g - 1]
	struct ShardRun {
		bool		busy = false;
		uint32_t	limit = 0, n_q = 0;
		std::vector<int> h;			// child handles
		std::vector<DevBuf> part;		// [children] [query][limit] records, child's device
		std::vector<cudaEvent_t> part_done;	// ... and "copied to the first device"
		DevBuf		gather;			// [children][query][limit] records, first device
		DevBuf		out;			// [query][limit] records | [query] counts
		PinnedBuf	h_out;
		size_t		out_bytes = 0, counts_at = 0;
		cudaEvent_t	done = nullptr;
	};
	ShardRun	shard_pipe[PIPE_DEPTH];

	/* vocabulary (fuzzy) */
	FuzzyImage	fz;
	FuzzyScratch	fz_scratch;
	FuzzyScratch	fz_pipe[PIPE_DEPTH];	// lookups queued with a search, per slot

	/*
	 * Delta segments (incremental refresh): child engines on the same
	 * device and stream, each holding the image of the documents added
	 * since the previous build.  dead[g] = ids removed from segment g
	 * (0 = this engine's own image) after it was built.
	 */
	std::vector<nxsb_engine *> segs;
	std::vector<uint64_t> dead[NXSB_MAX_SEGMENTS + 1];
	uint32_t	max_dead = 0, n_dead = 0;
	unsigned long long *d_dead = nullptr;
	uint32_t *	d_dead_off = nullptr;
	struct SegRun {
		Batch	own;		// this engine's image at limit + max_dead
		DevBuf	gather;		// [segment][query][limit + max_dead] records
	};
	SegRun		seg_oneshot, seg_pipe[PIPE_DEPTH];

	/* timing of the last run */
	/*
	 * A ring of per-run event sets, so that a caller can time many
	 * back-to-back runs without synchronising in between.
	 */
	struct RunEvents {
		cudaEvent_t	ev[EV_PER_RUN] = { nullptr };
		const char *	names[EV_PER_RUN] = { nullptr };
		int		n = 0;
	};
	RunEvents	runs[EV_RING];
	uint64_t	n_runs = 0;		// runs started so far
};

static int
fail(nxsb_engine_t *e, const char *fmt, ...)
{
	va_list ap;

	va_start(ap, fmt);
	vsnprintf(e->err, sizeof(e->err), fmt, ap);
	va_end(ap);
	snprintf(g_last_error, sizeof(g_last_error), "%s", e->err);
	return -1;
}

#define CK(e, call) do {						\
	cudaError_t _rc = (call);					\
	if (_rc != cudaSuccess)						\
		return fail((e), "%s: %s (%s:%d)", #call,		\
		    cudaGetErrorString(_rc), __FILE__, __LINE__);	\
} while (0)

template <typename T>
static cudaError_t
dev_alloc(T **p, size_t n)
{
	return counted_malloc(reinterpret_cast<void **>(p), (n ? n : 1) * sizeof(T));
}

template <typename T>
static void
dev_free(T *&p)
{
	counted_free(p);
	p = nullptr;
}

/*
 * Engine-wide arenas grow with headroom and are sized from the batch's
 * capacity (slots_for), not from what one batch happens to need, so that a
 * stream of batches of one shape never reallocates (cudaFree synchronises
 * the whole device).
 */
template <typename T>
static int
ensure_arena(nxsb_engine_t *e, T *&p, size_t &cap_bytes, size_t want, const char *what)
{
	if (want <= cap_bytes)
		return 0;
	dev_free(p);
	cap_bytes = 0;
	const size_t sz = want + want / 4 + 4096;

	if (counted_malloc(reinterpret_cast<void **>(&p), sz) != cudaSuccess) {
		p = nullptr;
		return fail(e, "%s allocation (%zu bytes) failed: %s", what, sz,
		    cudaGetErrorString(cudaGetLastError()));
	}
	cap_bytes = sz;
	return 0;
}

/* Query slots (real + virtual) the arenas are sized for. */
static inline uint32_t
virtual_cap(uint32_t n_q)
{
	return std::max(64u, n_q / 4);
}

static inline size_t
slots_for(uint32_t n_q)
{
	return (size_t)n_q + virtual_cap(n_q);
}

/* Dynamic shared memory opt-in and CTAs per SM of a kernel, cached. */
static int
kernel_fit(nxsb_engine_t *e, const void *fn, int threads, size_t smem, int *per_sm)
{
	auto it = e->kfit.find(fn);

	if (it != e->kfit.end() && it->second.smem == smem) {
		*per_sm = it->second.per_sm;
		return 0;
	}
	CK(e, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	CK(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, fn, threads, smem));
	e->kfit[fn] = { smem, *per_sm };
	return 0;
}


/* ---- replicated engines -------------------------------------------------- */

extern "C" nxsb_engine_t *nxsb_engine_create(int device);
extern "C" void nxsb_engine_destroy(nxsb_engine_t *e);

static inline bool
is_multi(const nxsb_engine_t *e)
{
	return !e->replicas.empty();
}

static int
multi_fail(nxsb_engine_t *e, const nxsb_engine_t *child, int r)
{
	return fail(e, "%s %d (device %d): %s", e->sharded ? "shard" : "replica", r, child->device,
	    child->err);
}

#define NOT_ON_REPLICATED(e, what) do {						\
	if (is_multi(e))							\
		return fail((e), what " is not available on a %s engine",	\
		    (e)->sharded ? "sharded" : "replicated");			\
} while (0)

/* fn(replica, index) on every replica, one host thread each; 0 if all returned 0. */
template <typename F>
static int
multi_each(nxsb_engine_t *e, bool parallel, F fn)
{
	const size_t n = e->replicas.size();
	std::vector<int> rc(n, 0);

	if (parallel && n > 1) {
		std::vector<std::thread> th;

		for (size_t r = 0; r < n; r++)
			th.emplace_back([&, r] { rc[r] = fn(e->replicas[r], (int)r); });
		for (auto &t : th)
			t.join();
	} else {
		for (size_t r = 0; r < n; r++)
			rc[r] = fn(e->replicas[r], (int)r);
	}
	for (size_t r = 0; r < n; r++)
		if (rc[r] != 0)
			return multi_fail(e, e->replicas[r], (int)r);
	return 0;
}

static int
multi_search_begin(nxsb_engine_t *e, const nxsb_batch_t *b)
{
	const uint32_t R = (uint32_t)e->replicas.size();
	int s = -1;

	for (int i = 0; i < PIPE_DEPTH; i++)
		if (!e->rep_pipe[i].busy) {
			s = i;
			break;
		}
	if (s < 0)
		return fail(e, "too many searches in flight (%d)", PIPE_DEPTH);
	nxsb_engine::RepSplit &sp = e->rep_pipe[s];

	sp.limit = b->limit;
	sp.q0.assign(R + 1, 0);
	sp.h.assign(R, -1);
	for (uint32_t r = 0; r <= R; r++)
		sp.q0[r] = (uint32_t)((uint64_t)b->n_queries * r / R);
	/* The replicas stage their shares side by side (descriptor layout, H2D, launches). */
	std::vector<std::thread> th;
	auto begin_share = [&](uint32_t r) {
		nxsb_batch_t sub = *b;

		sub.queries = b->queries + sp.q0[r];
		sub.n_queries = sp.q0[r + 1] - sp.q0[r];
		/* Token and program arrays go whole: the queries' offsets stay valid. */
		sp.h[r] = sub.n_queries ? nxsb_engine_search_begin(e->replicas[r], &sub) : -2;
	};
	/* Empty shares cost nothing (a single query wakes one replica, no thread). */
	uint32_t mine = R;
	for (uint32_t r = 0; r < R; r++) {
		if (sp.q0[r + 1] == sp.q0[r])
			sp.h[r] = -2;
		else if (mine == R)
			mine = r;
		else
			th.emplace_back(begin_share, r);
	}
	if (mine < R)
		begin_share(mine);
	for (auto &t : th)
		t.join();
	for (uint32_t r = 0; r < R; r++) {
		if (sp.h[r] != -1)
			continue;
		for (uint32_t x = 0; x < R; x++)
			if (sp.h[x] >= 0)
				nxsb_engine_search_end(e->replicas[x], sp.h[x], nullptr, nullptr, nullptr);
		return multi_fail(e, e->replicas[r], (int)r);
	}
	sp.busy = true;
	return s;
}

static int
multi_search_end(nxsb_engine_t *e, int s, uint32_t *counts, uint64_t *ids, float *scores)
{
	if (s < 0 || s >= PIPE_DEPTH || !e->rep_pipe[s].busy)
		return fail(e, "bad search handle %d", s);
	nxsb_engine::RepSplit &sp = e->rep_pipe[s];
	int rc = 0;

	sp.busy = false;
	for (size_t r = 0; r < e->replicas.size(); r++) {
		const size_t at = sp.q0[r];

		if (sp.h[r] < 0)
			continue;
		if (nxsb_engine_search_end(e->replicas[r], sp.h[r], counts ? counts + at : nullptr,
		    counts ? ids + at * sp.limit : nullptr, counts ? scores + at * sp.limit : nullptr) != 0 &&
		    rc == 0)
			rc = multi_fail(e, e->replicas[r], (int)r);
	}
	return rc;
}

/* ---- sharded engines ----------------------------------------------------- */

extern "C" int nxsb_engine_load_shard(nxsb_engine_t *e, const nxsb_shard_desc_t *sd);
extern "C" int nxsb_engine_get_df(nxsb_engine_t *e, uint32_t *df, uint32_t n_terms);
extern "C" int nxsb_engine_set_global_stats(nxsb_engine_t *e, const uint32_t *df,
    uint32_t n_terms, uint64_t token_count, uint32_t doc_count);
extern "C" int nxsb_engine_segment_add(nxsb_engine_t *e, const nxsb_shard_desc_t *sd);
extern "C" int nxsb_engine_set_dead(nxsb_engine_t *e, uint32_t segment, const uint64_t *ids,
    uint32_t n);
extern "C" int nxsb_engine_search_begin(nxsb_engine_t *e, const nxsb_batch_t *b);
extern "C" int nxsb_engine_search_begin_dev(nxsb_engine_t *e, const nxsb_batch_t *b, void *d_recs);
extern "C" int nxsb_engine_sync(nxsb_engine_t *e);
extern "C" int nxsb_engine_search_end(nxsb_engine_t *e, int s, uint32_t *counts, uint64_t *ids,
    float *scores);

/*
 * The documents of sd, ascending by id, are cut into one contiguous range per
 * child with about the same number of postings each; every child builds its
 * range as a shard of its own (the whole-index df / N / token count of sd, or
 * -- when sd carries no df -- the sum of the children's, as the all-reduce of
 * nxsearch_b200/dist.py does between processes).
 */
static int
shard_load(nxsb_engine_t *e, const nxsb_shard_desc_t *sd)
{
	const uint32_t R = (uint32_t)e->replicas.size(), N = sd->n_docs;
	const bool raw = sd->raw != nullptr;
	std::vector<uint32_t> cut(R + 1, N);
	std::vector<std::vector<uint64_t>> rebased(R);
	uint64_t P = 0, acc = 0;

	if (!raw && N && !sd->doc_off)
		return fail(e, "load_shard: neither raw blocks nor doc_off/pairs given");
	for (uint32_t d = 0; d < N; d++)
		P += raw ? sd->raw_n[d] : sd->doc_off[d + 1] - sd->doc_off[d];
	cut[0] = 0;
	for (uint32_t d = 0, r = 1; d < N && r < R; d++) {
		/* Child r starts at the first document at or past its share of the postings. */
		while (r < R && acc >= (P * r + R - 1) / R)
			cut[r++] = d;
		acc += raw ? sd->raw_n[d] : sd->doc_off[d + 1] - sd->doc_off[d];
	}
	e->shard_last_id.assign(R, 0);
	for (uint32_t r = 0; r < R; r++)
		e->shard_last_id[r] = cut[r + 1] > cut[r] ? sd->doc_ids[cut[r + 1] - 1]
		    : (r ? e->shard_last_id[r - 1] : 0);
	e->seg_at.clear();

	if (multi_each(e, true, [&](nxsb_engine_t *c, int r) {
		const uint32_t a = cut[r], n = cut[r + 1] - cut[r];
		nxsb_shard_desc_t sub = *sd;

		sub.n_docs = n;
		sub.doc_ids = sd->doc_ids + a;
		sub.doc_len = sd->doc_len + a;
		if (raw) {
			sub.raw_off = sd->raw_off + a;
			sub.raw_n = sd->raw_n + a;
		} else if (sd->doc_off) {
			std::vector<uint64_t> &off = rebased[r];

			off.resize((size_t)n + 1);
			for (uint32_t i = 0; i <= n; i++)
				off[i] = sd->doc_off[a + i] - sd->doc_off[a];
			sub.doc_off = off.data();
			sub.pairs = sd->pairs + 2 * sd->doc_off[a];
		}
		return nxsb_engine_load_shard(c, &sub);
	}) != 0)
		return -1;
	if (!sd->df) {
		std::vector<uint32_t> df(sd->n_terms);

		if (nxsb_engine_get_df(e, df.data(), sd->n_terms) != 0 ||
		    nxsb_engine_set_global_stats(e, df.data(), sd->n_terms, sd->token_count,
		    sd->doc_count) != 0)
			return -1;
	}
	return 0;
}

/* Removed ids of the base image go to the child whose range holds them. */
static int
shard_set_dead(nxsb_engine_t *e, uint32_t segment, const uint64_t *ids, uint32_t n)
{
	const uint32_t R = (uint32_t)e->replicas.size();

	if (segment > e->seg_at.size())
		return fail(e, "set_dead: no such segment %u", segment);
	if (segment) {
		const nxsb_engine::SegAt at = e->seg_at[segment - 1];
		nxsb_engine_t *c = e->replicas[at.child];

		return nxsb_engine_set_dead(c, at.local, ids, n) == 0 ? 0 : multi_fail(e, c, (int)at.child);
	}
	for (uint32_t i = 1; i < n; i++)
		if (ids[i - 1] >= ids[i])
			return fail(e, "set_dead: ids must be strictly ascending");
	uint32_t a = 0;
	for (uint32_t r = 0; r < R; r++) {
		const uint32_t z = r + 1 == R ? n
		    : (uint32_t)(std::upper_bound(ids + a, ids + n, e->shard_last_id[r]) - ids);

		if (nxsb_engine_set_dead(e->replicas[r], 0, ids + a, z - a) != 0)
			return multi_fail(e, e->replicas[r], (int)r);
		a = z;
	}
	return 0;
}

/* A delta segment goes whole to one child, round robin. */
static int
shard_segment_add(nxsb_engine_t *e, const nxsb_shard_desc_t *sd)
{
	if (e->seg_at.size() >= NXSB_MAX_SEGMENTS)
		return fail(e, "segment_add: %d delta segments already", NXSB_MAX_SEGMENTS);
	const uint32_t r = (uint32_t)(e->seg_at.size() % e->replicas.size());
	nxsb_engine_t *c = e->replicas[r];
	const int local = nxsb_engine_segment_add(c, sd);

	if (local < 0)
		return multi_fail(e, c, (int)r);
	e->seg_at.push_back({ r, (uint32_t)local });
	return (int)e->seg_at.size();
}

static int
shard_search_begin(nxsb_engine_t *e, const nxsb_batch_t *b)
{
	const uint32_t R = (uint32_t)e->replicas.size();
	const int dev0 = e->replicas[0]->device;
	int s = -1;

	for (int i = 0; i < PIPE_DEPTH; i++)
		if (!e->shard_pipe[i].busy) {
			s = i;
			break;
		}
	if (s < 0)
		return fail(e, "too many searches in flight (%d)", PIPE_DEPTH);
	if (b->limit == 0)
		return fail(e, "limit must be at least 1");
	nxsb_engine::ShardRun &S = e->shard_pipe[s];
	const size_t nq = std::max(b->n_queries, 1u);
	const size_t part_bytes = nq * b->limit * sizeof(Rec);

	CK(e, cudaSetDevice(dev0));
	if (!S.done)
		CK(e, cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming));
	S.limit = b->limit;
	S.n_q = b->n_queries;
	S.counts_at = align16(part_bytes);
	S.out_bytes = S.counts_at + nq * 4;
	if (S.gather.ensure(part_bytes * R) || S.out.ensure(S.out_bytes) || S.h_out.ensure(S.out_bytes))
		return fail(e, "allocation failed for a %u-shard search (%u queries, limit %u): %s", R,
		    b->n_queries, b->limit, cudaGetErrorString(cudaGetLastError()));
	S.h.assign(R, -1);
	S.part.resize(R);
	S.part_done.resize(R, nullptr);

	/* Every child scores the whole batch on its own documents, side by side. */
	std::vector<int> rc(R, 0);
	auto begin_part = [&](uint32_t r) {
		nxsb_engine_t *c = e->replicas[r];

		rc[r] = -1;
		if (cudaSetDevice(c->device) != cudaSuccess || S.part[r].ensure(part_bytes) != cudaSuccess ||
		    (!S.part_done[r] && cudaEventCreateWithFlags(&S.part_done[r],
		    cudaEventDisableTiming) != cudaSuccess)) {
			snprintf(c->err, sizeof(c->err), "record buffer: %s",
			    cudaGetErrorString(cudaGetLastError()));
			return;
		}
		if ((S.h[r] = nxsb_engine_search_begin_dev(c, b, S.part[r].p)) < 0)
			return;
		/* Behind the child's kernels, on its stream: its list to the first device. */
		if (cudaMemcpyPeerAsync((char *)S.gather.p + part_bytes * r, dev0, S.part[r].p,
		    c->device, part_bytes, c->stream) != cudaSuccess ||
		    cudaEventRecord(S.part_done[r], c->stream) != cudaSuccess) {
			snprintf(c->err, sizeof(c->err), "peer copy: %s",
			    cudaGetErrorString(cudaGetLastError()));
			return;
		}
		rc[r] = 0;
	};
	std::vector<std::thread> th;

	for (uint32_t r = 1; r < R; r++)
		th.emplace_back(begin_part, r);
	begin_part(0);
	for (auto &t : th)
		t.join();
	int bad = -1;
	for (uint32_t r = 0; r < R; r++)
		if (rc[r] != 0 && bad < 0)
			bad = (int)r;
	if (bad >= 0) {
		for (uint32_t r = 0; r < R; r++)
			if (S.h[r] >= 0)
				nxsb_engine_search_end(e->replicas[r], S.h[r], nullptr, nullptr, nullptr);
		return multi_fail(e, e->replicas[bad], bad);
	}

	CK(e, cudaSetDevice(dev0));
	for (uint32_t r = 0; r < R; r++)
		CK(e, cudaStreamWaitEvent(e->shard_stream, S.part_done[r], 0));
	CK(e, cudaMemsetAsync(S.out.p, 0, S.out_bytes, e->shard_stream));
	const unsigned long long total = (unsigned long long)b->n_queries * R * b->limit;
	if (total) {
		/* (score desc, id desc) over all lists; the ids of the shards never meet. */
		merge_segments_kernel<<<(unsigned)((total + 255) / 256), 256, 0, e->shard_stream>>>(
		    (const Rec *)S.gather.p, R, b->n_queries, b->limit, b->limit, (Rec *)S.out.p,
		    (uint32_t *)((char *)S.out.p + S.counts_at));
		e->launches++;
		CK(e, cudaGetLastError());
	}
	CK(e, cudaMemcpyAsync(S.h_out.p, S.out.p, S.out_bytes, cudaMemcpyDeviceToHost, e->shard_stream));
	CK(e, cudaEventRecord(S.done, e->shard_stream));
	S.busy = true;
	return s;
}

static int
shard_search_end(nxsb_engine_t *e, int s, uint32_t *counts, uint64_t *ids, float *scores)
{
	if (s < 0 || s >= PIPE_DEPTH || !e->shard_pipe[s].busy)
		return fail(e, "bad search handle %d", s);
	nxsb_engine::ShardRun &S = e->shard_pipe[s];
	int rc = 0;

	S.busy = false;
	CK(e, cudaSetDevice(e->replicas[0]->device));
	CK(e, cudaEventSynchronize(S.done));
	for (size_t r = 0; r < e->replicas.size(); r++)
		if (nxsb_engine_search_end(e->replicas[r], S.h[r], nullptr, nullptr, nullptr) != 0 && rc == 0)
			rc = multi_fail(e, e->replicas[r], (int)r);
	if (rc == 0 && counts) {
		const Rec *recs = (const Rec *)S.h_out.p;
		const size_t nrec = (size_t)S.n_q * S.limit;

		memcpy(counts, (const char *)S.h_out.p + S.counts_at, (size_t)S.n_q * 4);
		for (size_t i = 0; i < nrec; i++) {
			ids[i] = recs[i].doc_id;
			scores[i] = recs[i].score;
		}
	}
	return rc;
}

static void
shard_release(nxsb_engine_t *e)
{
	for (auto &S : e->shard_pipe) {
		for (size_t r = 0; r < S.part.size(); r++) {
			cudaSetDevice(e->replicas[r]->device);
			S.part[r].release();
			if (S.part_done[r])
				cudaEventDestroy(S.part_done[r]);
		}
		S.part.clear();
		S.part_done.clear();
		cudaSetDevice(e->replicas[0]->device);
		S.gather.release();
		S.out.release();
		S.h_out.release();
		if (S.done)
			cudaEventDestroy(S.done);
		S.done = nullptr;
	}
	if (e->shard_stream) {
		cudaSetDevice(e->replicas[0]->device);
		cudaStreamDestroy(e->shard_stream);
		e->shard_stream = nullptr;
	}
}

extern "C" nxsb_engine_t *nxsb_engine_create_replicated(const int *devices, int n);

extern "C" nxsb_engine_t *
nxsb_engine_create_sharded(const int *devices, int n)
{
	/* A device may be listed twice (two ranges on it): what a one-GPU box can test. */
	nxsb_engine_t *e = nxsb_engine_create_replicated(devices, n);

	if (!e)
		return nullptr;
	e->sharded = true;
	/* The lists travel device to device where the fabric allows it (else through the host). */
	for (int i = 1; i < n; i++) {
		int can = 0;

		if (cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) == cudaSuccess && can &&
		    cudaSetDevice(devices[i]) == cudaSuccess)
			cudaDeviceEnablePeerAccess(devices[0], 0);
		cudaGetLastError();
	}
	if (cudaSetDevice(devices[0]) != cudaSuccess ||
	    cudaStreamCreateWithFlags(&e->shard_stream, cudaStreamNonBlocking) != cudaSuccess) {
		snprintf(g_last_error, sizeof(g_last_error), "sharded engine: %s",
		    cudaGetErrorString(cudaGetLastError()));
		nxsb_engine_destroy(e);
		return nullptr;
	}
	return e;
}

extern "C" int
nxsb_engine_is_sharded(const nxsb_engine_t *e)
{
	return e->sharded ? 1 : 0;
}

extern "C" nxsb_engine_t *
nxsb_engine_create_replicated(const int *devices, int n)
{
	if (n < 1) {
		snprintf(g_last_error, sizeof(g_last_error), "a replicated engine needs at least one device");
		return nullptr;
	}
	nxsb_engine_t *e = new nxsb_engine();

	e->device = devices[0];
	for (int i = 0; i < n; i++) {
		nxsb_engine_t *c = nxsb_engine_create(devices[i]);

		if (!c) {
			for (auto *x : e->replicas)
				nxsb_engine_destroy(x);
			delete e;
			return nullptr;
		}
		e->replicas.push_back(c);
	}
	return e;
}

extern "C" int
nxsb_engine_replica_count(const nxsb_engine_t *e)
{
	return (int)e->replicas.size();
}

extern "C" int
nxsb_gpu_device_count(void)
{
	int n = 0;

	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

extern "C" const char *
nxsb_last_error(void)
{
	return g_last_error;
}

extern "C" const char *
nxsb_engine_errmsg(const nxsb_engine_t *e)
{
	return e ? e->err : g_last_error;
}

#ifdef ST_PROF
extern "C" __attribute__((visibility("default"))) int
nxsb_engine_prof(nxsb_engine_t *e, unsigned long long *out)
{
	cudaStreamSynchronize(e->stream);
	cudaMemcpy(out, e->d_prof, 32 * 8, cudaMemcpyDeviceToHost);
	cudaMemset(e->d_prof, 0, 32 * 8);
	return 0;
}
#endif

extern "C" uint64_t
nxsb_engine_launch_count(const nxsb_engine_t *e)
{
	if (is_multi(e)) {
		uint64_t n = 0;

		for (auto *c : e->replicas)
			n += nxsb_engine_launch_count(c);
		return n + e->launches;	/* + the cross-shard merges */
	}
	uint64_t n = e->launches;

	for (const nxsb_engine *c : e->segs)
		n += c->launches;
	return n;
}

extern "C" int
nxsb_engine_set_pruning(nxsb_engine_t *e, int on)
{
	if (is_multi(e)) {
		int was = 0;

		for (auto *c : e->replicas)
			was = nxsb_engine_set_pruning(c, on);
		return was;
	}
	const int was = e->bmw_enabled;

	e->bmw_enabled = on != 0;
	for (nxsb_engine *c : e->segs)
		c->bmw_enabled = e->bmw_enabled;
	return was;
}

static_assert(NXSB_KTH_STEPS == BMW_LADDER, "header and kernel agree on the ladder");

/* One thread per (tf, dl, idf) triple through the scorers' own arithmetic. */
__global__ void __launch_bounds__(256)
score_pairs_kernel(int algo, uint32_t n, const uint32_t *__restrict__ tf,
    const uint32_t *__restrict__ dl, const float *__restrict__ idf,
    const float *__restrict__ logtab, float K0, float K1, float *__restrict__ out)
{
	__shared__ float s_logtab[LOGTAB_N];
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

	for (uint32_t x = threadIdx.x; x < LOGTAB_N; x += blockDim.x)
		s_logtab[x] = logtab[x];
	__syncthreads();
	if (i >= n)
		return;
	StreamParams sp;
	sp.K0 = K0;
	sp.K1 = K1;
	sp.doc_len = nullptr;
	const uint2 v[1] = { make_uint2(0u, (tf[i] & 0xffffu) | (dl[i] << 16)) };
	float sc[1];

	if (algo == NXSB_ALGO_BM25)
		st_score<false, NXSB_ALGO_BM25, 1>(sp, s_logtab, v, idf[i], sc);
	else
		st_score<false, NXSB_ALGO_TFIDF, 1>(sp, s_logtab, v, idf[i], sc);
	out[i] = sc[0];
}

extern "C" int
nxsb_engine_score_pairs(nxsb_engine_t *e, int algo, uint32_t n, const uint32_t *tf,
    const uint32_t *dl, const float *idf, float *out)
{
	if (is_multi(e))
		return nxsb_engine_score_pairs(e->replicas[0], algo, n, tf, dl, idf, out) == 0 ? 0
		    : multi_fail(e, e->replicas[0], 0);
	uint32_t *d_tf = nullptr, *d_dl = nullptr;
	float *d_idf = nullptr, *d_out = nullptr;
	int rc = -1;

	if (!e->loaded)
		return fail(e, "no image loaded (the scores use its K0 / K1)");
	CK(e, cudaSetDevice(e->device));
	do {
		if (n == 0) {
			rc = 0;
			break;
		}
		if (cudaMalloc(&d_tf, (size_t)n * 4) || cudaMalloc(&d_dl, (size_t)n * 4) ||
		    cudaMalloc(&d_idf, (size_t)n * 4) || cudaMalloc(&d_out, (size_t)n * 4))
			break;
		cudaMemcpyAsync(d_tf, tf, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream);
		cudaMemcpyAsync(d_dl, dl, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream);
		cudaMemcpyAsync(d_idf, idf, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream);
		score_pairs_kernel<<<(n + 255) / 256, 256, 0, e->stream>>>(algo, n, d_tf, d_dl, d_idf,
		    e->d_logtab, e->K0, e->K1, d_out);
		cudaMemcpyAsync(out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, e->stream);
		if (cudaStreamSynchronize(e->stream) != cudaSuccess)
			break;
		rc = 0;
	} while (0);
	cudaFree(d_tf); cudaFree(d_dl); cudaFree(d_idf); cudaFree(d_out);
	return rc == 0 ? 0 : fail(e, "score probe failed: %s", cudaGetErrorString(cudaGetLastError()));
}

extern "C" int
nxsb_engine_term_kth(nxsb_engine_t *e, int algo, const uint32_t *term_ids, uint32_t n, float *out)
{
	if (is_multi(e))
		return nxsb_engine_term_kth(e->replicas[0], algo, term_ids, n, out) == 0 ? 0
		    : multi_fail(e, e->replicas[0], 0);
	const float *tab = algo == NXSB_ALGO_BM25 ? e->d_kth_bm25 : e->d_kth_tfidf;

	if (!e->loaded || !tab)
		return fail(e, "no k-th weight tables (no image, a wide image, or pruning off)");
	CK(e, cudaSetDevice(e->device));
	CK(e, cudaStreamSynchronize(e->stream));
	for (uint32_t i = 0; i < n; i++) {
		float *o = out + (size_t)i * BMW_LADDER;

		if (term_ids[i] == 0 || term_ids[i] > e->n_terms) {
			memset(o, 0, BMW_LADDER * sizeof(float));
			continue;
		}
		CK(e, cudaMemcpy(o, tab + (size_t)(term_ids[i] - 1) * BMW_LADDER,
		    BMW_LADDER * sizeof(float), cudaMemcpyDeviceToHost));
	}
	return 0;
}

extern "C" int
nxsb_engine_pruning_stats(nxsb_engine_t *e, uint64_t out[16], int reset)
{
	if (is_multi(e)) {
		uint64_t one[16];

		memset(out, 0, 16 * sizeof(uint64_t));
		for (size_t r = 0; r < e->replicas.size(); r++) {
			if (nxsb_engine_pruning_stats(e->replicas[r], one, reset) != 0)
				return multi_fail(e, e->replicas[r], (int)r);
			for (int i = 0; i < 16; i++)
				out[i] += one[i];
		}
		return 0;
	}
	unsigned long long v[16];

	CK(e, cudaSetDevice(e->device));
	CK(e, cudaStreamSynchronize(e->stream));
	CK(e, cudaMemcpy(v, e->d_bmw_stats, sizeof(v), cudaMemcpyDeviceToHost));
	for (int i = 0; i < 16; i++)
		out[i] = v[i];
	for (nxsb_engine *c : e->segs) {
		uint64_t sub[16];

		if (nxsb_engine_pruning_stats(c, sub, reset) == -1)
			return fail(e, "%s", c->err);
		for (int i = 0; i < 16; i++)
			out[i] += sub[i];
	}
	if (reset)
		CK(e, cudaMemset(e->d_bmw_stats, 0, sizeof(v)));
	return 0;
}

extern "C" nxsb_engine_t *
nxsb_engine_create(int device)
{
	int n = 0;
	cudaError_t rc = cudaGetDeviceCount(&n);

	if (rc != cudaSuccess || n == 0) {
		snprintf(g_last_error, sizeof(g_last_error),
		    "no CUDA device available (%s); the nxsearch-b200 engine has "
		    "no CPU fallback", rc != cudaSuccess ? cudaGetErrorString(rc) :
		    "device count is 0");
		cudaGetLastError();
		return nullptr;
	}
	if (device < 0 || device >= n) {
		snprintf(g_last_error, sizeof(g_last_error),
		    "CUDA device %d out of range (have %d)", device, n);
		return nullptr;
	}
	nxsb_engine_t *e = new nxsb_engine();
	cudaDeviceProp prop;

	e->device = device;
	if (cudaSetDevice(device) != cudaSuccess ||
	    cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
		snprintf(g_last_error, sizeof(g_last_error), "CUDA init failed: %s",
		    cudaGetErrorString(cudaGetLastError()));
		delete e;
		return nullptr;
	}
	e->stream = e->own_stream;
	e->n_sms = prop.multiProcessorCount;
#ifdef ST_PROF
	dev_alloc(&e->d_prof, 32);
	cudaMemset(e->d_prof, 0, 32 * 8);
#endif
	{
		/* Development switch: score with the older tiles.cuh kernel. */
		const char *kv = getenv("NXSB_KERNEL");
		e->force_v2 = kv && strcmp(kv, "v2") == 0;
		/* Development switch: density threshold of the dense columns (> 1: none). */
		if ((kv = getenv("NXSB_DENSE_MIN")) != NULL)
			e->dense_min = (float)atof(kv);
		/* Development switch: every query streams its own dense columns. */
		if ((kv = getenv("NXSB_SHARE_DENSE")) != NULL)
			e->share_dense = atoi(kv) != 0;
		/* NXSB_BMW=0: no block-max pruning, every query streams all its postings. */
		if ((kv = getenv("NXSB_BMW")) != NULL)
			e->bmw_enabled = atoi(kv) != 0;
		if ((kv = getenv("NXSB_BMW_LOGIC_POS")) != NULL)
			e->logic_pruned_max_pos = (uint32_t)atoi(kv);
		/* NXSB_LANES=1: every search on one stream (A/B of the overlap). */
		if ((kv = getenv("NXSB_LANES")) != NULL)
			e->lanes_enabled = atoi(kv) > 1;
		/* NXSB_PRIME=0: pruning thresholds start at zero (A/B of the priming). */
		if ((kv = getenv("NXSB_PRIME")) != NULL)
			e->prime_enabled = atoi(kv) != 0;
		/* Development switch: documents per block = 2^NXSB_BMW_SHIFT. */
		if ((kv = getenv("NXSB_BMW_SHIFT")) != NULL)
			e->bshift = (uint32_t)std::min(BMW_SHIFT_MAX, std::max(BMW_SHIFT_MIN, atoi(kv)));
	}
	if (dev_alloc(&e->d_bmw_stats, 16) != cudaSuccess ||
	    cudaMemset(e->d_bmw_stats, 0, 128) != cudaSuccess) {
		snprintf(g_last_error, sizeof(g_last_error), "CUDA alloc failed");
		delete e;
		return nullptr;
	}
	for (auto &r : e->runs)
		for (int i = 0; i < EV_PER_RUN; i++)
			cudaEventCreateWithFlags(&r.ev[i], cudaEventDefault);

	/* (float)log(c + 1): the tf weight of ref ranking.c:90,168. */
	float tab[LOGTAB_N];
	for (int c = 0; c < LOGTAB_N; c++)
		tab[c] = (float)log((double)(c + 1));
	if (dev_alloc(&e->d_logtab, LOGTAB_N) != cudaSuccess ||
	    cudaMemcpy(e->d_logtab, tab, sizeof(tab), cudaMemcpyHostToDevice) != cudaSuccess) {
		snprintf(g_last_error, sizeof(g_last_error), "CUDA alloc failed");
		delete e;
		return nullptr;
	}
	return e;
}

/* Park the current lane's members, bring lane i's in (creating its stream and
 * score columns on first use). */
static int
use_lane(nxsb_engine_t *e, int i)
{
	if (i != e->cur_lane) {
		nxsb_engine::Lane &o = e->lanes[e->cur_lane], &n = e->lanes[i];

		o.stream = e->own_stream;
#define LANE_SWAP(f)	do { o.f = e->f; e->f = n.f; } while (0)
		LANE_SWAP(d_cand); LANE_SWAP(cand_bytes); LANE_SWAP(d_sort_tmp); LANE_SWAP(sort_tmp_bytes);
		LANE_SWAP(d_cub_tmp); LANE_SWAP(cub_tmp_bytes); LANE_SWAP(d_plan); LANE_SWAP(plan_bytes);
		LANE_SWAP(d_tile_cnt); LANE_SWAP(tile_cnt_bytes); LANE_SWAP(d_tt); LANE_SWAP(tt_bytes);
		LANE_SWAP(d_dense_sc); LANE_SWAP(d_dense_used);
#undef LANE_SWAP
		if (!n.stream && cudaStreamCreateWithFlags(&n.stream, cudaStreamNonBlocking) != cudaSuccess)
			return fail(e, "stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
		e->own_stream = n.stream;
		e->cur_lane = i;
	}
	if (!e->external_stream)
		e->stream = e->own_stream;
	if (e->n_dense && !e->d_dense_sc) {
		/* This lane's per-batch score columns (load_shard made lane 0's). */
		const size_t col_words = (size_t)e->ntiles * TILE_DOCS;

		if (dev_alloc(&e->d_dense_sc, (size_t)e->n_dense * col_words) != cudaSuccess ||
		    dev_alloc(&e->d_dense_used, 512) != cudaSuccess)
			return fail(e, "score column allocation failed");
	}
	return 0;
}

/* Before the image changes: nothing may still be reading it on either lane. */
static void
sync_lanes(nxsb_engine_t *e)
{
	cudaSetDevice(e->device);
	if (e->own_stream)
		cudaStreamSynchronize(e->own_stream);
	for (int i = 0; i < N_LANES; i++)
		if (i != e->cur_lane && e->lanes[i].stream)
			cudaStreamSynchronize(e->lanes[i].stream);
	if (e->external_stream)
		cudaStreamSynchronize(e->stream);
}

static void
free_image(nxsb_engine_t *e)
{
	for (int i = 0; i < N_LANES; i++) {
		if (i == e->cur_lane)
			continue;
		dev_free(e->lanes[i].d_dense_sc);
		dev_free(e->lanes[i].d_dense_used);
	}
	dev_free(e->d_post);
	dev_free(e->d_term_off);
	dev_free(e->d_df_local);
	dev_free(e->d_idf_bm25);
	dev_free(e->d_idf_tfidf);
	dev_free(e->d_skip_row);
	dev_free(e->d_dense);
	dev_free(e->d_dense_col);
	dev_free(e->d_dense_sc);
	dev_free(e->d_dense_terms);
	dev_free(e->d_dense_used);
	e->n_dense = 0;
	dev_free(e->d_bcol);
	dev_free(e->d_bcol_terms);
	dev_free(e->d_boff);
	dev_free(e->d_bmax_bm25);
	dev_free(e->d_bmax_tfidf);
	dev_free(e->d_smax_bm25);
	dev_free(e->d_smax_tfidf);
	dev_free(e->d_mt_bm25);
	dev_free(e->d_mt_tfidf);
	dev_free(e->d_mt_bits);
	dev_free(e->d_long_terms);
	dev_free(e->d_wmax_bm25);
	dev_free(e->d_wmax_tfidf);
	dev_free(e->d_kth_bm25);
	dev_free(e->d_kth_tfidf);
	dev_free(e->d_kth_units);
	dev_free(e->d_kth_longs);
	dev_free(e->d_kth_scratch);
	e->n_kth_units = e->n_kth_longs = 0;
	e->tables_valid = false;
	e->n_bcol = 0;
	dev_free(e->d_skip);
	dev_free(e->d_skip_mt);
	dev_free(e->d_doc_ids);
	dev_free(e->d_doc_len);
	e->loaded = false;
}

static void drop_segments(nxsb_engine_t *e);

static void
free_batch(Batch &b)
{
	b.desc.release();
	b.h_desc.release();
	b.scratch.release();
	b.results.release();
	b.h_results.release();
	b = Batch();
}

/* Forget the delta segments and every removal note (the stream is idle). */
static void
drop_segments(nxsb_engine_t *e)
{
	for (nxsb_engine *c : e->segs) {
		e->launches += c->launches;
		c->stream = c->own_stream;
		nxsb_engine_destroy(c);
	}
	e->segs.clear();
	for (auto &d : e->dead)
		d.clear();
	e->max_dead = e->n_dead = 0;
	dev_free(e->d_dead);
	dev_free(e->d_dead_off);
}

extern "C" void
nxsb_engine_destroy(nxsb_engine_t *e)
{
	if (e && is_multi(e)) {
		for (auto *c : e->replicas)
			nxsb_engine_sync(c);
		shard_release(e);
		for (auto *c : e->replicas)
			nxsb_engine_destroy(c);
		delete e;
		return;
	}
	if (!e)
		return;
	sync_lanes(e);
	drop_segments(e);
	free_batch(e->seg_oneshot.own);
	e->seg_oneshot.gather.release();
	for (auto &sr : e->seg_pipe) {
		free_batch(sr.own);
		sr.gather.release();
	}
	for (auto &b : e->batches)
		if (b.used)
			free_batch(b);
	free_batch(e->oneshot);
	for (int i = 0; i < PIPE_DEPTH; i++) {
		free_batch(e->pipe[i]);
		if (e->pipe_done[i])
			cudaEventDestroy(e->pipe_done[i]);
	}
	free_image(e);
	fuzzy_free(e->fz);
	fuzzy_scratch_free(e->fz_scratch);
	for (auto &z : e->fz_pipe)
		fuzzy_scratch_free(z);
	for (int i = 0; i < N_LANES; i++) {
		nxsb_engine::Lane &l = e->lanes[i];

		if (i == e->cur_lane)
			continue;
		dev_free(l.d_cand);
		dev_free(l.d_sort_tmp);
		dev_free(l.d_plan);
		dev_free(l.d_tile_cnt);
		dev_free(l.d_tt);
		if (l.d_cub_tmp)
			cudaFree(l.d_cub_tmp);
		if (l.stream) {
			cudaStreamSynchronize(l.stream);
			cudaStreamDestroy(l.stream);
		}
	}
	dev_free(e->d_cand);
	dev_free(e->d_sort_tmp);
	dev_free(e->d_plan);
	dev_free(e->d_tile_cnt);
	dev_free(e->d_tt);
	if (e->d_cub_tmp)
		cudaFree(e->d_cub_tmp);
	dev_free(e->d_logtab);
	dev_free(e->d_bmw_stats);
	for (auto &r : e->runs)
		for (int i = 0; i < EV_PER_RUN; i++)
			if (r.ev[i])
				cudaEventDestroy(r.ev[i]);
	if (e->own_stream)
		cudaStreamDestroy(e->own_stream);
	delete e;
}

extern "C" int
nxsb_engine_set_stream(nxsb_engine_t *e, void *s)
{
	NOT_ON_REPLICATED(e, "an external stream");
	sync_lanes(e);
	if (use_lane(e, 0) == -1)
		return -1;
	e->external_stream = s != nullptr;
	e->stream = s ? (cudaStream_t)s : e->own_stream;
	for (nxsb_engine *c : e->segs)
		c->stream = e->stream;
	return 0;
}

extern "C" void *
nxsb_engine_lane_stream(nxsb_engine_t *e, int lane)
{
	if (is_multi(e) || lane < 0 || lane >= N_LANES)
		return nullptr;
	cudaSetDevice(e->device);
	if (lane == e->cur_lane)
		return e->own_stream;
	if (!e->lanes[lane].stream &&
	    cudaStreamCreateWithFlags(&e->lanes[lane].stream, cudaStreamNonBlocking) != cudaSuccess)
		return nullptr;
	return e->lanes[lane].stream;
}

extern "C" int
nxsb_engine_lanes_join(nxsb_engine_t *e)
{
	NOT_ON_REPLICATED(e, "a lane join");
	CK(e, cudaSetDevice(e->device));
	cudaStream_t s0 = (cudaStream_t)nxsb_engine_lane_stream(e, 0);

	for (int i = 1; i < N_LANES; i++) {
		cudaStream_t si = i == e->cur_lane ? e->own_stream : e->lanes[i].stream;
		cudaEvent_t ev;

		if (!si)
			continue;
		CK(e, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
		CK(e, cudaEventRecord(ev, si));
		CK(e, cudaStreamWaitEvent(s0, ev, 0));
		CK(e, cudaEventDestroy(ev));
	}
	return 0;
}

extern "C" int
nxsb_engine_sync(nxsb_engine_t *e)
{
	if (is_multi(e)) {
		if (multi_each(e, false, [](nxsb_engine_t *c, int) { return nxsb_engine_sync(c); }) != 0)
			return -1;
		if (e->shard_stream) {
			CK(e, cudaSetDevice(e->replicas[0]->device));
			CK(e, cudaStreamSynchronize(e->shard_stream));
		}
		return 0;
	}
	sync_lanes(e);
	CK(e, cudaGetLastError());
	return 0;
}

/*
 * Global statistics -> per-term idf tables and the BM25 constants, in
 * double on the host with the exact operand types of the reference
 * (ranking.c:91: float division inside a double log; ranking.c:163: integer
 * quotient; ranking.c:141-142: float literals widened to double).
 */
static int
upload_stats(nxsb_engine_t *e)
{
	const uint32_t V = e->n_terms;
	const unsigned long doc_count = e->doc_count;
	std::vector<float> bm(V), tfidf(V);

	e->stats_valid = false;
	if (doc_count) {
		const double adl = (double)(e->token_count / doc_count);
		static const double k = 1.2f, b = 0.75f;

		if (adl >= 1) {
			e->K0 = (float)(k * (1 - b));
			e->K1 = (float)(k * b / adl);
			e->stats_valid = true;
		}
	}
	/* Two libm log() calls per term: spread over host threads (an index
	 * refresh re-derives every term's idf because N moved). */
	auto fill = [&](uint32_t t0, uint32_t t1) {
		for (uint32_t t = t0; t < t1; t++) {
			const unsigned long df = e->h_df[t];

			if (!df || !doc_count) {
				bm[t] = tfidf[t] = 0.f;
				continue;
			}
			bm[t] = (float)log(((doc_count - df + 0.5) / (df + 0.5)) + 1);
			/*
			 * C semantics: float quotient, widened, DOUBLE log (in C++ a
			 * float argument would select logf and lose the last ulp).
			 */
			tfidf[t] = (float)(log((double)((float)doc_count / (float)df)) + 1);
		}
	};
	{
		const unsigned hw = std::thread::hardware_concurrency();
		const uint32_t nthr = V < 65536 ? 1 : std::min(8u, hw ? hw : 1u);
		std::vector<std::thread> pool;

		for (uint32_t i = 1; i < nthr; i++)
			pool.emplace_back(fill, (uint32_t)((uint64_t)V * i / nthr),
			    (uint32_t)((uint64_t)V * (i + 1) / nthr));
		fill(0, (uint32_t)((uint64_t)V / nthr));
		for (auto &th : pool)
			th.join();
	}
	CK(e, cudaMemcpyAsync(e->d_idf_bm25, bm.data(), V * sizeof(float),
	    cudaMemcpyHostToDevice, e->stream));
	CK(e, cudaMemcpyAsync(e->d_idf_tfidf, tfidf.data(), V * sizeof(float),
	    cudaMemcpyHostToDevice, e->stream));
	if (e->d_wmax_bm25 && !(e->tables_valid && e->tables_K0 == e->K0 && e->tables_K1 == e->K1)) {
		/*
		 * The BM25 weight depends on K0 / K1: the maxima follow the
		 * statistics.  K1 moves only when the integer average length does
		 * (ref ranking.c:163), so most refreshes keep the tables.
		 */
		const size_t words = (size_t)e->n_bcol * e->row_stride;

		if (e->n_bcol) {
			CK(e, cudaMemsetAsync(e->d_bmax_bm25, 0, words * 4, e->stream));
			CK(e, cudaMemsetAsync(e->d_bmax_tfidf, 0, words * 4, e->stream));
			block_max_kernel<<<dim3(16, e->n_bcol), 256, 0, e->stream>>>(e->d_post,
			    e->d_term_off, e->d_bcol_terms, e->row_stride, e->bshift, e->d_logtab,
			    e->K0, e->K1, e->d_bmax_bm25, e->d_bmax_tfidf);
			superblock_max_kernel<<<e->n_sms * 4, 256, 0, e->stream>>>(e->d_bmax_bm25,
			    e->d_bmax_tfidf, e->n_bcol, e->row_stride, e->sb_stride,
			    e->d_smax_bm25, e->d_smax_tfidf);
			e->launches += 2;
		}
		term_wmax_kernel<<<e->n_sms * 8, 256, 0, e->stream>>>(e->d_post, e->d_term_off,
		    e->d_bcol, V, e->d_logtab, e->K0, e->K1, e->d_bmax_bm25, e->d_bmax_tfidf,
		    e->row_stride, e->d_wmax_bm25, e->d_wmax_tfidf);
		e->launches++;
		if (e->d_mt_bm25) {
			const size_t bytes = (size_t)e->n_long * e->mt_stride;

			CK(e, cudaMemsetAsync(e->d_mt_bm25, 0, bytes, e->stream));
			CK(e, cudaMemsetAsync(e->d_mt_tfidf, 0, bytes, e->stream));
			CK(e, cudaMemsetAsync(e->d_mt_bits, 0, bytes * 4, e->stream));
			minitile_max_kernel<<<e->n_sms * 16, 256, 0, e->stream>>>(e->d_post, e->d_term_off,
			    e->d_long_terms, e->d_bcol, e->d_skip_mt, e->n_long, e->n_mt, e->mt_stride,
			    e->d_logtab, e->K0, e->K1, e->d_wmax_bm25, e->d_wmax_tfidf,
			    e->d_mt_bm25, e->d_mt_tfidf, e->d_mt_bits, e->bshift);
			e->launches++;
		}
		if (e->d_kth_bm25) {
			const size_t lad = (size_t)V * BMW_LADDER * 4;

			CK(e, cudaMemsetAsync(e->d_kth_bm25, 0, lad, e->stream));
			CK(e, cudaMemsetAsync(e->d_kth_tfidf, 0, lad, e->stream));
			term_kth_kernel<false><<<e->n_sms * 8, 256, 0, e->stream>>>(e->d_post,
			    e->d_term_off, V, nullptr, 0, e->d_logtab, e->K0, e->K1,
			    e->d_kth_bm25, e->d_kth_tfidf, nullptr);
			e->launches++;
			if (e->n_kth_units) {
				term_kth_kernel<true><<<std::min(e->n_sms * 8u, (e->n_kth_units + 7) / 8),
				    256, 0, e->stream>>>(e->d_post, e->d_term_off, V, e->d_kth_units,
				    e->n_kth_units, e->d_logtab, e->K0, e->K1, nullptr, nullptr,
				    e->d_kth_scratch);
				term_kth_merge_kernel<<<std::min(e->n_sms * 8u, (e->n_kth_longs + 7) / 8),
				    256, 0, e->stream>>>(e->d_kth_longs, e->n_kth_longs, e->d_kth_units,
				    e->n_kth_units, e->d_kth_scratch, e->d_kth_bm25, e->d_kth_tfidf);
				e->launches += 2;
			}
		}
		CK(e, cudaGetLastError());
		e->tables_valid = true;
		e->tables_K0 = e->K0;
		e->tables_K1 = e->K1;
	}
	CK(e, cudaStreamSynchronize(e->stream));
	return 0;
}

extern "C" int
nxsb_engine_load_shard(nxsb_engine_t *e, const nxsb_shard_desc_t *sd)
{
	if (is_multi(e) && e->sharded)
		return shard_load(e, sd);
	if (is_multi(e))	/* the same image on every device, built side by side */
		return multi_each(e, true, [&](nxsb_engine_t *c, int) { return nxsb_engine_load_shard(c, sd); });
	/* The image is about to change: both lanes drain first. */
	sync_lanes(e);
	if (use_lane(e, 0) == -1)
		return -1;
	const uint32_t N = sd->n_docs, V = sd->n_terms;
	const bool raw = sd->raw != nullptr;
	std::vector<uint64_t> raw_doc_off, raw_rel;
	uint64_t raw_lo = 0, raw_hi = 0;
	uint2 *d_pairs = nullptr, *d_vals_alt = nullptr;
	unsigned char *d_raw = nullptr;
	unsigned long long *d_doc_off = nullptr, *d_raw_off = nullptr;
	uint32_t *d_keys = nullptr, *d_keys_alt = nullptr, *d_long = nullptr;
	uint32_t *d_maxc = nullptr;
	int rc = -1;

	if (raw) {
		/* Offsets of the decoded pairs, and the byte range to copy. */
		raw_doc_off.resize((size_t)N + 1);
		raw_rel.resize((size_t)N + 1);
		raw_lo = N ? UINT64_MAX : 0;
		uint64_t acc = 0;
		for (uint32_t d = 0; d < N; d++) {
			raw_doc_off[d] = acc;
			acc += sd->raw_n[d];
			if (sd->raw_off[d] & 7)
				return fail(e, "load_shard: raw_off[%u] is not 8-byte aligned", d);
			raw_lo = std::min<uint64_t>(raw_lo, sd->raw_off[d]);
			raw_hi = std::max<uint64_t>(raw_hi, sd->raw_off[d] + 8ull * sd->raw_n[d]);
		}
		raw_doc_off[N] = acc;
		if (raw_hi < raw_lo)
			raw_hi = raw_lo;
		for (uint32_t d = 0; d < N; d++)
			raw_rel[d] = sd->raw_off[d] - raw_lo;
	}
	const uint64_t P = raw ? raw_doc_off[N] : (sd->doc_off ? sd->doc_off[N] : 0);
	const uint64_t *h_doc_off = raw ? raw_doc_off.data() : sd->doc_off;

	CK(e, cudaSetDevice(e->device));
	CK(e, cudaStreamSynchronize(e->stream));
	drop_segments(e);
	for (auto &b : e->batches)
		if (b.used)
			free_batch(b);
	free_image(e);

	e->n_docs = N;
	e->n_terms = V;
	e->n_post = P;
	e->ntiles = (N + TILE_DOCS - 1) / TILE_DOCS;
	if (e->ntiles == 0)
		e->ntiles = 1;
	e->token_count = sd->token_count;
	e->doc_count = sd->doc_count;

	/* Packed postings need every tf and doc length to fit 16 bits. */
	e->wide = false;
	for (uint32_t d = 0; d < N && !e->wide; d++)
		e->wide = sd->doc_len[d] > 0xffffu;
	for (uint64_t j = 0; !raw && j < P && !e->wide; j++)
		e->wide = sd->pairs[2 * j + 1] > 0xffffu;

	do {
		if ((raw ? (dev_alloc(&d_raw, raw_hi - raw_lo + 8) ||
		    dev_alloc(&d_raw_off, (size_t)N + 1) || dev_alloc(&d_maxc, 1))
		    : dev_alloc(&d_pairs, P)) || dev_alloc(&d_doc_off, (size_t)N + 1) ||
		    dev_alloc(&e->d_doc_len, N) || dev_alloc(&e->d_doc_ids, N) ||
		    dev_alloc(&d_keys, P) || dev_alloc(&d_keys_alt, P) ||
		    dev_alloc(&e->d_post, P + 2) || dev_alloc(&d_vals_alt, P + 2) ||
		    dev_alloc(&e->d_term_off, (size_t)V + 1) ||
		    dev_alloc(&e->d_df_local, V) || dev_alloc(&e->d_idf_bm25, V) ||
		    dev_alloc(&e->d_idf_tfidf, V) || dev_alloc(&e->d_skip_row, V)) {
			fail(e, "device allocation failed while loading a shard of "
			    "%u docs / %llu postings: %s", N, (unsigned long long)P,
			    cudaGetErrorString(cudaGetLastError()));
			break;
		}
		cudaStream_t st = e->stream;
		if ((raw ? (cudaMemcpyAsync(d_raw, (const char *)sd->raw + raw_lo, raw_hi - raw_lo, cudaMemcpyHostToDevice, st) ||
		    cudaMemcpyAsync(d_raw_off, raw_rel.data(), (size_t)N * 8, cudaMemcpyHostToDevice, st) ||
		    cudaMemsetAsync(d_maxc, 0, 4, st))
		    : cudaMemcpyAsync(d_pairs, sd->pairs, P * 8, cudaMemcpyHostToDevice, st)) ||
		    cudaMemcpyAsync(d_doc_off, h_doc_off, ((size_t)N + 1) * 8, cudaMemcpyHostToDevice, st) ||
		    cudaMemcpyAsync(e->d_doc_len, sd->doc_len, (size_t)N * 4, cudaMemcpyHostToDevice, st) ||
		    cudaMemcpyAsync(e->d_doc_ids, sd->doc_ids, (size_t)N * 8, cudaMemcpyHostToDevice, st)) {
			fail(e, "H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
			break;
		}

		if (N) {
			const int grid = e->n_sms * 8;
			if (raw && !e->wide) {
				uint32_t maxc = 0;

				raw_max_count_kernel<<<grid, 256, 0, st>>>(d_raw, d_raw_off,
				    d_doc_off, N, d_maxc);
				e->launches++;
				if (cudaMemcpyAsync(&maxc, d_maxc, 4, cudaMemcpyDeviceToHost, st) ||
				    cudaStreamSynchronize(st)) {
					fail(e, "raw block scan failed: %s",
					    cudaGetErrorString(cudaGetLastError()));
					break;
				}
				e->wide = maxc > 0xffffu;
			}
			if (raw && e->wide)
				expand_raw_kernel<true><<<grid, 256, 0, st>>>(d_raw, d_raw_off,
				    d_doc_off, e->d_doc_len, N, V, d_keys, d_vals_alt);
			else if (raw)
				expand_raw_kernel<false><<<grid, 256, 0, st>>>(d_raw, d_raw_off,
				    d_doc_off, e->d_doc_len, N, V, d_keys, d_vals_alt);
			else if (e->wide)
				expand_pairs_kernel<true><<<grid, 256, 0, st>>>(d_pairs,
				    d_doc_off, e->d_doc_len, N, V, d_keys, d_vals_alt);
			else
				expand_pairs_kernel<false><<<grid, 256, 0, st>>>(d_pairs,
				    d_doc_off, e->d_doc_len, N, V, d_keys, d_vals_alt);
			e->launches++;
		}

		/* Stable sort by term; document order inside a term survives. */
		int end_bit = 1;
		while ((1ull << end_bit) <= V)
			end_bit++;
		cub::DoubleBuffer<uint32_t> kb(d_keys, d_keys_alt);
		cub::DoubleBuffer<unsigned long long> vb(
		    reinterpret_cast<unsigned long long *>(d_vals_alt),
		    reinterpret_cast<unsigned long long *>(e->d_post));
		size_t tmp = 0;
		if (P) {
			if (cub::DeviceRadixSort::SortPairs(nullptr, tmp, kb, vb,
			    (unsigned long long)P, 0, end_bit, st) != cudaSuccess) {
				fail(e, "cub sort sizing failed");
				break;
			}
			if (tmp > e->cub_tmp_bytes) {
				if (e->d_cub_tmp)
					cudaFree(e->d_cub_tmp);
				e->d_cub_tmp = nullptr;
				if (cudaMalloc(&e->d_cub_tmp, tmp) != cudaSuccess) {
					fail(e, "cub temp allocation (%zu bytes) failed", tmp);
					break;
				}
				e->cub_tmp_bytes = tmp;
			}
			if (cub::DeviceRadixSort::SortPairs(e->d_cub_tmp, tmp, kb, vb,
			    (unsigned long long)P, 0, end_bit, st) != cudaSuccess) {
				fail(e, "cub sort failed: %s",
				    cudaGetErrorString(cudaGetLastError()));
				break;
			}
			e->launches += 8;
			if (vb.Current() != reinterpret_cast<unsigned long long *>(e->d_post)) {
				/* Result landed in the scratch buffer: swap roles. */
				std::swap(e->d_post, d_vals_alt);
			}
		}
		term_offsets_kernel<<<e->n_sms * 8, 256, 0, st>>>(kb.Current(), P, V,
		    e->d_term_off);
		local_df_kernel<<<(V + 255) / 256, 256, 0, st>>>(e->d_term_off, V,
		    e->d_df_local);
		e->launches += 2;

		e->h_df_local.resize(V);
		if (cudaMemcpyAsync(e->h_df_local.data(), e->d_df_local, (size_t)V * 4,
		    cudaMemcpyDeviceToHost, st) || cudaStreamSynchronize(st)) {
			fail(e, "image build failed: %s",
			    cudaGetErrorString(cudaGetLastError()));
			break;
		}

		/* Permanent skip rows for the long lists. */
		std::vector<int32_t> row(V, -1);
		std::vector<uint32_t> longs;
		for (uint32_t t = 0; t < V; t++) {
			if (e->h_df_local[t] >= DF_LONG) {
				row[t] = (int32_t)longs.size();
				longs.push_back(t);
			}
		}
		e->n_long = longs.size();
		e->n_mt = std::max(1u, (N + (1u << MT_SHIFT) - 1) >> MT_SHIFT);
		if (dev_alloc(&e->d_skip, (size_t)e->n_long * (e->ntiles + 1)) ||
		    dev_alloc(&e->d_skip_mt, (size_t)e->n_long * (e->n_mt + 1)) ||
		    dev_alloc(&d_long, e->n_long)) {
			fail(e, "skip table allocation failed");
			break;
		}
		cudaMemcpyAsync(e->d_skip_row, row.data(), (size_t)V * 4,
		    cudaMemcpyHostToDevice, st);
		if (e->n_long) {
			cudaMemcpyAsync(d_long, longs.data(), (size_t)e->n_long * 4,
			    cudaMemcpyHostToDevice, st);
			if (dev_alloc(&e->d_long_terms, e->n_long) != cudaSuccess) {
				fail(e, "skip table allocation failed");
				break;
			}
			cudaMemcpyAsync(e->d_long_terms, longs.data(), (size_t)e->n_long * 4,
			    cudaMemcpyHostToDevice, st);
			build_skip_rows_kernel<<<e->n_long, 256, 0, st>>>(e->d_post,
			    e->d_term_off, d_long, e->ntiles, e->d_skip, TILE_SHIFT);
			build_skip_rows_kernel<<<e->n_long, 256, 0, st>>>(e->d_post,
			    e->d_term_off, d_long, e->n_mt, e->d_skip_mt, MT_SHIFT);
			e->launches += 2;
		}
		if (cudaStreamSynchronize(st) != cudaSuccess) {
			fail(e, "skip table build failed: %s",
			    cudaGetErrorString(cudaGetLastError()));
			break;
		}

		/* Dense columns for the lists present in most documents. */
		{
			std::vector<int32_t> col(V, -1);
			std::vector<uint32_t> dterms;
			const unsigned long long col_words =
			    (unsigned long long)e->ntiles * TILE_DOCS;

			if (!e->wide && N >= TILE_DOCS && e->dense_min <= 1.f) {
				for (uint32_t t = 0; t < V; t++) {
					if ((double)e->h_df_local[t] >= (double)e->dense_min * N &&
					    dterms.size() < 256) {
						col[t] = (int32_t)dterms.size();
						dterms.push_back(t);
					}
				}
			}
			e->n_dense = dterms.size();
			uint32_t *d_dterms = nullptr;
			bool ok = dev_alloc(&e->d_dense_col, V) == cudaSuccess &&
			    dev_alloc(&e->d_dense, (size_t)e->n_dense * col_words) == cudaSuccess &&
			    dev_alloc(&e->d_dense_sc, (size_t)e->n_dense * col_words) == cudaSuccess &&
			    dev_alloc(&e->d_dense_used, 512) == cudaSuccess &&	/* used[256] | max score bits[256] */
			    dev_alloc(&d_dterms, e->n_dense) == cudaSuccess;
			if (ok) {
				cudaMemcpyAsync(e->d_dense_col, col.data(), (size_t)V * 4,
				    cudaMemcpyHostToDevice, st);
				e->h_dense_col = col;
				if (e->n_dense) {
					cudaMemsetAsync(e->d_dense, 0, (size_t)e->n_dense * col_words * 4, st);
					cudaMemcpyAsync(d_dterms, dterms.data(), (size_t)e->n_dense * 4,
					    cudaMemcpyHostToDevice, st);
					build_dense_columns_kernel<<<dim3(e->n_sms * 2, e->n_dense), 256, 0, st>>>(
					    e->d_post, e->d_term_off, d_dterms, col_words, e->d_dense);
					e->launches++;
				}
				ok = cudaStreamSynchronize(st) == cudaSuccess;
			}
			e->d_dense_terms = d_dterms;
			if (!ok) {
				fail(e, "dense column build failed: %s",
				    cudaGetErrorString(cudaGetLastError()));
				break;
			}
		}

		/* Block arrays of the column terms (bmw.cuh). */
		{
			e->nblocks = std::max(1u, (N + (1u << e->bshift) - 1) >> e->bshift);
			e->row_stride = (e->nblocks + 31u) & ~31u;
			e->nchunks = (e->nblocks + BMW_CH_BLOCKS - 1) / BMW_CH_BLOCKS;
			e->sb_stride = e->nchunks * BMW_CH_SB;
			/* Whole superblocks of mini-tiles per row, zero padded. */
			e->mt_stride = e->sb_stride << (e->bshift + 5 - MT_SHIFT);
			std::vector<uint32_t> bcol(V, BMW_BCOL_NONE), bterms;
			const size_t stride = e->row_stride;

			if (!e->wide && e->bmw_enabled) {
				/* A posting per two blocks on average at least; 2 GiB of arrays at most. */
				const uint64_t min_df = std::max<uint64_t>(e->nblocks / 2, 64);
				const size_t cap = std::max<size_t>(1, (2ull << 30) /
				    (4ull * (e->nblocks + 1) + 8ull * stride));

				for (uint32_t t = 0; t < V; t++)
					if (e->h_df_local[t] >= min_df)
						bterms.push_back(t);
				if (bterms.size() > cap) {
					std::nth_element(bterms.begin(), bterms.begin() + cap, bterms.end(),
					    [&](uint32_t a, uint32_t b) {
						return e->h_df_local[a] > e->h_df_local[b];
					    });
					bterms.resize(cap);
					std::sort(bterms.begin(), bterms.end());
				}
				for (size_t c = 0; c < bterms.size(); c++)
					bcol[bterms[c]] = (uint32_t)c;
			}
			e->n_bcol = bterms.size();
			bool ok = dev_alloc(&e->d_bcol, V) == cudaSuccess &&
			    dev_alloc(&e->d_bcol_terms, e->n_bcol) == cudaSuccess &&
			    dev_alloc(&e->d_boff, (size_t)e->n_bcol * (e->nblocks + 1)) == cudaSuccess &&
			    dev_alloc(&e->d_bmax_bm25, (size_t)e->n_bcol * stride) == cudaSuccess &&
			    dev_alloc(&e->d_bmax_tfidf, (size_t)e->n_bcol * stride) == cudaSuccess &&
			    dev_alloc(&e->d_smax_bm25, (size_t)e->n_bcol * e->sb_stride) == cudaSuccess &&
			    dev_alloc(&e->d_smax_tfidf, (size_t)e->n_bcol * e->sb_stride) == cudaSuccess &&
			    (e->wide || (dev_alloc(&e->d_wmax_bm25, V) == cudaSuccess &&
			    dev_alloc(&e->d_wmax_tfidf, V) == cudaSuccess));
			if (ok && !e->wide && e->bmw_enabled && e->n_long)
				ok = dev_alloc(&e->d_mt_bm25, (size_t)e->n_long * e->mt_stride) == cudaSuccess &&
				    dev_alloc(&e->d_mt_tfidf, (size_t)e->n_long * e->mt_stride) == cudaSuccess &&
				    dev_alloc(&e->d_mt_bits, (size_t)e->n_long * e->mt_stride) == cudaSuccess;
			/* Threshold priming (bmw.cuh): the ladder tables and, for the
			 * lists longer than one part, the units of the two-level pass. */
			std::vector<uint2> units, longs;
			if (ok && !e->wide && e->bmw_enabled && e->prime_enabled) {
				for (uint32_t t = 0; t < V; t++) {
					const uint32_t df = e->h_df_local[t];

					if (df <= BMW_KTH_PART)
						continue;
					longs.push_back(make_uint2(t, (uint32_t)units.size()));
					for (uint32_t part = 0; part * BMW_KTH_PART < df; part++)
						units.push_back(make_uint2(t, part));
				}
				e->n_kth_units = units.size();
				e->n_kth_longs = longs.size();
				ok = dev_alloc(&e->d_kth_bm25, (size_t)V * BMW_LADDER) == cudaSuccess &&
				    dev_alloc(&e->d_kth_tfidf, (size_t)V * BMW_LADDER) == cudaSuccess &&
				    dev_alloc(&e->d_kth_units, units.size()) == cudaSuccess &&
				    dev_alloc(&e->d_kth_longs, longs.size()) == cudaSuccess &&
				    dev_alloc(&e->d_kth_scratch, units.size() * 2 * BMW_KTH_KEEP) == cudaSuccess;
				if (ok && !units.empty()) {
					cudaMemcpyAsync(e->d_kth_units, units.data(), units.size() * sizeof(uint2),
					    cudaMemcpyHostToDevice, st);
					cudaMemcpyAsync(e->d_kth_longs, longs.data(), longs.size() * sizeof(uint2),
					    cudaMemcpyHostToDevice, st);
				}
			}
			if (ok) {
				cudaMemcpyAsync(e->d_bcol, bcol.data(), (size_t)V * 4,
				    cudaMemcpyHostToDevice, st);
				if (e->n_bcol) {
					cudaMemcpyAsync(e->d_bcol_terms, bterms.data(), (size_t)e->n_bcol * 4,
					    cudaMemcpyHostToDevice, st);
					build_block_offsets_kernel<<<e->n_bcol, 256, 0, st>>>(e->d_post,
					    e->d_term_off, e->d_bcol_terms, e->nblocks, e->bshift, e->d_boff);
					e->launches++;
				}
				ok = cudaStreamSynchronize(st) == cudaSuccess;
			}
			if (!ok) {
				fail(e, "block array build failed: %s",
				    cudaGetErrorString(cudaGetLastError()));
				break;
			}
		}

		if (sd->df)
			e->h_df.assign(sd->df, sd->df + V);
		else
			e->h_df = e->h_df_local;
		if (upload_stats(e) == -1)
			break;
		e->loaded = true;
		rc = 0;
	} while (0);

	dev_free(d_pairs);
	dev_free(d_raw);
	dev_free(d_raw_off);
	dev_free(d_maxc);
	dev_free(d_doc_off);
	dev_free(d_keys);
	dev_free(d_keys_alt);
	dev_free(d_vals_alt);
	dev_free(d_long);
	if (rc != 0)
		free_image(e);
	return rc;
}

extern "C" int
nxsb_engine_get_df(nxsb_engine_t *e, uint32_t *df, uint32_t n_terms)
{
	if (is_multi(e) && e->sharded) {
		/* Whole-index df = the sum over the shards. */
		std::vector<uint32_t> one(n_terms);

		memset(df, 0, (size_t)n_terms * 4);
		for (size_t r = 0; r < e->replicas.size(); r++) {
			if (nxsb_engine_get_df(e->replicas[r], one.data(), n_terms) != 0)
				return multi_fail(e, e->replicas[r], (int)r);
			for (uint32_t t = 0; t < n_terms; t++)
				df[t] += one[t];
		}
		return 0;
	}
	if (is_multi(e))
		return nxsb_engine_get_df(e->replicas[0], df, n_terms) == 0 ? 0 : multi_fail(e, e->replicas[0], 0);
	if (!e->loaded || n_terms != e->n_terms)
		return fail(e, "get_df: no image or vocabulary size mismatch");
	memcpy(df, e->h_df_local.data(), (size_t)n_terms * 4);
	return 0;
}

extern "C" int
nxsb_engine_set_global_stats(nxsb_engine_t *e, const uint32_t *df,
    uint32_t n_terms, uint64_t token_count, uint32_t doc_count)
{
	if (is_multi(e))
		return multi_each(e, true, [&](nxsb_engine_t *c, int) {
			return nxsb_engine_set_global_stats(c, df, n_terms, token_count, doc_count);
		});
	/* The image is about to change: both lanes drain first. */
	sync_lanes(e);
	if (use_lane(e, 0) == -1)
		return -1;
	/*
	 * The vocabulary only grows: a segment built when it had fewer terms
	 * takes the leading part of a longer table.
	 */
	if (!e->loaded || n_terms < e->n_terms)
		return fail(e, "set_global_stats: no image or vocabulary size mismatch");
	CK(e, cudaSetDevice(e->device));
	e->h_df.assign(df, df + e->n_terms);
	e->token_count = token_count;
	e->doc_count = doc_count;
	if (upload_stats(e) == -1)
		return -1;
	for (nxsb_engine *c : e->segs)
		if (nxsb_engine_set_global_stats(c, df, n_terms, token_count, doc_count) == -1)
			return fail(e, "%s", c->err);
	return 0;
}

/*
 * Incremental refresh: delta segments and removal notes.
 */

extern "C" int
nxsb_engine_segment_add(nxsb_engine_t *e, const nxsb_shard_desc_t *sd)
{
	if (is_multi(e) && e->sharded)
		return shard_segment_add(e, sd);
	if (is_multi(e)) {
		if (multi_each(e, true, [&](nxsb_engine_t *c, int) {
			return nxsb_engine_segment_add(c, sd) < 0 ? -1 : 0;
		}) != 0)
			return -1;
		return nxsb_engine_segment_count(e->replicas[0]);
	}
	/* The image is about to change: both lanes drain first. */
	sync_lanes(e);
	if (use_lane(e, 0) == -1)
		return -1;
	if (!e->loaded)
		return fail(e, "segment_add: no base image loaded");
	if (e->segs.size() >= NXSB_MAX_SEGMENTS)
		return fail(e, "segment_add: %d delta segments already", NXSB_MAX_SEGMENTS);
	if (!sd->df)
		return fail(e, "segment_add: whole-index df[] is required");
	CK(e, cudaSetDevice(e->device));
	CK(e, cudaStreamSynchronize(e->stream));

	nxsb_engine *c = nxsb_engine_create(e->device);
	if (!c)
		return fail(e, "segment_add: %s", g_last_error);
	c->stream = e->stream;
	c->force_v2 = e->force_v2;
	c->bmw_enabled = e->bmw_enabled;
	c->bshift = e->bshift;
	if (nxsb_engine_load_shard(c, sd) == -1) {
		fail(e, "segment_add: %s", c->err);
		c->stream = c->own_stream;
		nxsb_engine_destroy(c);
		return -1;
	}
	e->segs.push_back(c);
	return (int)e->segs.size();
}

extern "C" int
nxsb_engine_segment_count(const nxsb_engine_t *e)
{
	if (is_multi(e) && e->sharded)
		return (int)e->seg_at.size();
	if (is_multi(e))
		return nxsb_engine_segment_count(e->replicas[0]);
	return (int)e->segs.size();
}

extern "C" int
nxsb_engine_segments_drop(nxsb_engine_t *e)
{
	if (is_multi(e)) {
		e->seg_at.clear();
		return multi_each(e, false, [](nxsb_engine_t *c, int) { return nxsb_engine_segments_drop(c); });
	}
	/* The image is about to change: both lanes drain first. */
	sync_lanes(e);
	if (use_lane(e, 0) == -1)
		return -1;
	CK(e, cudaSetDevice(e->device));
	CK(e, cudaStreamSynchronize(e->stream));
	drop_segments(e);
	return 0;
}

extern "C" int
nxsb_engine_set_dead(nxsb_engine_t *e, uint32_t segment, const uint64_t *ids,
    uint32_t n)
{
	if (is_multi(e) && e->sharded)
		return shard_set_dead(e, segment, ids, n);
	if (is_multi(e))
		return multi_each(e, false, [&](nxsb_engine_t *c, int) { return nxsb_engine_set_dead(c, segment, ids, n); });
	/* The image is about to change: both lanes drain first. */
	sync_lanes(e);
	if (use_lane(e, 0) == -1)
		return -1;
	if (!e->loaded || segment > e->segs.size())
		return fail(e, "set_dead: no such segment %u", segment);
	for (uint32_t i = 1; i < n; i++)
		if (ids[i - 1] >= ids[i])
			return fail(e, "set_dead: ids must be strictly ascending");
	CK(e, cudaSetDevice(e->device));
	CK(e, cudaStreamSynchronize(e->stream));
	e->dead[segment].assign(ids, ids + n);

	std::vector<unsigned long long> all;
	uint32_t off[NXSB_MAX_SEGMENTS + 2] = { 0 };

	e->max_dead = 0;
	for (uint32_t g = 0; g <= NXSB_MAX_SEGMENTS; g++) {
		off[g] = all.size();
		all.insert(all.end(), e->dead[g].begin(), e->dead[g].end());
		e->max_dead = std::max<uint32_t>(e->max_dead, e->dead[g].size());
	}
	off[NXSB_MAX_SEGMENTS + 1] = all.size();
	e->n_dead = all.size();
	dev_free(e->d_dead);
	if (!e->d_dead_off)
		CK(e, dev_alloc(&e->d_dead_off, NXSB_MAX_SEGMENTS + 2));
	CK(e, dev_alloc(&e->d_dead, all.size()));
	CK(e, cudaMemcpy(e->d_dead, all.data(), all.size() * 8, cudaMemcpyHostToDevice));
	CK(e, cudaMemcpy(e->d_dead_off, off, sizeof(off), cudaMemcpyHostToDevice));
	return 0;
}

/*
 * Batches.
 */

/* The query's postfix program for a document held by exactly the token slots in m. */
static bool
eval_program(const nxsb_batch_t *b, const nxsb_query_t &q, uint32_t m)
{
	bool st[NXSB_MAX_QUERY_PROG / 2 + 2];
	int sp = 0;

	for (uint32_t c = 0; c < q.n_prog; c++) {
		const int32_t op = b->prog[q.prog_off + c];

		if (op >= 0) {
			st[sp++] = (m >> op) & 1u;
		} else if (op == NXSB_OP_EMPTY) {
			st[sp++] = false;
		} else {
			const bool y = st[--sp];
			const bool x = st[sp - 1];

			st[sp - 1] = op == NXSB_OP_AND ? (x && y) :
			    op == NXSB_OP_OR ? (x || y) : (x && !y);
		}
	}
	return sp ? st[sp - 1] : false;
}

static bool
is_pure_or(const nxsb_batch_t *b, const nxsb_query_t &q)
{
	/* Every token pushed at least once, only OR operators. */
	uint32_t seen = 0;

	if (q.n_tokens > 32)
		return false;
	for (uint32_t i = 0; i < q.n_prog; i++) {
		const int32_t op = b->prog[q.prog_off + i];

		if (op >= 0)
			seen |= 1u << op;
		else if (op != NXSB_OP_OR)
			return false;
	}
	return seen == (q.n_tokens == 32 ? 0xffffffffu : (1u << q.n_tokens) - 1);
}

static int
validate_batch(nxsb_engine_t *e, const nxsb_batch_t *b)
{
	if (!e->loaded)
		return fail(e, "no shard image loaded");
	if (b->limit == 0)
		return fail(e, "limit must be >= 1");
	if (b->algo != NXSB_ALGO_BM25 && b->algo != NXSB_ALGO_TFIDF)
		return fail(e, "unknown ranking algorithm %d", b->algo);
	for (uint32_t i = 0; i < b->n_queries; i++) {
		const nxsb_query_t &q = b->queries[i];
		int depth = 0, maxdepth = 0;

		if (q.n_tokens > NXSB_MAX_QUERY_TOKENS || q.n_prog > NXSB_MAX_QUERY_PROG)
			return fail(e, "query %u exceeds the engine limits "
			    "(%u tokens / %u operators)", i, NXSB_MAX_QUERY_TOKENS,
			    NXSB_MAX_QUERY_PROG);
		if ((uint64_t)q.tok_off + q.n_tokens > b->n_tokens ||
		    (uint64_t)q.prog_off + q.n_prog > b->n_prog)
			return fail(e, "query %u: descriptor out of range", i);
		for (uint32_t c = 0; c < q.n_prog; c++) {
			const int32_t op = b->prog[q.prog_off + c];

			if (op >= 0) {
				if ((uint32_t)op >= q.n_tokens)
					return fail(e, "query %u: bad token slot", i);
				depth++;
			} else if (op == NXSB_OP_EMPTY) {
				depth++;
			} else if (op >= NXSB_OP_ANDNOT) {
				if (depth < 2)
					return fail(e, "query %u: malformed program", i);
				depth--;
			} else {
				return fail(e, "query %u: unknown operator", i);
			}
			maxdepth = std::max(maxdepth, depth);
		}
		if (q.n_prog && depth != 1)
			return fail(e, "query %u: malformed program", i);
		if (maxdepth > NXSB_MAX_QUERY_TOKENS + 1)
			return fail(e, "query %u: expression too deep", i);
	}
	return 0;
}

/*
 * Stage a batch on the device: classify the queries, size the (grow-only)
 * buffers, pack all descriptors into one pinned block and copy it with a
 * single cudaMemcpyAsync.
 */
/*
 * How many token slots of a boolean query can be present in a matching
 * document (a slot only ever under NOT cannot): polarity through the postfix
 * program, AND NOT swapping its right operand's.  A routing estimate only.
 */
static uint32_t
positive_tokens(const nxsb_batch_t *b, const nxsb_query_t &q)
{
	uint32_t pos[NXSB_MAX_QUERY_PROG / 2 + 2], neg[NXSB_MAX_QUERY_PROG / 2 + 2];
	int sp = 0;

	for (uint32_t c = 0; c < q.n_prog; c++) {
		const int32_t op = b->prog[q.prog_off + c];

		if (op >= 0) {
			pos[sp] = 1u << (op & 31);
			neg[sp++] = 0;
		} else if (op == NXSB_OP_EMPTY) {
			pos[sp] = neg[sp] = 0;
			sp++;
		} else if (sp >= 2) {
			sp--;
			if (op == NXSB_OP_ANDNOT) {
				pos[sp - 1] |= neg[sp];
				neg[sp - 1] |= pos[sp];
			} else {
				pos[sp - 1] |= pos[sp];
				neg[sp - 1] |= neg[sp];
			}
		}
	}
	return sp ? (uint32_t)__builtin_popcount(pos[sp - 1]) : 0;
}

static int
fill_batch(nxsb_engine_t *e, Batch &B, const nxsb_batch_t *b)
{
	if (validate_batch(e, b) == -1)
		return -1;
	B.used = true;
	B.algo = b->algo;
	B.limit = b->limit;
	B.n_q = b->n_queries;
	B.n_tok = b->n_tokens;
	B.n_prog = b->n_prog;
	B.max_tokens = 1;
	B.max_tokens_all = 1;
	B.bytes = 0;
	B.q_or.clear();
	B.q_logic.clear();
	/* OR queries with a small limit are pruned (bmw.cuh); the rest stream. */
	B.bmw = e->bmw_enabled && !e->wide && !e->force_v2 && b->limit <= BMW_K_MAX &&
	    e->d_wmax_bm25 != nullptr;

	for (uint32_t i = 0; i < b->n_queries; i++) {
		const nxsb_query_t &q = b->queries[i];

		/* search.c:224-226: nothing resolved => empty result. */
		if (q.n_tokens == 0 || q.n_prog == 0)
			continue;
		B.max_tokens_all = std::max(B.max_tokens_all, q.n_tokens);
		if (is_pure_or(b, q)) {
			B.q_or.push_back(i);
		} else {
			B.q_logic.push_back(i);
			B.max_tokens = std::max(B.max_tokens, q.n_tokens);
		}
		for (uint32_t j = 0; j < q.n_tokens; j++) {
			const uint32_t id = b->tokens[q.tok_off + j];
			if (id >= 1 && id <= e->n_terms)
				B.bytes += 8ull * e->h_df[id - 1];
		}
	}
	static_assert(sizeof(QDesc) == sizeof(nxsb_query_t), "descriptor layout");

	/*
	 * Boolean queries: pruning pays while few tokens can add up in one
	 * document (the block bound is the sum of their maxima and loosens with
	 * every one: measured, 10M documents, top-100 -- a AND b 4.8 ms per
	 * batch pruned against 9.1 streamed, (a OR b) AND c 8.0 against 13.1,
	 * a AND NOT b 2.4 against 12.5, four positive terms 30 against 27).
	 * Those go first in the list; the rest stream.
	 */
	B.n_logic_pruned = 0;
	if (B.bmw && e->bshift == BMW_LOGIC_SHIFT) {
		auto mid = std::stable_partition(B.q_logic.begin(), B.q_logic.end(), [&](uint32_t i) {
			const nxsb_query_t &q = b->queries[i];

			return q.n_tokens <= BMW_LOGIC_TOKENS && positive_tokens(b, q) <= e->logic_pruned_max_pos;
		});
		B.n_logic_pruned = (uint32_t)(mid - B.q_logic.begin());
	}

	/*
	 * Shared dense prefixes.  The dense terms of an OR query contribute
	 * the same column values to whatever query names them, so when they
	 * LEAD the token list their sum is computed once per distinct ordered
	 * prefix of the batch (a "virtual" query of just those terms) instead
	 * of once per query: the query itself then scores only the documents
	 * its other terms name, starting each from the prefix sum (a gather
	 * from the score columns), and takes the rest of its top-k from the
	 * prefix's own top-k.  Float sums keep the token-list order: a single
	 * non-dense token ahead of a single dense one also qualifies, because
	 * the first addition of a sum is commutative.
	 */
	std::vector<QDesc> vq;			// virtual queries
	std::vector<uint32_t> vtok;		// their tokens
	std::vector<uint2> qbase(B.q_or.size(), make_uint2(0u, 0u));
	if (e->share_dense && e->n_dense && !e->wide && !e->force_v2 && !B.bmw &&
	    B.limit <= ST_K_MAX && !e->h_dense_col.empty()) {
		std::vector<std::pair<uint64_t, uint32_t>> seen;	// (terms packed, prefix number)
		const uint32_t vcap = virtual_cap(B.n_q);	/* the arenas' headroom */

		for (size_t i = 0; i < B.q_or.size(); i++) {
			const nxsb_query_t &q = b->queries[B.q_or[i]];
			uint32_t terms[4], nd = 0, first_sparse = q.n_tokens, last_dense = 0;
			bool ok = true;

			for (uint32_t j = 0; j < q.n_tokens; j++) {
				const uint32_t id = b->tokens[q.tok_off + j];
				const bool dense = id >= 1 && id <= e->n_terms &&
				    e->h_dense_col[id - 1] >= 0;

				if (dense) {
					if (nd == 4) {
						ok = false;
						break;
					}
					/* A term listed twice would be added twice: leave it. */
					for (uint32_t x = 0; x < nd; x++)
						ok &= terms[x] != id;
					terms[nd++] = id;
					last_dense = j;
				} else if (first_sparse == q.n_tokens) {
					first_sparse = j;
				}
			}
			if (!ok || nd == 0)
				continue;
			/*
			 * All dense tokens lead, or the list is [sparse, dense,
			 * sparse...] (the first addition commutes): the sum starts
			 * from them.  All dense tokens LAST: they are added to the
			 * finished sums in the epilogue (0x80).  Anything else streams.
			 */
			uint32_t trailing = 0;
			if (!(last_dense < first_sparse ||
			    (nd == 1 && last_dense == 1 && first_sparse == 0))) {
				uint32_t first_dense = q.n_tokens, last_sparse = 0;

				for (uint32_t j = 0; j < q.n_tokens; j++) {
					const uint32_t id = b->tokens[q.tok_off + j];

					if (id >= 1 && id <= e->n_terms && e->h_dense_col[id - 1] >= 0)
						first_dense = std::min(first_dense, j);
					else
						last_sparse = j;
				}
				if (first_dense < last_sparse)
					continue;
				trailing = 0x80u;
			}
			uint64_t key = nd;
			uint32_t cols = 0;
			for (uint32_t x = 0; x < nd; x++) {
				key = key * 0x100000001b3ull + terms[x];
				cols |= (uint32_t)e->h_dense_col[terms[x] - 1] << (8 * x);
			}
			uint32_t pn = UINT32_MAX;
			for (auto &sp : seen)
				if (sp.first == key) {
					const QDesc &v = vq[sp.second];
					bool same = v.n_tokens == nd;
					for (uint32_t x = 0; same && x < nd; x++)
						same = vtok[v.tok_off - B.n_tok + x] == terms[x];
					if (same) {
						pn = sp.second;
						break;
					}
				}
			if (pn == UINT32_MAX && vq.size() >= vcap)
				continue;	/* out of virtual slots: the query streams its columns */
			if (pn == UINT32_MAX) {
				QDesc v;

				pn = vq.size();
				v.tok_off = B.n_tok + vtok.size();
				v.n_tokens = nd;
				v.prog_off = 0;
				v.n_prog = 0;
				vq.push_back(v);
				vtok.insert(vtok.end(), terms, terms + nd);
				seen.emplace_back(key, pn);
			}
			qbase[i] = make_uint2(cols, nd | trailing | (pn << 8));
		}
	}
	/*
	 * Boolean queries: the same base columns when the dense terms lead AND no
	 * document can satisfy the query through dense terms alone -- then only
	 * the documents the other terms name can match, the dense terms' scores
	 * and membership bits are gathered for those, and nothing else is needed
	 * (no virtual query).
	 */
	B.q_base_logic.assign(B.q_logic.size(), make_uint2(0u, 0u));
	if (e->share_dense && e->n_dense && !e->wide && !e->force_v2 &&
	    B.limit <= ST_K_MAX && B.max_tokens <= ST_LOGIC_TOKENS && !e->h_dense_col.empty()) {
		for (size_t i = 0; i < B.q_logic.size(); i++) {
			const nxsb_query_t &q = b->queries[B.q_logic[i]];
			uint32_t terms[4], slots[4], nd = 0, first_sparse = q.n_tokens, last_dense = 0;
			uint32_t dmask = 0;
			bool ok = true;

			for (uint32_t j = 0; j < q.n_tokens && ok; j++) {
				const uint32_t id = b->tokens[q.tok_off + j];
				const bool dense = id >= 1 && id <= e->n_terms &&
				    e->h_dense_col[id - 1] >= 0;

				if (dense) {
					if (nd == 4) {
						ok = false;
						break;
					}
					for (uint32_t x = 0; x < nd; x++)
						ok &= terms[x] != id;
					terms[nd] = id;
					slots[nd++] = j;
					dmask |= 1u << j;
					last_dense = j;
				} else if (first_sparse == q.n_tokens) {
					first_sparse = j;
				}
			}
			if (!ok || nd == 0 || first_sparse == q.n_tokens)
				continue;
			if (!(last_dense < first_sparse ||
			    (nd == 1 && last_dense == 1 && first_sparse == 0)))
				continue;
			/* Every subset of the dense terms, the empty one included. */
			for (uint32_t m = dmask;; m = (m - 1) & dmask) {
				if (eval_program(b, q, m)) {
					ok = false;
					break;
				}
				if (m == 0)
					break;
			}
			if (!ok)
				continue;
			uint32_t cols = 0, y = nd;
			for (uint32_t x = 0; x < nd; x++) {
				cols |= (uint32_t)e->h_dense_col[terms[x] - 1] << (8 * x);
				y |= slots[x] << (8 + 3 * x);
			}
			B.q_base_logic[i] = make_uint2(cols, y);
		}
	}
	B.n_virtual = vq.size();
	B.n_tok_all = B.n_tok + vtok.size();
	/* q_or = [virtual queries | the batch's OR queries]; q_base likewise. */
	B.q_base.assign(B.n_virtual, make_uint2(0u, 0u));
	B.q_base.insert(B.q_base.end(), qbase.begin(), qbase.end());
	{
		std::vector<uint32_t> lst(B.n_virtual);

		for (uint32_t v = 0; v < B.n_virtual; v++)
			lst[v] = B.n_q + v;
		B.q_or.insert(B.q_or.begin(), lst.begin(), lst.end());
	}

	const size_t nq = std::max(B.n_q, 1u);
	const size_t nslots = nq + B.n_virtual;
	const size_t o_q = 0;
	const size_t o_tok = o_q + align16(nslots * sizeof(QDesc));
	const size_t o_prog = o_tok + align16((size_t)B.n_tok_all * 4);
	const size_t o_or = o_prog + align16((size_t)B.n_prog * 4);
	const size_t o_base = o_or + align16(B.q_or.size() * 4);
	const size_t o_lg = o_base + align16(B.q_or.size() * sizeof(uint2));
	const size_t o_lgb = o_lg + align16(B.q_logic.size() * 4);
	const size_t desc_bytes = o_lgb + align16(B.q_logic.size() * sizeof(uint2));

	const size_t s_toks = 0;
	const size_t s_skip = s_toks + align16((size_t)B.n_tok_all * sizeof(DTok));
	const size_t s_pkey = s_skip + align16((size_t)B.n_tok_all * (e->ntiles + 1) * 4);
	const size_t s_pcnt = s_pkey + align16((size_t)B.n_virtual * B.limit * 8);
	const size_t s_thr = s_pcnt + align16((size_t)B.n_virtual * 4);
	const size_t s_cnt = s_thr + align16(nslots * 8);
	const size_t s_work = s_cnt + align16(nslots * 4);
	const size_t scratch_bytes = s_work + 16;

	const size_t r_counts = align16(nq * B.limit * sizeof(Rec));
	B.results_bytes = r_counts + align16(nq * 4);

	if (B.desc.ensure(desc_bytes) || B.h_desc.ensure(desc_bytes) ||
	    B.scratch.ensure(scratch_bytes) || B.results.ensure(B.results_bytes) ||
	    B.h_results.ensure(B.results_bytes))
		return fail(e, "device allocation failed for a batch of %u queries "
		    "(limit %u): %s", b->n_queries, B.limit,
		    cudaGetErrorString(cudaGetLastError()));

	char *d = (char *)B.desc.p, *h = (char *)B.h_desc.p, *sc = (char *)B.scratch.p;
	B.d_queries = (QDesc *)(d + o_q);
	B.d_tokens = (uint32_t *)(d + o_tok);
	B.d_prog = (int32_t *)(d + o_prog);
	B.d_qlist_or = (uint32_t *)(d + o_or);
	B.d_qbase = (uint2 *)(d + o_base);
	B.d_qlist_logic = (uint32_t *)(d + o_lg);
	B.d_qbase_logic = (uint2 *)(d + o_lgb);
	B.d_toks = (DTok *)(sc + s_toks);
	B.d_tmp_skip = (uint32_t *)(sc + s_skip);
	B.d_prefix_keys = (unsigned long long *)(sc + s_pkey);
	B.d_prefix_cnt = (uint32_t *)(sc + s_pcnt);
	B.d_thr = (unsigned long long *)(sc + s_thr);
	B.d_cand_count = (uint32_t *)(sc + s_cnt);
	B.d_work = (uint32_t *)(sc + s_work);
	B.zero_bytes = scratch_bytes - s_thr;
	B.d_recs = (Rec *)B.results.p;
	B.d_counts = (uint32_t *)((char *)B.results.p + r_counts);

	memcpy(h + o_q, b->queries, (size_t)B.n_q * sizeof(QDesc));
	memcpy(h + o_q + (size_t)B.n_q * sizeof(QDesc), vq.data(), vq.size() * sizeof(QDesc));
	memcpy(h + o_tok, b->tokens, (size_t)B.n_tok * 4);
	memcpy(h + o_tok + (size_t)B.n_tok * 4, vtok.data(), vtok.size() * 4);
	memcpy(h + o_prog, b->prog, (size_t)B.n_prog * 4);
	memcpy(h + o_or, B.q_or.data(), B.q_or.size() * 4);
	memcpy(h + o_base, B.q_base.data(), B.q_base.size() * sizeof(uint2));
	memcpy(h + o_lg, B.q_logic.data(), B.q_logic.size() * 4);
	memcpy(h + o_lgb, B.q_base_logic.data(), B.q_base_logic.size() * sizeof(uint2));
	CK(e, cudaMemcpyAsync(d, h, desc_bytes, cudaMemcpyHostToDevice, e->stream));
	return 0;
}

extern "C" int
nxsb_engine_batch_upload(nxsb_engine_t *e, const nxsb_batch_t *b)
{
	NOT_ON_REPLICATED(e, "a resident batch");
	int h = -1;

	CK(e, cudaSetDevice(e->device));
	for (int i = 0; i < MAX_HANDLES; i++)
		if (!e->batches[i].used) {
			h = i;
			break;
		}
	if (h < 0)
		return fail(e, "too many resident batches");
	if (fill_batch(e, e->batches[h], b) == -1) {
		free_batch(e->batches[h]);
		return -1;
	}
	if (cudaStreamSynchronize(e->stream) != cudaSuccess) {
		free_batch(e->batches[h]);
		return fail(e, "batch upload failed: %s",
		    cudaGetErrorString(cudaGetLastError()));
	}
	return h;
}

extern "C" int
nxsb_engine_batch_release(nxsb_engine_t *e, int h)
{
	if (h < 0 || h >= MAX_HANDLES || !e->batches[h].used)
		return fail(e, "bad batch handle %d", h);
	cudaSetDevice(e->device);
	cudaStreamSynchronize(e->stream);
	free_batch(e->batches[h]);
	return 0;
}

extern "C" uint64_t
nxsb_engine_batch_bytes(nxsb_engine_t *e, int h)
{
	if (h < 0 || h >= MAX_HANDLES || !e->batches[h].used)
		return 0;
	return e->batches[h].bytes;
}

static void
begin_run(nxsb_engine_t *e)
{
	e->n_runs++;
	e->runs[(e->n_runs - 1) % EV_RING].n = 0;
}

static void
mark(nxsb_engine_t *e, const char *name)
{
	nxsb_engine::RunEvents &r = e->runs[(e->n_runs - 1) % EV_RING];

	if (r.n < EV_PER_RUN) {
		r.names[r.n] = name;
		cudaEventRecord(r.ev[r.n], e->stream);
		r.n++;
	}
}

template <bool LOGIC>
static int
launch_tiles(nxsb_engine_t *e, Batch &B, const uint32_t *d_qlist, uint32_t n_q,
    uint32_t k_tile, uint64_t cand_cap)
{
	ScoreParams p;
	size_t smem = (size_t)TILE_DOCS * 4;

	p.post = e->d_post;
	p.toks = B.d_toks;
	p.queries = B.d_queries;
	p.prog = B.d_prog;
	p.qlist = d_qlist;
	p.n_q = n_q;
	p.ntiles = e->ntiles;
	p.k = k_tile;
	p.thr = B.d_thr;
	p.cand_count = B.d_cand_count;
	p.cand = e->d_cand;
	p.cand_cap = cand_cap;
	p.work_counter = B.d_work;
	p.logtab = e->d_logtab;
	p.doc_len = e->d_doc_len;
	p.K0 = e->K0;
	p.K1 = e->K1;
	p.algo = B.algo;
	p.max_tokens = B.max_tokens;
	if (LOGIC)
		smem += ((size_t)B.max_tokens + 1) * TILE_WORDS * 4;

	auto kern = B.algo == NXSB_ALGO_BM25
	    ? (e->wide ? score_tiles_kernel<LOGIC, true, NXSB_ALGO_BM25>
	       : score_tiles_kernel<LOGIC, false, NXSB_ALGO_BM25>)
	    : (e->wide ? score_tiles_kernel<LOGIC, true, NXSB_ALGO_TFIDF>
	       : score_tiles_kernel<LOGIC, false, NXSB_ALGO_TFIDF>);
	int per_sm = 0;

	if (kernel_fit(e, (const void *)kern, TILE_THREADS, smem, &per_sm) == -1)
		return -1;
	if (per_sm < 1)
		return fail(e, "scoring kernel does not fit an SM (smem %zu)", smem);

	const unsigned long long items = (unsigned long long)n_q * e->ntiles;
	unsigned grid = (unsigned)std::min<unsigned long long>(items,
	    (unsigned long long)e->n_sms * per_sm);

	/* thresholds, candidate counts and the work counter: one memset. */
	CK(e, cudaMemsetAsync(B.d_thr, 0, B.zero_bytes, e->stream));
	mark(e, "score_tiles");
	kern<<<grid, TILE_THREADS, smem, e->stream>>>(p);
	e->launches++;
	CK(e, cudaGetLastError());
	return 0;
}

/*
 * The TMA-fed scorer (stream.cuh): plan the items, then one persistent
 * launch, two CTAs per SM.
 */
template <bool LOGIC>
static int
launch_stream(nxsb_engine_t *e, Batch &B, const uint32_t *d_qlist,
    const uint2 *d_qbase, uint32_t n_q, uint32_t k_tile, uint64_t cand_cap,
    size_t chunk_cap)
{
	const uint64_t items = (uint64_t)n_q * e->ntiles;
	const uint32_t stride = 16u * (1u + std::max(B.max_tokens_all, 1u));
	const size_t want_cnt = (size_t)items * 4;
	StreamParams p;

	/* Sized for the chunk capacity run_list works with, and 4-token records at least. */
	{
		const size_t cap_q = std::max<size_t>(n_q, std::min<size_t>(slots_for(B.n_q), chunk_cap));
		const size_t cap_items = cap_q * e->ntiles;

		if (ensure_arena(e, e->d_plan, e->plan_bytes,
		    cap_items * std::max(stride, 16u * 5u), "plan arena") == -1 ||
		    ensure_arena(e, e->d_tile_cnt, e->tile_cnt_bytes, cap_items * 4,
		    "tile count") == -1)
			return -1;
	}
	p.post = e->d_post;
	p.plan = e->d_plan;
	p.plan_stride = stride;
	p.n_q = n_q;
	p.ntiles = e->ntiles;
	p.k = k_tile;
	p.thr = B.d_thr;
	p.tile_count = e->d_tile_cnt;
	p.cand = e->d_cand;
	p.work_counter = B.d_work;
	p.logtab = e->d_logtab;
	p.doc_len = e->d_doc_len;
	p.K0 = e->K0;
	p.K1 = e->K1;
	(void)cand_cap;		/* = ntiles * k_tile: one k-cell per (query, tile) */
	p.prof = e->d_prof;

	if (LOGIC) {
		/* Truth tables of the boolean programs and their subset closures:
		 * 8 + 8 words per query. */
		if (ensure_arena(e, e->d_tt, e->tt_bytes,
		    std::max<size_t>(n_q, B.n_q) * 16 * 4, "truth table") == -1)
			return -1;
	}
	p.tt = e->d_tt;
	p.dense = reinterpret_cast<const uint32_t *>(e->d_dense_sc);
	p.col_words = (unsigned long long)e->ntiles * TILE_DOCS;
	p.dense_max = e->d_dense_used ? e->d_dense_used + 256 : nullptr;

	auto kern = B.algo == NXSB_ALGO_BM25
	    ? (e->wide ? score_stream_kernel<LOGIC, true, NXSB_ALGO_BM25>
	       : score_stream_kernel<LOGIC, false, NXSB_ALGO_BM25>)
	    : (e->wide ? score_stream_kernel<LOGIC, true, NXSB_ALGO_TFIDF>
	       : score_stream_kernel<LOGIC, false, NXSB_ALGO_TFIDF>);
	const size_t smem = StCfg<LOGIC>::SMEM;
	int per_sm = 0;

	if (kernel_fit(e, (const void *)kern, ST_THREADS, smem, &per_sm) == -1)
		return -1;
	if (per_sm < 1)
		return fail(e, "stream kernel does not fit an SM (smem %zu)", smem);
	const unsigned grid = (unsigned)std::min<uint64_t>(items,
	    (uint64_t)e->n_sms * per_sm);

	CK(e, cudaMemsetAsync(B.d_thr, 0, B.zero_bytes, e->stream));
	CK(e, cudaMemsetAsync(e->d_tile_cnt, 0, want_cnt, e->stream));
	mark(e, "plan");
	if (LOGIC) {
		/* tt[n_q][8] truth tables, then tu[n_q][8] their subset closures. */
		truth_tables_kernel<<<n_q, 256, 0, e->stream>>>(B.d_queries, d_qlist,
		    B.d_prog, e->d_tt, e->d_tt + (size_t)n_q * 8);
		e->launches++;
	}
	plan_items_kernel<<<(unsigned)((items + 255) / 256), 256, 0, e->stream>>>(
	    B.d_queries, d_qlist, d_qbase, LOGIC ? e->d_tt + (size_t)n_q * 8 : nullptr,
	    B.d_toks, n_q, e->ntiles, stride, e->d_plan);
	mark(e, "score_tiles");
	kern<<<grid, ST_THREADS, smem, e->stream>>>(p);
	e->launches += 2;
	CK(e, cudaGetLastError());
	return 0;
}

/*
 * OR queries through the block-max scorer (bmw.cuh): one persistent launch
 * over (query, chunk) items, then the same per-query merge of the cells.
 */
static int
run_bmw(nxsb_engine_t *e, Batch &B, const uint32_t *d_qlist, uint32_t n_list, Rec *d_recs,
    bool logic)
{
	cudaStream_t st = e->stream;
	const uint32_t k = B.limit;
	const size_t cells = (size_t)e->nchunks * k;
	/* Item numbers are 32-bit. */
	const uint32_t chunk = (uint32_t)std::min<uint64_t>(n_list, 0xfffffff0ull / (e->nchunks + 1));
	const size_t slots = std::max<size_t>(chunk, std::min<size_t>(slots_for(B.n_q), chunk));

	if (ensure_arena(e, e->d_cand, e->cand_bytes, slots * cells * 8, "candidate arena") == -1 ||
	    ensure_arena(e, e->d_tile_cnt, e->tile_cnt_bytes, slots * e->nchunks * 4,
	    "tile count") == -1)
		return -1;
	/* Boolean queries: truth tables and their subset closures, 8 + 8 words per query. */
	if (logic && ensure_arena(e, e->d_tt, e->tt_bytes,
	    std::max<size_t>(chunk, B.n_q) * 16 * 4, "truth table") == -1)
		return -1;

	for (uint32_t q0 = 0; q0 < n_list; q0 += chunk) {
		const uint32_t n = std::min(chunk, n_list - q0);
		BmwParams p;

		p.post = e->d_post;
		p.toks = B.d_toks;
		p.queries = B.d_queries;
		p.qlist = d_qlist + q0;
		p.n_q = n;
		p.nchunks = e->nchunks;
		p.nblocks = e->nblocks;
		p.row_stride = e->row_stride;
		p.n_docs = e->n_docs;
		p.ntiles = e->ntiles;
		p.k = k;
		p.boff = e->d_boff;
		p.bmax = B.algo == NXSB_ALGO_BM25 ? e->d_bmax_bm25 : e->d_bmax_tfidf;
		p.smax = B.algo == NXSB_ALGO_BM25 ? e->d_smax_bm25 : e->d_smax_tfidf;
		p.sb_stride = e->sb_stride;
		p.n_mt = e->n_mt;
		p.tt = e->d_tt;
		p.thr = B.d_thr;
		p.tile_count = e->d_tile_cnt;
		p.cand = e->d_cand;
		p.work_counter = B.d_work;
		p.logtab = e->d_logtab;
		p.K0 = e->K0;
		p.K1 = e->K1;
		p.stats = e->d_bmw_stats;

		const void *kern = nullptr;
		size_t smem = 0;
#define BMW_PICK(S) do {								\
		kern = B.algo == NXSB_ALGO_BM25						\
		    ? (const void *)score_bmw_kernel<NXSB_ALGO_BM25, S, false>		\
		    : (const void *)score_bmw_kernel<NXSB_ALGO_TFIDF, S, false>;	\
		smem = BmwCfg<S>::SMEM;							\
	} while (0)
		if (logic) {
			/* Boolean queries are served at the default block size only. */
			kern = B.algo == NXSB_ALGO_BM25
			    ? (const void *)score_bmw_kernel<NXSB_ALGO_BM25, BMW_LOGIC_SHIFT, true>
			    : (const void *)score_bmw_kernel<NXSB_ALGO_TFIDF, BMW_LOGIC_SHIFT, true>;
			smem = BmwCfg<BMW_LOGIC_SHIFT, true>::SMEM;
		} else
		switch (e->bshift) {
		case 5: BMW_PICK(5); break;
		case 6: BMW_PICK(6); break;
		case 7: BMW_PICK(7); break;
		default: BMW_PICK(8); break;
		}
#undef BMW_PICK
		int per_sm = 0;

		if (kernel_fit(e, kern, BMW_THREADS, smem, &per_sm) == -1)
			return -1;
		if (per_sm < 1)
			return fail(e, "block-max kernel does not fit an SM (smem %zu)", smem);
		const uint64_t items = (uint64_t)n * e->nchunks;
		const unsigned grid = (unsigned)std::min<uint64_t>(items, (uint64_t)e->n_sms * per_sm);

#ifdef BMW_EXP_KEEP_THR
		/* Experiment: a rerun of the batch starts from its final thresholds. */
		CK(e, cudaMemsetAsync(B.d_cand_count, 0, B.zero_bytes - ((char *)B.d_cand_count - (char *)B.d_thr), st));
#else
		CK(e, cudaMemsetAsync(B.d_thr, 0, B.zero_bytes, st));
#endif
		CK(e, cudaMemsetAsync(e->d_tile_cnt, 0, (size_t)items * 4, st));
		if (logic) {
			mark(e, "plan");
			truth_tables_kernel<<<n, 256, 0, st>>>(B.d_queries, d_qlist + q0, B.d_prog,
			    e->d_tt, e->d_tt + (size_t)n * 8);
			e->launches++;
		}
		mark(e, "score_tiles");
		void *args[] = { &p };
		CK(e, cudaLaunchKernel(kern, dim3(grid), dim3(BMW_THREADS), args, smem, st));
		e->launches++;
		mark(e, "topk");

		FinalizeShared fs;
		fs.queries = B.d_queries;
		fs.toks = B.d_toks;
		fs.post = e->d_post;
		fs.qbase = nullptr;
		fs.prefix_keys = B.d_prefix_keys;
		fs.prefix_cnt = B.d_prefix_cnt;
		fs.n_real = B.n_q;
		finalize_cells_kernel<<<n, 256, 0, st>>>(e->d_cand, e->d_tile_cnt, e->nchunks,
		    d_qlist + q0, k, e->d_doc_ids, d_recs, B.d_counts, fs);
		e->launches++;
		CK(e, cudaGetLastError());
	}
	return 0;
}

/*
 * Score one list of queries (pure-OR or boolean) in chunks bounded by the
 * candidate arena: tiles -> per-query final top-k.
 */
template <bool LOGIC>
static int
run_list(nxsb_engine_t *e, Batch &B, const uint32_t *d_qlist, uint32_t n_list,
    Rec *d_recs, bool pruned, uint32_t list_off = 0)
{
	cudaStream_t st = e->stream;
	const uint32_t k = B.limit;
	const bool small_k = k <= SMALL_K_MAX;
	const uint32_t k_tile = small_k ? k : TILE_DOCS;
	const uint64_t cand_cap = (uint64_t)e->ntiles * k_tile;
	const size_t per_q = cand_cap * 8;

	if (n_list == 0)
		return 0;

	if (pruned)
		return run_bmw(e, B, d_qlist, n_list, d_recs, LOGIC);

	const size_t slots = std::max<size_t>(n_list, slots_for(B.n_q));
	if (ensure_arena(e, e->d_cand, e->cand_bytes,
	    std::max(per_q, std::min<size_t>(CAND_ARENA_BYTES, per_q * slots)),
	    "candidate arena") == -1)
		return -1;
	uint32_t chunk = (uint32_t)std::min<size_t>(n_list, e->cand_bytes / per_q);
	/* Boolean queries fit the stream kernel when a byte holds their tokens. */
	const bool stream = k <= ST_K_MAX && !e->force_v2 &&
	    (!LOGIC || B.max_tokens <= ST_LOGIC_TOKENS);

	if (stream) {
		/* Item numbers are 32-bit; keep the plan arena under 1 GiB. */
		const uint64_t per_q_plan = (uint64_t)e->ntiles * 16u * (1u + std::max(B.max_tokens_all, 1u));
		const uint64_t cap = std::min<uint64_t>((1ull << 30) / per_q_plan,
		    0xfffffff0ull / e->ntiles);

		chunk = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(chunk, cap));
	}

	/*
	 * Shared dense prefixes need the virtual queries (the head of the
	 * list) finalized before the queries that merge with them: possible
	 * when the list goes through in one chunk, else every query streams
	 * its own dense columns as before.
	 */
	const uint32_t n_virtual = LOGIC ? 0 : B.n_virtual;
	const uint2 *d_qbase = !stream ? nullptr
	    : LOGIC ? B.d_qbase_logic + list_off	/* no virtual queries: any chunking */
	    : (n_virtual && chunk >= n_list) ? B.d_qbase : nullptr;

	for (uint32_t q0 = 0; q0 < n_list; q0 += chunk) {
		const uint32_t n = std::min(chunk, n_list - q0);

		if ((stream ? launch_stream<LOGIC>(e, B, d_qlist + q0,
		    d_qbase ? d_qbase + q0 : nullptr, n, k_tile, cand_cap, chunk)
		    : launch_tiles<LOGIC>(e, B, d_qlist + q0, n, k_tile, cand_cap)) == -1)
			return -1;
		mark(e, "topk");
		if (stream) {
			FinalizeShared fs;

			fs.queries = B.d_queries;
			fs.toks = B.d_toks;
			fs.post = e->d_post;
			fs.qbase = (!LOGIC && d_qbase) ? d_qbase + q0 : nullptr;
			fs.prefix_keys = B.d_prefix_keys;
			fs.prefix_cnt = B.d_prefix_cnt;
			fs.n_real = B.n_q;
			/* Virtual queries lead the list: [q0, q0 + nv) of this chunk. */
			const uint32_t nv = q0 < n_virtual ? std::min(n, n_virtual - q0) : 0;

			if (nv) {
				finalize_cells_kernel<<<nv, 256, 0, st>>>(e->d_cand, e->d_tile_cnt,
				    e->ntiles, d_qlist + q0, k, e->d_doc_ids, d_recs, B.d_counts, fs);
				e->launches++;
			}
			if (n > nv) {
				/* Slot numbers stay relative to the chunk: shift the views. */
				FinalizeShared fr = fs;

				if (fr.qbase)
					fr.qbase += nv;
				finalize_cells_kernel<<<n - nv, 256, 0, st>>>(
				    e->d_cand + (size_t)nv * cand_cap,
				    e->d_tile_cnt + (size_t)nv * e->ntiles, e->ntiles,
				    d_qlist + q0 + nv, k, e->d_doc_ids, d_recs, B.d_counts, fr);
				e->launches++;
			}
			CK(e, cudaGetLastError());
			continue;
		}
		if (small_k) {
			finalize_topk_kernel<<<n, 256, 0, st>>>(e->d_cand, cand_cap,
			    B.d_cand_count, d_qlist + q0, k, e->d_doc_ids, d_recs,
			    B.d_counts);
			e->launches++;
			CK(e, cudaGetLastError());
			continue;
		}

		/*
		 * Large limits: every matching document was emitted; sort each
		 * query's keys on the device (CUB) and cut at the limit.
		 */
		std::vector<uint32_t> cnt(n), qids(n);
		CK(e, cudaMemcpyAsync(cnt.data(), B.d_cand_count, (size_t)n * 4,
		    cudaMemcpyDeviceToHost, st));
		CK(e, cudaMemcpyAsync(qids.data(), d_qlist + q0, (size_t)n * 4,
		    cudaMemcpyDeviceToHost, st));
		CK(e, cudaStreamSynchronize(st));
		if (per_q > e->sort_tmp_bytes) {
			dev_free(e->d_sort_tmp);
			e->sort_tmp_bytes = 0;
			if (dev_alloc(&e->d_sort_tmp, cand_cap) != cudaSuccess)
				return fail(e, "sort buffer allocation failed");
			e->sort_tmp_bytes = per_q;
		}
		for (uint32_t i = 0; i < n; i++) {
			const unsigned long long *keys = e->d_cand + (size_t)i * cand_cap;
			size_t tmp = 0;

			if (cnt[i]) {
				cub::DeviceRadixSort::SortKeysDescending(nullptr, tmp, keys,
				    e->d_sort_tmp, (unsigned long long)cnt[i], 0, 64, st);
				if (tmp > e->cub_tmp_bytes) {
					if (e->d_cub_tmp)
						cudaFree(e->d_cub_tmp);
					e->d_cub_tmp = nullptr;
					e->cub_tmp_bytes = 0;
					if (cudaMalloc(&e->d_cub_tmp, tmp) != cudaSuccess)
						return fail(e, "cub temp allocation failed");
					e->cub_tmp_bytes = tmp;
				}
				cub::DeviceRadixSort::SortKeysDescending(e->d_cub_tmp, tmp,
				    keys, e->d_sort_tmp, (unsigned long long)cnt[i], 0, 64, st);
				e->launches += 8;
			}
			emit_sorted_kernel<<<std::max(1u, std::min(1024u, (k + 255) / 256)),
			    256, 0, st>>>(e->d_sort_tmp, cnt[i], k, e->d_doc_ids,
			    d_recs + (size_t)qids[i] * k, B.d_counts + qids[i]);
			e->launches++;
			CK(e, cudaGetLastError());
		}
	}
	return 0;
}

static int
run_batch(nxsb_engine_t *e, Batch &B, Rec *d_recs)
{
	cudaStream_t st = e->stream;
	const uint32_t n_active = B.q_or.size() + B.q_logic.size();
	/* ranking.c:86,156-166: N == 0 (or, BM25, adl < 1) => nothing scores. */
	const bool can_score = B.algo == NXSB_ALGO_BM25 ? e->stats_valid
	    : e->doc_count != 0;

	begin_run(e);
	mark(e, "resolve");

	/* Results start out empty; queries that score nothing stay so. */
	if (d_recs == B.d_recs) {
		CK(e, cudaMemsetAsync(B.results.p, 0, B.results_bytes, st));
	} else {
		CK(e, cudaMemsetAsync(d_recs, 0,
		    (size_t)std::max(B.n_q, 1u) * B.limit * sizeof(Rec), st));
		CK(e, cudaMemsetAsync(B.d_counts, 0, (size_t)std::max(B.n_q, 1u) * 4, st));
	}
	if (n_active == 0 || !can_score) {
		mark(e, "end");
		return 0;
	}

	if (e->n_dense)
		CK(e, cudaMemsetAsync(e->d_dense_used, 0, 512 * 4, st));
	resolve_tokens_kernel<<<(B.n_tok_all + 255) / 256, 256, 0, st>>>(
	    B.d_tokens, B.n_tok_all, e->n_terms, e->d_term_off, e->d_skip_row,
	    e->d_dense_col, e->d_bcol, e->d_dense_used, (unsigned long long)e->ntiles * TILE_DOCS,
	    e->d_skip, e->d_skip_mt, e->n_mt, B.d_tmp_skip,
	    B.algo == NXSB_ALGO_BM25 ? e->d_idf_bm25 : e->d_idf_tfidf,
	    B.algo == NXSB_ALGO_BM25 ? e->d_wmax_bm25 : e->d_wmax_tfidf,
	    B.bmw ? (B.algo == NXSB_ALGO_BM25 ? e->d_kth_bm25 : e->d_kth_tfidf) : nullptr,
	    bmw_ladder_step(B.limit),
	    B.bmw ? (B.algo == NXSB_ALGO_BM25 ? e->d_mt_bm25 : e->d_mt_tfidf) : nullptr,
	    B.bmw ? e->d_mt_bits : nullptr, e->mt_stride,
	    e->ntiles, B.d_toks);
	build_temp_skips_kernel<<<B.n_tok_all, 128, 0, st>>>(e->d_post, B.d_toks,
	    B.d_tmp_skip, e->ntiles);
	e->launches += 2;
	CK(e, cudaGetLastError());
	const uint32_t n_lp = B.n_logic_pruned, n_ls = (uint32_t)B.q_logic.size() - n_lp;
	if (e->n_dense && !(B.bmw && n_ls == 0)) {
		/* Score columns of the dense terms this batch refers to. */
		const unsigned long long col_words = (unsigned long long)e->ntiles * TILE_DOCS;
		const dim3 grid(e->n_sms * 4, e->n_dense);

		mark(e, "dense_scores");
		if (B.algo == NXSB_ALGO_BM25)
			dense_scores_kernel<NXSB_ALGO_BM25><<<grid, 256, 0, st>>>(e->d_dense,
			    e->d_dense_terms, e->d_dense_used, e->d_idf_bm25, e->d_logtab,
			    e->K0, e->K1, col_words, e->d_dense_sc, e->d_dense_used + 256);
		else
			dense_scores_kernel<NXSB_ALGO_TFIDF><<<grid, 256, 0, st>>>(e->d_dense,
			    e->d_dense_terms, e->d_dense_used, e->d_idf_tfidf, e->d_logtab,
			    e->K0, e->K1, col_words, e->d_dense_sc, e->d_dense_used + 256);
		e->launches++;
		CK(e, cudaGetLastError());
	}

	if (run_list<false>(e, B, B.d_qlist_or, B.q_or.size(), d_recs, B.bmw) == -1 ||
	    run_list<true>(e, B, B.d_qlist_logic, n_lp, d_recs, true) == -1 ||
	    run_list<true>(e, B, B.d_qlist_logic + n_lp, n_ls, d_recs, false, n_lp) == -1)
		return -1;
	mark(e, "end");
	return 0;
}


/*
 * Search over a segmented image: every segment scores the batch at
 * k_in = limit + max_dead (so that a segment's list still holds `limit` live
 * documents after the removed ones are dropped), the lists are merged on the
 * device.  B receives the final records (or d_out, if given) and counts;
 * slot selects the children's pooled Batch (-1: the one-shot one).
 */
static inline bool
segmented(const nxsb_engine_t *e)
{
	return !e->segs.empty() || e->n_dead != 0;
}

static int
run_segmented(nxsb_engine_t *e, Batch &B, nxsb_engine::SegRun &S,
    const nxsb_batch_t *b, Rec *d_out, int slot)
{
	const uint32_t n_segs = 1 + e->segs.size();
	const uint64_t k_in64 = (uint64_t)b->limit + e->max_dead;
	const uint32_t k_in = k_in64 > UINT32_MAX ? UINT32_MAX : (uint32_t)k_in64;
	const size_t nq = std::max(b->n_queries, 1u);
	const size_t stride = nq * k_in;
	nxsb_batch_t b2 = *b;

	/* Final buffers (and validation) at the caller's limit. */
	if (fill_batch(e, B, b) == -1)
		return -1;
	b2.limit = k_in;
	if (S.gather.ensure(stride * n_segs * sizeof(Rec)))
		return fail(e, "device allocation failed for a %u-segment search "
		    "(%u queries, limit %u)", n_segs, b->n_queries, k_in);
	Rec *gather = (Rec *)S.gather.p;

	if (fill_batch(e, S.own, &b2) == -1 || run_batch(e, S.own, gather) == -1)
		return -1;
	for (uint32_t i = 0; i < e->segs.size(); i++) {
		nxsb_engine *c = e->segs[i];
		Batch &CB = slot < 0 ? c->oneshot : c->pipe[slot];

		c->stream = e->stream;
		if (fill_batch(c, CB, &b2) == -1 ||
		    run_batch(c, CB, gather + (size_t)(i + 1) * stride) == -1)
			return fail(e, "segment %u: %s", i + 1, c->err);
	}

	cudaStream_t st = e->stream;
	Rec *out = d_out ? d_out : B.d_recs;
	const unsigned long long total = (unsigned long long)b->n_queries * n_segs * k_in;

	if (d_out) {
		CK(e, cudaMemsetAsync(d_out, 0, nq * b->limit * sizeof(Rec), st));
		CK(e, cudaMemsetAsync(B.d_counts, 0, nq * 4, st));
	} else {
		CK(e, cudaMemsetAsync(B.results.p, 0, B.results_bytes, st));
	}
	if (total == 0)
		return 0;
	if (e->n_dead) {
		const uint32_t lists = n_segs * b->n_queries;

		drop_dead_kernel<<<(lists + 3) / 4, 128, 0, st>>>(gather, n_segs,
		    b->n_queries, k_in, e->d_dead, e->d_dead_off);
		e->launches++;
	}
	merge_segments_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gather,
	    n_segs, b->n_queries, k_in, b->limit, out, B.d_counts);
	e->launches++;
	CK(e, cudaGetLastError());
	return 0;
}

extern "C" int
nxsb_engine_batch_run(nxsb_engine_t *e, int h, void *d_recs)
{
	if (h < 0 || h >= MAX_HANDLES || !e->batches[h].used)
		return fail(e, "bad batch handle %d", h);
	if (segmented(e))
		return fail(e, "resident batches cannot run on a segmented image; "
		    "use nxsb_engine_search / _search_begin");
	CK(e, cudaSetDevice(e->device));
	Batch &B = e->batches[h];
	/* A handle always runs on the same lane (its scratch is its own, not its lane's). */
	if (use_lane(e, (e->lanes_enabled && !e->external_stream) ? h % N_LANES : 0) == -1)
		return -1;
	return run_batch(e, B, d_recs ? (Rec *)d_recs : B.d_recs);
}

/* Results of a batch: one D2H copy into pinned memory, then host arrays. */
static int
enqueue_fetch(nxsb_engine_t *e, Batch &B)
{
	CK(e, cudaMemcpyAsync(B.h_results.p, B.results.p, B.results_bytes,
	    cudaMemcpyDeviceToHost, e->stream));
	return 0;
}

static void
unpack_results(const Batch &B, uint32_t *counts, uint64_t *ids, float *scores)
{
	const size_t nrec = (size_t)B.n_q * B.limit;
	const Rec *recs = (const Rec *)B.h_results.p;
	const uint32_t *cnt = (const uint32_t *)((const char *)B.h_results.p +
	    ((const char *)B.d_counts - (const char *)B.results.p));

	memcpy(counts, cnt, (size_t)B.n_q * 4);
	for (size_t i = 0; i < nrec; i++) {
		ids[i] = recs[i].doc_id;
		scores[i] = recs[i].score;
	}
}

static int
fetch_results(nxsb_engine_t *e, Batch &B, uint32_t *counts, uint64_t *ids,
    float *scores)
{
	if (enqueue_fetch(e, B) == -1)
		return -1;
	CK(e, cudaStreamSynchronize(e->stream));
	unpack_results(B, counts, ids, scores);
	return 0;
}

extern "C" int
nxsb_engine_batch_fetch(nxsb_engine_t *e, int h, uint32_t *counts,
    uint64_t *ids, float *scores)
{
	if (h < 0 || h >= MAX_HANDLES || !e->batches[h].used)
		return fail(e, "bad batch handle %d", h);
	CK(e, cudaSetDevice(e->device));
	return fetch_results(e, e->batches[h], counts, ids, scores);
}

extern "C" int
nxsb_engine_search(nxsb_engine_t *e, const nxsb_batch_t *b, uint32_t *counts,
    uint64_t *ids, float *scores)
{
	if (is_multi(e)) {
		const int h = nxsb_engine_search_begin(e, b);

		return h < 0 ? -1 : nxsb_engine_search_end(e, h, counts, ids, scores);
	}
	Batch &B = e->oneshot;

	CK(e, cudaSetDevice(e->device));
	if (use_lane(e, 0) == -1)
		return -1;
	if (segmented(e)) {
		if (run_segmented(e, B, e->seg_oneshot, b, nullptr, -1) == -1)
			return -1;
	} else if (fill_batch(e, B, b) == -1 || run_batch(e, B, B.d_recs) == -1) {
		return -1;
	}
	return fetch_results(e, B, counts, ids, scores);
}

/*
 * The same search split in two, so that a caller can prepare and submit the
 * next batch while this one is on the device: begin enqueues the descriptor
 * copy, every kernel and the result copy on the engine's stream and records
 * an event; end waits for that event only.
 */
/* Lookups whose answers the batch's token list is waiting for (search_begin_fz). */
struct FzPending {
	uint32_t		n;
	const char *		blob;
	const uint32_t *	off;	/* [n + 1] */
	const uint32_t *	pos;	/* [n] index into batch.tokens */
};

static int
search_begin(nxsb_engine_t *e, const nxsb_batch_t *b, Rec *d_recs, const FzPending *fz = nullptr)
{
	int s = -1;

	CK(e, cudaSetDevice(e->device));
	for (int i = 0; i < PIPE_DEPTH; i++)
		if (!e->pipe_busy[i]) {
			s = i;
			break;
		}
	if (s < 0)
		return fail(e, "too many searches in flight (%d)", PIPE_DEPTH);
	if (!e->pipe_done[s])
		CK(e, cudaEventCreateWithFlags(&e->pipe_done[s], cudaEventDisableTiming));
	Batch &B = e->pipe[s];

	/* Slots alternate lanes, so that two searches in flight overlap on the GPU. */
	if (use_lane(e, (e->lanes_enabled && !e->external_stream && !segmented(e)) ? s % N_LANES : 0) == -1)
		return -1;

	if (segmented(e)) {
		if (run_segmented(e, B, e->seg_pipe[s], b, d_recs, s) == -1)
			return -1;
	} else {
		if (fill_batch(e, B, b) == -1)
			return -1;
		if (fz && fz->n) {
			/*
			 * Behind the descriptor copy, ahead of the token resolve, on
			 * the same stream: the vocabulary scan of the missing terms
			 * and a kernel that writes the picks into the token list.
			 * Nothing waits on the host.
			 */
			FuzzyOut o;
			int launches = 0;

			if (fuzzy_enqueue(e->fz, e->fz_pipe[s], fz->n, fz->blob, fz->off, 0, fz->n,
			    e->stream, e->n_sms, &launches, &o) != 0)
				return fail(e, "fuzzy scan failed: %s", cudaGetErrorString(cudaGetLastError()));
			CK(e, cudaMemcpyAsync(o.extra, fz->pos, (size_t)fz->n * 4, cudaMemcpyHostToDevice, e->stream));
			fuzzy_patch_tokens_kernel<<<(fz->n + 255) / 256, 256, 0, e->stream>>>(B.d_tokens,
			    o.extra, o.term, fz->n);
			e->launches += launches + 1;
			CK(e, cudaGetLastError());
		}
		if (run_batch(e, B, d_recs ? d_recs : B.d_recs) == -1)
			return -1;
	}
	if (!d_recs && enqueue_fetch(e, B) == -1)
		return -1;
	CK(e, cudaEventRecord(e->pipe_done[s], e->stream));
	e->pipe_busy[s] = true;
	return s;
}

extern "C" int
nxsb_engine_search_begin(nxsb_engine_t *e, const nxsb_batch_t *b)
{
	if (is_multi(e))
		return e->sharded ? shard_search_begin(e, b) : multi_search_begin(e, b);
	return search_begin(e, b, nullptr);
}

/*
 * As search_begin, but this shard's records stay on the device in d_recs
 * (the send buffer of the cross-shard all-gather); end the search with
 * counts = NULL once whatever consumes d_recs has been enqueued.
 */
extern "C" int
nxsb_engine_search_begin_fz(nxsb_engine_t *e, const nxsb_batch_t *b, uint32_t n_miss,
    const char *miss_blob, const uint32_t *miss_off, const uint32_t *miss_pos)
{
	if (n_miss == 0)
		return nxsb_engine_search_begin(e, b);
	for (uint32_t i = 0; i < n_miss; i++)
		if (miss_pos[i] >= b->n_tokens) {
			return fail(e, "search_begin_fz: token position %u out of range", miss_pos[i]);
		}
	nxsb_engine_t *fe = is_multi(e) ? e->replicas[0] : e;	/* who holds the vocabulary */

	if (!fe->fz.loaded)
		return fail(e, "no vocabulary image loaded");
	if (!is_multi(e) && !segmented(e)) {
		const FzPending fz = { n_miss, miss_blob, miss_off, miss_pos };

		return search_begin(e, b, nullptr, &fz);
	}
	/*
	 * Replicas (the vocabulary lives on the first) and segmented images
	 * (every segment stages its own copy of the tokens): the lookups run
	 * first and their answers go into a host copy of the token list.
	 */
	std::vector<uint32_t> term(n_miss), dist(n_miss), toks(b->tokens, b->tokens + b->n_tokens);
	nxsb_batch_t b2 = *b;

	if (nxsb_engine_fuzzy(fe, n_miss, miss_blob, miss_off, term.data(), dist.data(), nullptr) != 0)
		return is_multi(e) ? multi_fail(e, fe, 0) : -1;
	for (uint32_t i = 0; i < n_miss; i++)
		toks[miss_pos[i]] = term[i];
	b2.tokens = toks.data();
	return nxsb_engine_search_begin(e, &b2);
}

extern "C" int
nxsb_engine_search_begin_dev(nxsb_engine_t *e, const nxsb_batch_t *b, void *d_recs)
{
	NOT_ON_REPLICATED(e, "a search into caller-owned device memory");
	if (!d_recs)
		return fail(e, "search_begin_dev needs a device buffer");
	return search_begin(e, b, (Rec *)d_recs);
}

extern "C" int
nxsb_engine_search_end(nxsb_engine_t *e, int s, uint32_t *counts, uint64_t *ids,
    float *scores)
{
	if (is_multi(e))
		return e->sharded ? shard_search_end(e, s, counts, ids, scores)
		    : multi_search_end(e, s, counts, ids, scores);
	if (s < 0 || s >= PIPE_DEPTH || !e->pipe_busy[s])
		return fail(e, "bad search handle %d", s);
	e->pipe_busy[s] = false;
	CK(e, cudaSetDevice(e->device));
	CK(e, cudaEventSynchronize(e->pipe_done[s]));
	if (counts)
		unpack_results(e->pipe[s], counts, ids, scores);
	return 0;
}

extern "C" int
nxsb_engine_merge_topk(nxsb_engine_t *e, const void *d_in, uint32_t n_shards,
    uint32_t n_queries, uint32_t limit, void *d_out)
{
	NOT_ON_REPLICATED(e, "the cross-shard merge");
	const unsigned long long total = (unsigned long long)n_queries * n_shards * limit;

	CK(e, cudaSetDevice(e->device));
	CK(e, cudaMemsetAsync(d_out, 0, (size_t)n_queries * limit * sizeof(Rec), e->stream));
	if (total) {
		merge_topk_kernel<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(
		    (const Rec *)d_in, n_shards, n_queries, limit, (Rec *)d_out);
		e->launches++;
		CK(e, cudaGetLastError());
	}
	return 0;
}

extern "C" int
nxsb_engine_timings(nxsb_engine_t *e, uint32_t last_runs, const char **names,
    float *ms, int cap)
{
	if (is_multi(e))
		return nxsb_engine_timings(e->replicas[0], last_runs, names, ms, cap);
	int n = 0;

	if (cudaStreamSynchronize(e->stream) != cudaSuccess)
		return 0;
	if (last_runs > e->n_runs)
		last_runs = (uint32_t)e->n_runs;
	if (last_runs > EV_RING)
		last_runs = EV_RING;
	for (uint64_t run = e->n_runs - last_runs; run < e->n_runs; run++) {
		const nxsb_engine::RunEvents &r = e->runs[run % EV_RING];

		for (int i = 0; i + 1 < r.n; i++) {
			float t = 0;
			int k;

			if (cudaEventElapsedTime(&t, r.ev[i], r.ev[i + 1]) != cudaSuccess)
				continue;
			for (k = 0; k < n; k++)
				if (strcmp(names[k], r.names[i]) == 0)
					break;
			if (k == n) {
				if (n == cap)
					continue;
				names[n] = r.names[i];
				ms[n] = 0;
				n++;
			}
			ms[k] += t;
		}
	}
	return n;
}

extern "C" int
nxsb_engine_last_timings(nxsb_engine_t *e, const char **names, float *ms, int cap)
{
	return nxsb_engine_timings(e, 1, names, ms, cap);
}

/*
 * Fuzzy matching entry points (kernels in fuzzy.cuh).
 */

extern "C" int
nxsb_engine_load_vocab(nxsb_engine_t *e, uint32_t n_terms, const char *blob,
    const uint32_t *term_off, const uint64_t *term_total,
    const uint32_t *bk_parent, const uint8_t *bk_edge, const uint32_t *bk_rank)
{
	if (is_multi(e))	/* lookups are a small share of a batch: one replica serves them */
		return nxsb_engine_load_vocab(e->replicas[0], n_terms, blob, term_off, term_total,
		    bk_parent, bk_edge, bk_rank) == 0 ? 0 : multi_fail(e, e->replicas[0], 0);
	/* The image is about to change: both lanes drain first. */
	sync_lanes(e);
	if (use_lane(e, 0) == -1)
		return -1;
	CK(e, cudaSetDevice(e->device));
	if (fuzzy_load(e->fz, n_terms, blob, term_off, term_total, bk_parent,
	    bk_edge, bk_rank, e->stream) != 0)
		return fail(e, "vocabulary upload failed: %s",
		    cudaGetErrorString(cudaGetLastError()));
	return 0;
}

extern "C" int
nxsb_engine_update_term_totals(nxsb_engine_t *e, uint32_t n_terms, const uint64_t *term_total)
{
	if (is_multi(e))
		return nxsb_engine_update_term_totals(e->replicas[0], n_terms, term_total) == 0 ? 0
		    : multi_fail(e, e->replicas[0], 0);
	/* The image is about to change: both lanes drain first. */
	sync_lanes(e);
	if (use_lane(e, 0) == -1)
		return -1;
	CK(e, cudaSetDevice(e->device));
	if (fuzzy_update_live(e->fz, n_terms, term_total, e->stream) != 0)
		return fail(e, "term totals do not match the vocabulary image (%u terms)", n_terms);
	return 0;
}

extern "C" int
nxsb_engine_fuzzy(nxsb_engine_t *e, uint32_t n, const char *qblob,
    const uint32_t *qoff, uint32_t *out_term, uint32_t *out_dist,
    uint32_t *out_true)
{
	if (is_multi(e))
		return nxsb_engine_fuzzy(e->replicas[0], n, qblob, qoff, out_term, out_dist, out_true) == 0 ? 0
		    : multi_fail(e, e->replicas[0], 0);
	CK(e, cudaSetDevice(e->device));
	if (!e->fz.loaded)
		return fail(e, "no vocabulary image loaded");
	begin_run(e);
	mark(e, "fuzzy_scan");
	int launches = 0;
	if (fuzzy_run(e->fz, e->fz_scratch, n, qblob, qoff, out_term, out_dist, out_true,
	    0, nullptr, nullptr, e->stream, e->n_sms, &launches) != 0)
		return fail(e, "fuzzy scan failed: %s",
		    cudaGetErrorString(cudaGetLastError()));
	e->launches += launches;
	mark(e, "end");
	return 0;
}

extern "C" int
nxsb_engine_fuzzy_candidates(nxsb_engine_t *e, uint32_t n, const char *qblob,
    const uint32_t *qoff, uint32_t cap, uint32_t *out_term, uint32_t *out_dist,
    uint32_t *out_n, uint32_t *cand_term, uint8_t *cand_dist, uint8_t *cand_flags)
{
	if (is_multi(e))
		return nxsb_engine_fuzzy_candidates(e->replicas[0], n, qblob, qoff, cap, out_term, out_dist,
		    out_n, cand_term, cand_dist, cand_flags) == 0 ? 0 : multi_fail(e, e->replicas[0], 0);
	CK(e, cudaSetDevice(e->device));
	if (!e->fz.loaded)
		return fail(e, "no vocabulary image loaded");
	if (cap == 0 || (uint64_t)n * cap > (1ull << 28))
		return fail(e, "candidate capacity %u x %u queries is out of range", cap, n);
	std::vector<uint4> recs((size_t)n * cap);
	std::vector<uint32_t> cnt(n), term(n), dist(n);
	int launches = 0;

	begin_run(e);
	mark(e, "fuzzy_scan");
	if (fuzzy_run(e->fz, e->fz_scratch, n, qblob, qoff, out_term ? out_term : term.data(),
	    out_dist ? out_dist : dist.data(), nullptr, cap, cnt.data(), recs.data(),
	    e->stream, e->n_sms, &launches) != 0)
		return fail(e, "fuzzy scan failed: %s", cudaGetErrorString(cudaGetLastError()));
	e->launches += launches;
	mark(e, "end");
	for (uint32_t i = 0; i < n; i++) {
		uint4 *r = recs.data() + (size_t)i * cap;
		const uint32_t m = std::min(cnt[i], cap);

		/* BFS rank order: the reached ones, in this order, are the reference's deque. */
		std::sort(r, r + m, [](const uint4 &a, const uint4 &b) { return a.x < b.x; });
		out_n[i] = cnt[i];
		for (uint32_t j = 0; j < m; j++) {
			cand_term[(size_t)i * cap + j] = r[j].y + 1;
			cand_dist[(size_t)i * cap + j] = (uint8_t)(r[j].z & 0xffu);
			cand_flags[(size_t)i * cap + j] = (uint8_t)((r[j].z >> 8) & 3u);
		}
	}
	return 0;
}
