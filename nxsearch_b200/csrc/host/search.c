/*
 * Searching: the host half of the hot path.
 *
 * Mirrors ref src/query/search.c:285-342 step for step -- parameters, index
 * sync, parse, token preparation and resolution -- and then hands the
 * document-set logic, scoring and top-k (ref run_query_logic,
 * search.c:210-278, and nxs_resp_build, results.c:182-220) to the GPU engine
 * as one batch.  Nothing is scored on the CPU; without a CUDA device the
 * search fails with NXS_ERR_SYSTEM.
 */
#define _GNU_SOURCE
#include <limits.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <unistd.h>

#include "index.h"
#include "query.h"
#include "nxsb200_tools.h"

/*
 * NXSB_HOST_PROF=1: where a batch call's host time goes, printed at exit
 * (development aid; a clock_gettime pair per stage when enabled, one branch
 * when not).
 */
#include <time.h>
#define NXS_FUZZY_MAX_LEN	64u	/* fuzzy.cuh FZ_MAX_QLEN: the Myers pattern is one 64-bit word */

enum { HP_PREPARE, HP_MERGE, HP_FUZZY, HP_BEGIN, HP_WAIT, HP_RESP, HP_N };
static struct {
	int		on;		/* 0 unknown, 1 yes, -1 no */
	double		us[HP_N];
	unsigned long	calls, queries;
} g_hp;

static void
hp_report(void)
{
	static const char *names[HP_N] = { "prepare (pool)", "merge shares", "fuzzy lookups",
	    "engine begin (layout, H2D, launches)", "engine end (wait + unpack)", "responses" };

	if (!g_hp.calls)
		return;
	fprintf(stderr, "nxsearch-b200 host profile: %lu batch calls, %lu queries\n", g_hp.calls, g_hp.queries);
	for (int i = 0; i < HP_N; i++)
		fprintf(stderr, "  %-38s %9.1f us/call  %6.3f us/query\n", names[i],
		    g_hp.us[i] / g_hp.calls, g_hp.us[i] / (g_hp.queries ? g_hp.queries : 1));
}

static inline double
hp_now(void)
{
	struct timespec ts;

	if (g_hp.on == 0) {
		g_hp.on = getenv("NXSB_HOST_PROF") ? 1 : -1;
		if (g_hp.on == 1)
			atexit(hp_report);
	}
	if (g_hp.on != 1)
		return 0;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e6 + ts.tv_nsec / 1e3;
}
#define HP_ADD(stage, t0) do { if (g_hp.on == 1) { const double _t = hp_now(); g_hp.us[stage] += _t - (t0); (t0) = _t; } } while (0)

typedef struct {
	uint64_t	limit;
	int		algo;
	bool		fuzzymatch;
} search_params_t;

static int
get_search_params(nxs_index_t *idx, nxs_params_t *params, search_params_t *sp)
{
	const char *s;
	bool fl;

	/* Defaults (search.c:88-91). */
	sp->limit = NXS_DEFAULT_RESULTS_LIMIT;
	sp->fuzzymatch = true;
	sp->algo = idx->algo;
	if (!params)
		return 0;

	if (nxs_params_get_uint(params, "limit", &sp->limit) == 0 &&
	    (sp->limit == 0 || sp->limit > UINT_MAX)) {
		nxs_set_error(idx->nxs, NXS_ERR_INVALID, "invalid limit");
		return -1;
	}
	if ((s = nxs_params_get_str(params, "algo")) != NULL) {
		if (strcasecmp(s, "TF-IDF") == 0) {
			sp->algo = NXSB_ALGO_TFIDF;
		} else if (strcasecmp(s, "BM25") == 0) {
			sp->algo = NXSB_ALGO_BM25;
		} else {
			nxs_set_error(idx->nxs, NXS_ERR_INVALID, "invalid algorithm");
			return -1;
		}
	}
	if (nxs_params_get_bool(params, "fuzzymatch", &fl) == 0 && !fl)
		sp->fuzzymatch = false;
	return 0;
}

/*
 * The host half of a batch is parse + tokenize + exact term lookup per query,
 * independent read-only work: it is spread over a persistent pool of worker
 * threads, each building the descriptors, token ids and programs of its share
 * of the queries in its own buffers; the shares are then laid end to end.
 * A typical query allocates nothing but its leaf strings.
 */

/* The distinct tokens of one query, in token-list order (<= the engine's limit). */
typedef struct {
	struct {
		const char *	str;
		uint32_t	len;
		uint32_t	term_id;	/* 0 = no exact term */
	} tok[NXSB_MAX_QUERY_TOKENS];
	uint32_t	count;
	bool		overflow;	/* more distinct tokens than the engine takes */
	char		buf[640];	/* filtered strings; longer ones go to the heap */
	uint32_t	used;
	char *		heap[NXSB_MAX_QUERY_TOKENS];
	uint32_t	n_heap;
} qtokens_t;

typedef struct {
	nxs_err_t	code;
	char *		msg;
} qerror_t;

/* One worker's share of a batch: queries [lo, hi). */
typedef struct {
	size_t		lo, hi;
	uint32_t *	tokens;
	int32_t *	prog;
	uint32_t	n_tok, cap_tok, n_prog, cap_prog;
	/* tokens without an exact term (fuzzymatch): position in tokens[] + string */
	uint32_t *	miss_pos;
	uint32_t *	miss_off;	/* [n_miss + 1] into miss_blob */
	char *		miss_blob;
	uint32_t	n_miss, cap_miss, blob_len, blob_cap;
	size_t		n_run;
	bool		oom;
} share_t;

typedef struct {
	nxs_index_t *		idx;
	const search_params_t *	sp;
	const char *const *	queries;
	size_t			n;
	nxsb_query_t *		descs;		/* [n], offsets relative to the share */
	qerror_t *		errs;		/* [n], code 0 = fine */
	share_t *		shares;
} prep_job_t;

/* Queries per worker below which threads are not worth waking. */
#define PREP_MIN_PER_THREAD	64
#define PREP_MAX_THREADS	16

static void
query_fail(qerror_t *e, nxs_err_t code, const char *fmt, ...)
{
	va_list ap;

	e->code = code;
	va_start(ap, fmt);
	if (vasprintf(&e->msg, fmt, ap) == -1)
		e->msg = NULL;
	va_end(ap);
}

/*
 * One leaf through the filter pipeline into the query's token list
 * (tokenizer.c tokenize_value without the allocations): *slot = its index, or
 * -1 if a filter discarded it.
 */
static int
leaf_token(const filter_pipeline_t *fp, qtokens_t *qt, const char *val,
    size_t len, int32_t *slot)
{
	char *buf;
	bool on_heap = false;

	*slot = -1;
	if (qt->used + len + 1 <= sizeof(qt->buf)) {
		buf = qt->buf + qt->used;
	} else {
		if ((buf = malloc(len + 1)) == NULL)
			return -1;
		on_heap = true;
	}
	memcpy(buf, val, len);
	buf[len] = '\0';
	if (!filter_apply(fp, buf, &len)) {
		if (on_heap)
			free(buf);
		return 0;
	}
	/* Same string, same token: scored once (ref tokenizer.c:94-117). */
	for (uint32_t i = 0; i < qt->count; i++) {
		if (qt->tok[i].len == len && memcmp(qt->tok[i].str, buf, len) == 0) {
			if (on_heap)
				free(buf);
			*slot = (int32_t)i;
			return 0;
		}
	}
	if (qt->count == NXSB_MAX_QUERY_TOKENS) {
		if (on_heap)
			free(buf);
		qt->overflow = true;
		return 0;
	}
	if (on_heap)
		qt->heap[qt->n_heap++] = buf;
	else
		qt->used += len + 1;
	qt->tok[qt->count].str = buf;
	qt->tok[qt->count].len = len;
	qt->tok[qt->count].term_id = 0;
	*slot = (int32_t)qt->count++;
	return 0;
}

/*
 * query_prepare (ref query.c:75-115): walk the tree with a LIFO so that the
 * RIGHT-most leaf is tokenized first; the order of first appearance in that
 * walk is the token-list order, i.e. the score summation order.
 */
static int
prepare_query(const filter_pipeline_t *fp, qtree_t *t, qtokens_t *qt)
{
	int32_t small[2 * QTREE_INLINE_NODES + 2], *stack = small;
	int32_t sp = 0;
	int ret = 0;

	qt->count = qt->used = qt->n_heap = 0;
	qt->overflow = false;
	if (t->root < 0)
		return 0;
	if (t->n_nodes > 2 * QTREE_INLINE_NODES &&
	    (stack = malloc(sizeof(int32_t) * (t->n_nodes + 1))) == NULL)
		return -1;
	stack[sp++] = t->root;
	while (sp) {
		qnode_t *n = &t->nodes[stack[--sp]];

		if (n->type != QN_VALUE) {
			stack[sp++] = n->left;
			stack[sp++] = n->right;
			continue;
		}
		if (leaf_token(fp, qt, n->value, strlen(n->value), &n->token) == -1) {
			ret = -1;
			break;
		}
	}
	if (stack != small)
		free(stack);
	return ret;
}

/*
 * Operand stack the postfix program of qtree_emit_program() needs: the engine
 * evaluates it with a fixed stack (engine.cu validate_batch), and a query that
 * does not fit must fail on its own, not take the batch down.
 */
static uint32_t
program_stack_depth(const qtree_t *t, int32_t node)
{
	const qnode_t *nd = &t->nodes[node];

	if (nd->type == QN_VALUE)
		return 1;
	const uint32_t l = program_stack_depth(t, nd->left);
	const uint32_t r = program_stack_depth(t, nd->right) + 1;

	return l > r ? l : r;
}

static int
share_reserve(share_t *sh, uint32_t ntok, uint32_t nprog, uint32_t nmiss, uint32_t nblob)
{
#define GROW(ptr, cap, need, type) do {						\
	if ((need) > (cap)) {							\
		const uint32_t _nc = (need) + (need) / 2 + 64;			\
		type *_p = realloc((ptr), sizeof(type) * _nc);			\
		if (!_p)							\
			return -1;						\
		(ptr) = _p;							\
		(cap) = _nc;							\
	}									\
} while (0)
	GROW(sh->tokens, sh->cap_tok, sh->n_tok + ntok, uint32_t);
	GROW(sh->prog, sh->cap_prog, sh->n_prog + nprog, int32_t);
	if (nmiss) {
		const uint32_t need = sh->n_miss + nmiss + 1;

		if (need > sh->cap_miss) {
			const uint32_t nc = need + need / 2 + 16;
			uint32_t *a = realloc(sh->miss_pos, sizeof(uint32_t) * nc);
			uint32_t *b = a ? realloc(sh->miss_off, sizeof(uint32_t) * nc) : NULL;

			if (a)
				sh->miss_pos = a;
			if (b)
				sh->miss_off = b;
			if (!a || !b)
				return -1;
			sh->cap_miss = nc;
		}
		GROW(sh->miss_blob, sh->blob_cap, sh->blob_len + nblob, char);
	}
#undef GROW
	return 0;
}

/*
 * Everything the host does for one query (ref search.c:285-342 up to
 * run_query_logic): parse, query_prepare, exact tokenset_resolve, the limits,
 * and its part of the batch arrays.
 */
static void
prepare_one(const prep_job_t *job, share_t *sh, size_t i)
{
	const nxs_index_t *idx = job->idx;
	nxsb_query_t *d = &job->descs[i];
	qerror_t *err = &job->errs[i];
	qtree_t tree;
	qtokens_t qt;
	uint32_t n_resolved = 0, n_miss = 0, miss_bytes = 0, np = 0;

	qt.n_heap = 0;
	qtree_parse(&tree, job->queries[i]);
	if (tree.error) {
		query_fail(err, NXS_ERR_INVALID, "query failed with %s",
		    tree.errmsg ? tree.errmsg : "out of memory");
		goto out;
	}
	if (prepare_query(idx->fp, &tree, &qt) == -1) {
		query_fail(err, NXS_ERR_FATAL, "query_prepare() failed");
		goto out;
	}
	/* tokenset_resolve (tokenizer.c:160-199): exact lookups. */
	for (uint32_t j = 0; j < qt.count; j++) {
		qt.tok[j].term_id = idx_term_lookup(idx, qt.tok[j].str, qt.tok[j].len);
		if (qt.tok[j].term_id) {
			n_resolved++;
		} else if (job->sp->fuzzymatch) {
			/*
			 * The GPU scan takes patterns of up to NXS_FUZZY_MAX_LEN
			 * bytes.  A longer one can only match a term within 2
			 * bytes of its length: none that long in the vocabulary
			 * means "no match" is the exact answer; otherwise the
			 * query fails alone, loudly, rather than miss silently.
			 */
			if (qt.tok[j].len > NXS_FUZZY_MAX_LEN &&
			    idx->max_term_len + 2 >= qt.tok[j].len) {
				query_fail(err, NXS_ERR_LIMIT, "fuzzy match of a %u-byte term is "
				    "not supported (limit %u bytes)", (unsigned)qt.tok[j].len,
				    NXS_FUZZY_MAX_LEN);
				goto out;
			}
			n_miss++;
			miss_bytes += qt.tok[j].len;
		}
	}
	/* search.c:224-226: nothing usable => empty result, no error. */
	if (tree.root < 0 || (n_resolved == 0 && n_miss == 0))
		goto out;
	/* search.c:126-131 (the recursion guard of get_expr_bitmap). */
	if (tree.depth > NXS_QUERY_RLIMIT) {
		query_fail(err, NXS_ERR_LIMIT, "query nesting limit reached (%u levels)",
		    NXS_QUERY_RLIMIT);
		goto out;
	}
	if (qt.overflow || (uint32_t)tree.n_nodes > NXSB_MAX_QUERY_PROG ||
	    program_stack_depth(&tree, tree.root) > NXSB_MAX_QUERY_TOKENS + 1) {
		query_fail(err, NXS_ERR_LIMIT, "query too large for the GPU engine "
		    "(%s%d nodes; limits %u terms / %u nodes)",
		    qt.overflow ? "too many terms, " : "", tree.n_nodes,
		    NXSB_MAX_QUERY_TOKENS, NXSB_MAX_QUERY_PROG);
		goto out;
	}
	if (share_reserve(sh, qt.count, (uint32_t)tree.n_nodes, n_miss, miss_bytes) == -1) {
		sh->oom = true;
		goto out;
	}
	d->tok_off = sh->n_tok;
	d->n_tokens = qt.count;
	for (uint32_t j = 0; j < qt.count; j++) {
		if (!qt.tok[j].term_id && job->sp->fuzzymatch) {
			sh->miss_pos[sh->n_miss] = sh->n_tok;
			sh->miss_off[sh->n_miss++] = sh->blob_len;
			memcpy(sh->miss_blob + sh->blob_len, qt.tok[j].str, qt.tok[j].len);
			sh->blob_len += qt.tok[j].len;
		}
		sh->tokens[sh->n_tok++] = qt.tok[j].term_id;
	}
	d->prog_off = sh->n_prog;
	qtree_emit_program(&tree, tree.root, sh->prog + sh->n_prog, &np);
	d->n_prog = np;
	sh->n_prog += np;
	sh->n_run++;
out:
	for (uint32_t j = 0; j < qt.n_heap; j++)
		free(qt.heap[j]);
	qtree_free(&tree);
}

static void
prepare_share(void *arg, unsigned w, unsigned nw)
{
	prep_job_t *job = arg;
	share_t *sh = &job->shares[w];

	sh->lo = job->n * w / nw;
	sh->hi = job->n * (w + 1) / nw;
	for (size_t i = sh->lo; i < sh->hi; i++)
		prepare_one(job, sh, i);
}

/*
 * A small persistent pool: the reference is single-threaded, and so is this
 * API (one nxs_t per thread); the pool only ever serves one batch at a time
 * -- a second caller finding it busy does its work itself.
 */
typedef void (*pool_fn_t)(void *, unsigned worker, unsigned nworkers);

static struct {
	pthread_mutex_t	user;		/* one batch at a time */
	pthread_mutex_t	lock;
	pthread_cond_t	wake, done;
	pthread_t	threads[PREP_MAX_THREADS];
	unsigned	n_threads;	/* started so far */
	unsigned	generation, helpers, pending;
	pool_fn_t	fn;
	void *		arg;
} g_pool = {
	.user = PTHREAD_MUTEX_INITIALIZER, .lock = PTHREAD_MUTEX_INITIALIZER,
	.wake = PTHREAD_COND_INITIALIZER, .done = PTHREAD_COND_INITIALIZER,
};

static void *
pool_worker(void *arg)
{
	const unsigned me = (unsigned)(uintptr_t)arg;
	unsigned seen = 0;

	pthread_mutex_lock(&g_pool.lock);
	for (;;) {
		while (g_pool.generation == seen)
			pthread_cond_wait(&g_pool.wake, &g_pool.lock);
		seen = g_pool.generation;
		if (me >= g_pool.helpers)
			continue;
		pool_fn_t fn = g_pool.fn;
		void *a = g_pool.arg;
		const unsigned nw = g_pool.helpers + 1;

		pthread_mutex_unlock(&g_pool.lock);
		fn(a, me + 1, nw);
		pthread_mutex_lock(&g_pool.lock);
		if (--g_pool.pending == 0)
			pthread_cond_signal(&g_pool.done);
	}
	return NULL;
}

/* Run fn(arg, w, nw) for w in [0, nw) with nw <= want; returns nw. */
static unsigned
pool_run(unsigned want, pool_fn_t fn, void *arg, unsigned (*plan)(void *, unsigned))
{
	unsigned helpers;

	if (want < 2 || pthread_mutex_trylock(&g_pool.user) != 0) {
		plan(arg, 1);
		fn(arg, 0, 1);
		return 1;
	}
	pthread_mutex_lock(&g_pool.lock);
	while (g_pool.n_threads < want - 1 && g_pool.n_threads < PREP_MAX_THREADS) {
		if (pthread_create(&g_pool.threads[g_pool.n_threads], NULL, pool_worker,
		    (void *)(uintptr_t)g_pool.n_threads) != 0)
			break;
		pthread_detach(g_pool.threads[g_pool.n_threads]);
		g_pool.n_threads++;
	}
	helpers = want - 1 < g_pool.n_threads ? want - 1 : g_pool.n_threads;
	plan(arg, helpers + 1);
	g_pool.fn = fn;
	g_pool.arg = arg;
	g_pool.helpers = helpers;
	g_pool.pending = helpers;
	g_pool.generation++;
	pthread_cond_broadcast(&g_pool.wake);
	pthread_mutex_unlock(&g_pool.lock);

	fn(arg, 0, helpers + 1);

	pthread_mutex_lock(&g_pool.lock);
	while (g_pool.pending)
		pthread_cond_wait(&g_pool.done, &g_pool.lock);
	pthread_mutex_unlock(&g_pool.lock);
	pthread_mutex_unlock(&g_pool.user);
	return helpers + 1;
}

static unsigned
prepare_plan(void *arg, unsigned nw)
{
	prep_job_t *job = arg;

	job->shares = calloc(nw, sizeof(share_t));
	return nw;
}

/*
 * A batch between its two halves: what _end needs to turn the engine's
 * arrays into responses.
 */
struct nxs_batch {
	nxs_index_t *	idx;
	size_t		n;
	bool *		failed;		/* per query: no response for it */
	uint32_t	k;		/* result stride */
	int		handle;		/* engine search in flight, or -1 */
};

static void
batch_free(nxs_batch_t *bt)
{
	if (bt) {
		free(bt->failed);
		free(bt);
	}
}

/*
 * First half: everything on the host (parameters, index sync, parse, token
 * resolution) and the submission of the batch to the device.  Returns without
 * waiting for the GPU.
 */
NXS_API nxs_batch_t *
nxs_index_search_batch_begin(nxs_index_t *idx, nxs_params_t *params,
    const char *const *queries, size_t n)
{
	nxs_t *nxs = idx->nxs;
	search_params_t sp;
	prep_job_t job = { 0 };
	uint32_t *tokens = NULL, *miss_pos = NULL, *miss_off = NULL;
	int32_t *prog = NULL;
	char *miss_blob = NULL;
	size_t n_tok = 0, n_prog = 0, n_miss = 0, n_run = 0, blob_len = 0;
	unsigned nw = 0;
	nxs_batch_t *bt = NULL;
	uint32_t k;
	int ret = -1;
	double hp_t = hp_now();

	nxs_clear_error(nxs);
	if (get_search_params(idx, params, &sp) == -1)
		return NULL;
	if (sp.algo != NXSB_ALGO_BM25 && sp.algo != NXSB_ALGO_TFIDF) {
		nxs_set_error(nxs, NXS_ERR_INVALID, "invalid algorithm");
		return NULL;
	}

	/* Pick up what other processes appended (search.c:309-310). */
	if (idx_terms_sync(idx) == -1 || idx_dtmap_sync(idx, true) == -1)
		return NULL;

	if ((bt = calloc(1, sizeof(*bt))) == NULL ||
	    (bt->failed = calloc(n ? n : 1, sizeof(bool))) == NULL)
		goto oom;
	bt->idx = idx;
	bt->n = n;
	bt->handle = -1;

	job.idx = idx;
	job.sp = &sp;
	job.queries = queries;
	job.n = n;
	job.descs = calloc(n ? n : 1, sizeof(nxsb_query_t));
	job.errs = calloc(n ? n : 1, sizeof(qerror_t));
	if (!job.descs || !job.errs)
		goto oom;
	{
		long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
		size_t want = n / PREP_MIN_PER_THREAD;

		if (ncpu < 1)
			ncpu = 1;
		if (want > (size_t)ncpu)
			want = ncpu;
		if (want > PREP_MAX_THREADS)
			want = PREP_MAX_THREADS;
		nw = pool_run((unsigned)want, prepare_share, &job, prepare_plan);
		if (!job.shares)
			goto oom;
	}
	HP_ADD(HP_PREPARE, hp_t);

	/* Errors in query order (the slot keeps the last one, as a loop of single searches would). */
	for (size_t i = 0; i < n; i++) {
		if (job.errs[i].code) {
			nxs_set_error(nxs, job.errs[i].code, "%s",
			    job.errs[i].msg ? job.errs[i].msg : "out of memory");
			bt->failed[i] = true;
		}
	}
	/* Lay the shares end to end. */
	for (unsigned w = 0; w < nw; w++) {
		const share_t *sh = &job.shares[w];

		if (sh->oom)
			goto oom;
		n_tok += sh->n_tok;
		n_prog += sh->n_prog;
		n_miss += sh->n_miss;
		blob_len += sh->blob_len;
		n_run += sh->n_run;
	}
	tokens = malloc(sizeof(uint32_t) * (n_tok + 1));
	prog = malloc(sizeof(int32_t) * (n_prog + 1));
	if (!tokens || !prog)
		goto oom;
	if (n_miss) {
		miss_pos = malloc(sizeof(uint32_t) * n_miss);
		miss_off = malloc(sizeof(uint32_t) * (n_miss + 1));
		miss_blob = malloc(blob_len + 1);
		if (!miss_pos || !miss_off || !miss_blob)
			goto oom;
	}
	n_tok = n_prog = n_miss = blob_len = 0;
	for (unsigned w = 0; w < nw; w++) {
		const share_t *sh = &job.shares[w];

		memcpy(tokens + n_tok, sh->tokens, sizeof(uint32_t) * sh->n_tok);
		memcpy(prog + n_prog, sh->prog, sizeof(int32_t) * sh->n_prog);
		for (size_t i = sh->lo; i < sh->hi; i++) {
			if (job.descs[i].n_tokens) {
				job.descs[i].tok_off += n_tok;
				job.descs[i].prog_off += n_prog;
			}
		}
		for (uint32_t m = 0; m < sh->n_miss; m++) {
			miss_pos[n_miss + m] = sh->miss_pos[m] + n_tok;
			miss_off[n_miss + m] = sh->miss_off[m] + blob_len;
		}
		if (sh->n_miss)
			memcpy(miss_blob + blob_len, sh->miss_blob, sh->blob_len);
		n_tok += sh->n_tok;
		n_prog += sh->n_prog;
		n_miss += sh->n_miss;
		blob_len += sh->blob_len;
	}

	HP_ADD(HP_MERGE, hp_t);
	/*
	 * Tokens that missed keep id 0 here; their strings go to the engine with
	 * the batch, which scans the vocabulary for them on the GPU and writes
	 * the picks into the device copy of the token list before it scores --
	 * no wait on the host (nxsb_engine_search_begin_fz).
	 */
	if (n_miss && idx->n_terms) {
		miss_off[n_miss] = blob_len;
		if (idx_gpu_prepare(idx, true) == -1)
			goto out;
	} else {
		n_miss = 0;
	}
	HP_ADD(HP_FUZZY, hp_t);
	/* More results than live documents cannot exist: clamp the limit. */
	k = sp.limit > idx->n_live ? idx->n_live : (uint32_t)sp.limit;
	if (k == 0)
		k = 1;
	bt->k = k;

	if (n_run) {
		const nxsb_batch_t batch = {
			.algo = sp.algo, .limit = k, .n_queries = n,
			.queries = job.descs, .tokens = tokens, .n_tokens = n_tok,
			.prog = prog, .n_prog = n_prog,
		};

		if (idx_gpu_prepare(idx, false) == -1)
			goto out;
		bt->handle = nxsb_engine_search_begin_fz(idx->engine, &batch, (uint32_t)n_miss,
		    miss_blob, miss_off, miss_pos);
		if (bt->handle == -1) {
			nxs_set_error(nxs, NXS_ERR_SYSTEM, "GPU search failed: %s",
			    nxsb_engine_errmsg(idx->engine));
			goto out;
		}
	}
	ret = 0;
	HP_ADD(HP_BEGIN, hp_t);
	if (g_hp.on == 1) {
		g_hp.calls++;
		g_hp.queries += n;
	}
	goto out;
oom:
	nxs_set_error(nxs, NXS_ERR_SYSTEM, "out of memory");
out:
	if (ret != 0) {
		batch_free(bt);
		bt = NULL;
		nxs_error_checkpoint(nxs);
	}
	for (unsigned w = 0; job.shares && w < nw; w++) {
		share_t *sh = &job.shares[w];

		free(sh->tokens);
		free(sh->prog);
		free(sh->miss_pos);
		free(sh->miss_off);
		free(sh->miss_blob);
	}
	free(job.shares);
	for (size_t i = 0; job.errs && i < n; i++)
		free(job.errs[i].msg);
	free(job.errs);
	free(job.descs);
	free(tokens);
	free(prog);
	free(miss_pos);
	free(miss_off);
	free(miss_blob);
	return bt;
}

/*
 * Second half: wait for the batch, build the responses, release the batch
 * (also on failure).  resps may be NULL to abandon the results.
 */
NXS_API int
nxs_index_search_batch_end(nxs_batch_t *bt, nxs_resp_t **resps)
{
	nxs_index_t *idx;
	nxs_t *nxs;
	uint32_t *counts = NULL;
	uint64_t *ids = NULL;
	float *scores = NULL;
	size_t n;
	uint32_t k;
	int ret = -1;

	if (!bt)
		return -1;
	double hp_t = hp_now();
	idx = bt->idx;
	nxs = idx->nxs;
	n = bt->n;
	k = bt->k;
	for (size_t i = 0; resps && i < n; i++)
		resps[i] = NULL;
	if (bt->handle >= 0) {
		if (resps) {
			counts = calloc(n, sizeof(uint32_t));
			ids = malloc(sizeof(uint64_t) * n * k);
			scores = malloc(sizeof(float) * n * k);
			if (!counts || !ids || !scores) {
				nxsb_engine_search_end(idx->engine, bt->handle, NULL, NULL, NULL);
				nxs_set_error(nxs, NXS_ERR_SYSTEM, "out of memory");
				goto out;
			}
		}
		if (nxsb_engine_search_end(idx->engine, bt->handle, counts, ids,
		    scores) == -1) {
			nxs_set_error(nxs, NXS_ERR_SYSTEM, "GPU search failed: %s",
			    nxsb_engine_errmsg(idx->engine));
			goto out;
		}
	}
	HP_ADD(HP_WAIT, hp_t);
	for (size_t i = 0; resps && i < n; i++) {
		if (bt->failed[i])
			continue;
		resps[i] = nxs_resp_from_arrays(ids ? ids + i * k : NULL,
		    scores ? scores + i * k : NULL, counts ? counts[i] : 0);
		if (!resps[i]) {
			nxs_set_error(nxs, NXS_ERR_SYSTEM, "out of memory");
			goto out;
		}
	}
	ret = 0;
	HP_ADD(HP_RESP, hp_t);
out:
	if (ret != 0) {
		for (size_t i = 0; resps && i < n; i++) {
			if (resps[i])
				nxs_resp_release(resps[i]);
			resps[i] = NULL;
		}
		nxs_error_checkpoint(nxs);
	}
	free(counts);
	free(ids);
	free(scores);
	batch_free(bt);
	return ret;
}

NXS_API int
nxs_index_search_batch(nxs_index_t *idx, nxs_params_t *params,
    const char *const *queries, size_t n, nxs_resp_t **resps)
{
	nxs_batch_t *bt;

	for (size_t i = 0; i < n; i++)
		resps[i] = NULL;
	if ((bt = nxs_index_search_batch_begin(idx, params, queries, n)) == NULL)
		return -1;
	return nxs_index_search_batch_end(bt, resps);
}

NXS_API nxs_resp_t *
nxs_index_search(nxs_index_t *idx, nxs_params_t *params, const char *query,
    size_t len)
{
	nxs_resp_t *resp = NULL;

	(void)len;	/* ignored, as in the reference (search.c:177) */
	if (nxs_index_search_batch(idx, params, &query, 1, &resp) == -1)
		return NULL;
	return resp;
}
