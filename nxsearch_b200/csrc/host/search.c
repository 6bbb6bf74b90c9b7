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

typedef struct {
	uint64_t	limit;
	int		algo;
	bool		fuzzymatch;
} search_params_t;

typedef struct {
	qtree_t		tree;
	tokenset_t *	tokens;
	int32_t *	slot_map;	/* token slot -> compacted slot, -1 = trimmed */
	uint32_t	n_resolved;
	uint32_t	n_miss;		/* tokens without an exact term */
	bool		failed;
	/* Error of this query, reported in query order after the workers join. */
	nxs_err_t	err_code;
	char *		err_msg;
} prepared_t;

/* Queries per worker below which threads are not worth starting. */
#define PREP_MIN_PER_THREAD	128
#define PREP_MAX_THREADS	16

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
 * query_prepare (ref query.c:75-115): walk the tree with a LIFO so that the
 * RIGHT-most leaf is tokenized first; the order of first appearance in that
 * walk is the token-list order, i.e. the score summation order.
 */
static int
prepare_query(filter_pipeline_t *fp, prepared_t *pq)
{
	qtree_t *t = &pq->tree;
	int32_t *stack;
	int32_t sp = 0;

	if ((pq->tokens = tokenset_create()) == NULL)
		return -1;
	if (t->root < 0)
		return 0;
	if ((stack = malloc(sizeof(int32_t) * (t->n_nodes + 1))) == NULL)
		return -1;
	stack[sp++] = t->root;
	while (sp) {
		qnode_t *n = &t->nodes[stack[--sp]];

		if (n->type != QN_VALUE) {
			stack[sp++] = n->left;
			stack[sp++] = n->right;
			continue;
		}
		if (tokenize_value(fp, pq->tokens, n->value, strlen(n->value),
		    &n->token) == -1) {
			free(stack);
			return -1;
		}
	}
	free(stack);
	return 0;
}

/* Post-order emission of the boolean program (depth is bounded by then). */
static void
emit_program(const prepared_t *pq, int32_t node, int32_t *prog, uint32_t *n)
{
	const qnode_t *nd = &pq->tree.nodes[node];

	if (nd->type == QN_VALUE) {
		const int32_t slot = nd->token >= 0 ? pq->slot_map[nd->token] : -1;

		/*
		 * A leaf without a usable term is the empty set.  The reference
		 * does this for filter-discarded leaves (search.c:133-141); for
		 * unresolved ones it reads freed memory -- the
		 * empty set is the defined behaviour here.
		 */
		prog[(*n)++] = slot >= 0 ? slot : NXSB_OP_EMPTY;
		return;
	}
	emit_program(pq, nd->left, prog, n);
	emit_program(pq, nd->right, prog, n);
	prog[(*n)++] = nd->type == QN_AND ? NXSB_OP_AND :
	    nd->type == QN_OR ? NXSB_OP_OR : NXSB_OP_ANDNOT;
}

/*
 * Operand stack the postfix program of emit_program() needs: the engine
 * evaluates it with a fixed stack (engine.cu validate_batch), and a query that
 * does not fit must fail on its own, not take the batch down.
 */
static uint32_t
program_stack_depth(const prepared_t *pq, int32_t node)
{
	const qnode_t *nd = &pq->tree.nodes[node];

	if (nd->type == QN_VALUE)
		return 1;
	const uint32_t l = program_stack_depth(pq, nd->left);
	const uint32_t r = program_stack_depth(pq, nd->right) + 1;

	return l > r ? l : r;
}

static void
prepared_release(prepared_t *pq)
{
	qtree_free(&pq->tree);
	tokenset_destroy(pq->tokens);
	free(pq->slot_map);
	free(pq->err_msg);
}

static void
prep_fail(prepared_t *pq, nxs_err_t code, const char *fmt, ...)
{
	va_list ap;

	pq->failed = true;
	pq->err_code = code;
	va_start(ap, fmt);
	if (vasprintf(&pq->err_msg, fmt, ap) == -1)
		pq->err_msg = NULL;
	va_end(ap);
}

/* Steps 1-3 of one query: parse, query_prepare, tokenset_resolve (exact). */
static void
prepare_one(nxs_index_t *idx, const search_params_t *sp, const char *query,
    prepared_t *pq)
{
	qtree_parse(&pq->tree, query);
	if (pq->tree.error) {
		prep_fail(pq, NXS_ERR_INVALID, "query failed with %s",
		    pq->tree.errmsg ? pq->tree.errmsg : "out of memory");
		return;
	}
	if (prepare_query(idx->fp, pq) == -1) {
		prep_fail(pq, NXS_ERR_FATAL, "query_prepare() failed");
		return;
	}
	/* tokenset_resolve (tokenizer.c:160-199): exact lookups. */
	for (uint32_t j = 0; j < pq->tokens->count; j++) {
		token_t *t = &pq->tokens->list[j];

		t->term_id = idx_term_lookup(idx, t->str, t->len);
		if (!t->term_id && sp->fuzzymatch)
			pq->n_miss++;
	}
}

typedef struct {
	nxs_index_t *		idx;
	const search_params_t *	sp;
	const char *const *	queries;
	prepared_t *		pq;
	size_t			lo, hi;
} prep_job_t;

static void *
prepare_worker(void *arg)
{
	prep_job_t *j = arg;

	for (size_t i = j->lo; i < j->hi; i++)
		prepare_one(j->idx, j->sp, j->queries[i], &j->pq[i]);
	return NULL;
}

static void
prepare_all(nxs_index_t *idx, const search_params_t *sp,
    const char *const *queries, size_t n, prepared_t *pq)
{
	pthread_t tid[PREP_MAX_THREADS];
	prep_job_t job[PREP_MAX_THREADS];
	long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
	size_t nt = n / PREP_MIN_PER_THREAD;

	if (ncpu < 1)
		ncpu = 1;
	if (nt > (size_t)ncpu)
		nt = ncpu;
	if (nt > PREP_MAX_THREADS)
		nt = PREP_MAX_THREADS;
	if (nt < 2) {
		prep_job_t all = { idx, sp, queries, pq, 0, n };

		prepare_worker(&all);
		return;
	}
	bool started[PREP_MAX_THREADS] = { false };

	for (size_t t = 0; t < nt; t++) {
		job[t] = (prep_job_t){ idx, sp, queries, pq, n * t / nt, n * (t + 1) / nt };
		/* The caller takes the last share, and any share whose thread fails to start. */
		if (t + 1 < nt)
			started[t] = pthread_create(&tid[t], NULL, prepare_worker, &job[t]) == 0;
		if (!started[t])
			prepare_worker(&job[t]);
	}
	for (size_t t = 0; t < nt; t++)
		if (started[t])
			pthread_join(tid[t], NULL);
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
	prepared_t *pq = NULL;
	nxsb_query_t *descs = NULL;
	uint32_t *tokens = NULL;
	int32_t *prog = NULL;
	size_t n_tok = 0, n_prog = 0, n_miss = 0, n_run = 0;
	nxs_batch_t *bt = NULL;
	uint32_t k;
	int ret = -1;

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
		goto out;
	bt->idx = idx;
	bt->n = n;
	bt->handle = -1;
	if ((pq = calloc(n ? n : 1, sizeof(prepared_t))) == NULL)
		goto out;

	/*
	 * Parse, tokenize and resolve every query (exact lookups).  The
	 * reference is single-threaded; a batch is independent read-only work
	 * per query, so it is spread over the host cores here.
	 */
	prepare_all(idx, &sp, queries, n, pq);
	for (size_t i = 0; i < n; i++) {
		if (pq[i].failed) {
			nxs_set_error(nxs, pq[i].err_code, "%s",
			    pq[i].err_msg ? pq[i].err_msg : "out of memory");
		}
		n_miss += pq[i].n_miss;
	}

	/* One batched fuzzy scan for every token that missed. */
	if (n_miss && idx->n_terms) {
		uint32_t *qoff = malloc(sizeof(uint32_t) * (n_miss + 1));
		uint32_t *oterm = malloc(sizeof(uint32_t) * n_miss);
		uint32_t *odist = malloc(sizeof(uint32_t) * n_miss);
		size_t blob_len = 0, m = 0;
		char *blob;
		int rc = -1;

		for (size_t i = 0; i < n; i++) {
			for (uint32_t j = 0; !pq[i].failed && j < pq[i].tokens->count; j++) {
				if (!pq[i].tokens->list[j].term_id)
					blob_len += pq[i].tokens->list[j].len;
			}
		}
		blob = malloc(blob_len + 1);
		if (qoff && oterm && odist && blob) {
			blob_len = 0;
			for (size_t i = 0; i < n; i++) {
				for (uint32_t j = 0; !pq[i].failed && j < pq[i].tokens->count; j++) {
					const token_t *t = &pq[i].tokens->list[j];

					if (t->term_id)
						continue;
					qoff[m++] = blob_len;
					memcpy(blob + blob_len, t->str, t->len);
					blob_len += t->len;
				}
			}
			qoff[m] = blob_len;
			if (idx_gpu_prepare(idx, true) == 0) {
				rc = nxsb_engine_fuzzy(idx->engine, m, blob, qoff, oterm,
				    odist, NULL);
				if (rc == -1)
					nxs_set_error(nxs, NXS_ERR_SYSTEM, "GPU fuzzy match "
					    "failed: %s", nxsb_engine_errmsg(idx->engine));
			}
		}
		if (rc == 0) {
			m = 0;
			for (size_t i = 0; i < n; i++) {
				for (uint32_t j = 0; !pq[i].failed && j < pq[i].tokens->count; j++) {
					token_t *t = &pq[i].tokens->list[j];

					if (!t->term_id)
						t->term_id = oterm[m++];
				}
			}
		}
		free(qoff); free(oterm); free(odist); free(blob);
		if (rc != 0)
			goto out;
	}

	/*
	 * TOKENSET_TRIM: unresolved tokens leave the list; the rest keep
	 * their order.  Size the batch arrays.
	 */
	for (size_t i = 0; i < n; i++) {
		tokenset_t *ts = pq[i].tokens;

		if (pq[i].failed)
			continue;
		if ((pq[i].slot_map = malloc(sizeof(int32_t) * (ts->count + 1))) == NULL)
			goto out;
		for (uint32_t j = 0; j < ts->count; j++)
			pq[i].slot_map[j] = ts->list[j].term_id ?
			    (int32_t)pq[i].n_resolved++ : -1;

		/* search.c:224-226: nothing usable => empty result, no error. */
		if (pq[i].tree.root < 0 || pq[i].n_resolved == 0)
			continue;
		/* search.c:126-131 (the recursion guard of get_expr_bitmap). */
		if (pq[i].tree.depth > NXS_QUERY_RLIMIT) {
			nxs_set_error(nxs, NXS_ERR_LIMIT,
			    "query nesting limit reached (%u levels)", NXS_QUERY_RLIMIT);
			pq[i].failed = true;
			continue;
		}
		if (pq[i].n_resolved > NXSB_MAX_QUERY_TOKENS ||
		    (uint32_t)pq[i].tree.n_nodes > NXSB_MAX_QUERY_PROG ||
		    program_stack_depth(&pq[i], pq[i].tree.root) > NXSB_MAX_QUERY_TOKENS + 1) {
			nxs_set_error(nxs, NXS_ERR_LIMIT, "query too large for the GPU "
			    "engine (%u terms, %d nodes; limits %u / %u)",
			    pq[i].n_resolved, pq[i].tree.n_nodes,
			    NXSB_MAX_QUERY_TOKENS, NXSB_MAX_QUERY_PROG);
			pq[i].failed = true;
			continue;
		}
		n_tok += pq[i].n_resolved;
		n_prog += pq[i].tree.n_nodes;
		n_run++;
	}

	descs = calloc(n ? n : 1, sizeof(nxsb_query_t));
	tokens = malloc(sizeof(uint32_t) * (n_tok + 1));
	prog = malloc(sizeof(int32_t) * (n_prog + 1));
	if (!descs || !tokens || !prog)
		goto out;
	n_tok = n_prog = 0;
	for (size_t i = 0; i < n; i++) {
		const tokenset_t *ts = pq[i].tokens;
		uint32_t np = 0;

		if (pq[i].failed || pq[i].tree.root < 0 || pq[i].n_resolved == 0)
			continue;	/* descriptor stays all-zero: empty result */
		descs[i].tok_off = n_tok;
		descs[i].n_tokens = pq[i].n_resolved;
		for (uint32_t j = 0; j < ts->count; j++) {
			if (ts->list[j].term_id)
				tokens[n_tok++] = ts->list[j].term_id;
		}
		descs[i].prog_off = n_prog;
		emit_program(&pq[i], pq[i].tree.root, prog + n_prog, &np);
		descs[i].n_prog = np;
		n_prog += np;
	}

	/* More results than live documents cannot exist: clamp the limit. */
	k = sp.limit > idx->n_live ? idx->n_live : (uint32_t)sp.limit;
	if (k == 0)
		k = 1;

	bt->k = k;

	if (n_run) {
		const nxsb_batch_t batch = {
			.algo = sp.algo, .limit = k, .n_queries = n,
			.queries = descs, .tokens = tokens, .n_tokens = n_tok,
			.prog = prog, .n_prog = n_prog,
		};

		if (idx_gpu_prepare(idx, false) == -1)
			goto out;
		bt->handle = nxsb_engine_search_begin(idx->engine, &batch);
		if (bt->handle == -1) {
			nxs_set_error(nxs, NXS_ERR_SYSTEM, "GPU search failed: %s",
			    nxsb_engine_errmsg(idx->engine));
			goto out;
		}
	}
	for (size_t i = 0; i < n; i++)
		bt->failed[i] = pq[i].failed;
	ret = 0;
out:
	if (ret != 0) {
		batch_free(bt);
		bt = NULL;
		nxs_error_checkpoint(nxs);
	}
	for (size_t i = 0; pq && i < n; i++)
		prepared_release(&pq[i]);
	free(pq);
	free(descs);
	free(tokens);
	free(prog);
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

/*
 * Introspection for the tests (include/nxsb200_tools.h).
 */

NXS_API size_t
nxsb_query_lex(const char *query, int *kinds, size_t cap)
{
	qlexer_t lx;
	qtok_t tok;
	size_t n = 0;

	qlex_init(&lx, query);
	while ((tok = qlex_next(&lx)) != QTOK_EOF) {
		if (tok == QTOK_FF_STRING || tok == QTOK_QUOTED_STRING) {
			free(lx.str);
			lx.str = NULL;
		}
		if (n < cap)
			kinds[n] = (int)tok;
		n++;
	}
	return n;
}

NXS_API char *
nxsb_query_dump(const char *query, char **errmsg)
{
	qtree_t t;
	char *out = NULL;

	if (errmsg)
		*errmsg = NULL;
	qtree_parse(&t, query);
	if (t.error) {
		if (errmsg && t.errmsg)
			*errmsg = strdup(t.errmsg);
	} else {
		out = qtree_dump(&t);
	}
	qtree_free(&t);
	return out;
}

NXS_API int
nxsb_query_compile(const char *query, char *tokens_buf, size_t buf_len,
    uint32_t *n_tokens, int32_t *prog, uint32_t prog_cap, uint32_t *n_prog)
{
	filter_pipeline_t nofilters = { 0 };
	prepared_t pq = { 0 };
	size_t off = 0;
	int ret = -1;

	*n_tokens = *n_prog = 0;
	qtree_parse(&pq.tree, query);
	if (pq.tree.error || prepare_query(&nofilters, &pq) == -1)
		goto out;
	if ((pq.slot_map = malloc(sizeof(int32_t) * (pq.tokens->count + 1))) == NULL)
		goto out;
	for (uint32_t j = 0; j < pq.tokens->count; j++) {
		const token_t *t = &pq.tokens->list[j];

		pq.slot_map[j] = j;
		if (off + t->len + 1 > buf_len)
			goto out;
		memcpy(tokens_buf + off, t->str, t->len + 1);
		off += t->len + 1;
	}
	if (pq.tree.root >= 0) {
		if ((uint32_t)pq.tree.n_nodes > prog_cap ||
		    pq.tree.depth > NXS_QUERY_RLIMIT)
			goto out;
		emit_program(&pq, pq.tree.root, prog, n_prog);
	}
	*n_tokens = pq.tokens->count;
	ret = 0;
out:
	prepared_release(&pq);
	return ret;
}
