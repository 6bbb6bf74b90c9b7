/*
 * ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See nxs_oracle.h.
 *
 * Scalar, single-threaded C restatement of the reference's scoring path.
 * Every function cites the reference lines whose behaviour it restates.
 * Deliberately simple data structures (bitsets, sorted arrays) -- the point
 * is to be obviously right, not fast.
 */
#include <math.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>

#include "nxs_oracle.h"

#define	LEV_TOLERANCE	2	/* ref index/index.h:26 */
#define	BK_EDGE_MAX	63	/* ref algo/bktree.h:11 (BKT_DIST_LIMIT) */

typedef struct bknode {
	uint32_t	term;		/* 0-based term index */
	uint8_t		nchild;
	uint8_t *	edge;		/* ascending edge labels */
	struct bknode **child;
} bknode_t;

struct ora_index {
	uint32_t	n_docs, n_terms, doc_count;
	uint64_t	token_count;
	uint64_t *	doc_ids;	/* ascending */
	uint32_t *	doc_len;
	/* Term-major postings, dense doc index ascending within a term. */
	uint64_t *	post_off;	/* [n_terms + 1] */
	uint32_t *	post_doc;
	uint32_t *	post_tf;
	/* Vocabulary. */
	char *		blob;
	uint32_t *	term_off;
	uint64_t *	term_total;
	uint32_t *	hash_slots;	/* open addressing, term index + 1 */
	size_t		hash_n;
	bknode_t *	bk_root;
};

/*
 * Index construction: the equivalent of idx_terms_sync + idx_dtmap_sync
 * (ref index/terms.c:320-414, index/dtmap.c:386-544) -- every (term, count)
 * of every live document lands in that term's document set.  Documents are
 * renumbered by ascending external id, which is the order roaring64 iterates
 * them in (ref query/search.c:235-272).
 */

typedef struct { uint64_t id; uint32_t pos; } idpos_t;

static int
idpos_cmp(const void *a, const void *b)
{
	const idpos_t *x = a, *y = b;
	return (x->id > y->id) - (x->id < y->id);
}

static uint64_t
fnv1a(const char *s, size_t n)
{
	uint64_t h = UINT64_C(0xcbf29ce484222325);

	for (size_t i = 0; i < n; i++)
		h = (h ^ (unsigned char)s[i]) * UINT64_C(0x100000001b3);
	return h;
}

ora_index_t *
ora_index_build(uint32_t n_docs, const uint64_t *doc_ids,
    const uint32_t *doc_len, const uint64_t *doc_off, const uint32_t *pairs,
    uint32_t n_terms, const char *term_blob, const uint32_t *term_off,
    const uint64_t *term_total, uint64_t token_count, uint32_t doc_count)
{
	ora_index_t *ix = calloc(1, sizeof(*ix));
	idpos_t *order = malloc(sizeof(idpos_t) * ((size_t)n_docs + 1));
	const uint64_t n_pairs = doc_off[n_docs];
	uint64_t *cursor;

	ix->n_docs = n_docs;
	ix->n_terms = n_terms;
	ix->doc_count = doc_count;
	ix->token_count = token_count;

	for (uint32_t i = 0; i < n_docs; i++)
		order[i] = (idpos_t){ doc_ids[i], i };
	qsort(order, n_docs, sizeof(idpos_t), idpos_cmp);

	ix->doc_ids = malloc(sizeof(uint64_t) * ((size_t)n_docs + 1));
	ix->doc_len = malloc(sizeof(uint32_t) * ((size_t)n_docs + 1));
	ix->post_off = calloc((size_t)n_terms + 2, sizeof(uint64_t));
	ix->post_doc = malloc(sizeof(uint32_t) * (n_pairs + 1));
	ix->post_tf = malloc(sizeof(uint32_t) * (n_pairs + 1));

	for (uint64_t j = 0; j < n_pairs; j++) {
		const uint32_t t = pairs[2 * j] - 1;
		if (t < n_terms)
			ix->post_off[t + 2]++;
	}
	for (uint32_t t = 0; t < n_terms; t++)
		ix->post_off[t + 2] += ix->post_off[t + 1];
	/*
	 * cursor[t] starts at the first slot of 0-based term t and ends one
	 * past its last, so afterwards post_off[t] .. post_off[t + 1] bounds
	 * 0-based term t (1-based id t + 1).
	 */
	cursor = ix->post_off + 1;

	for (uint32_t d = 0; d < n_docs; d++) {
		const uint32_t src = order[d].pos;

		ix->doc_ids[d] = order[d].id;
		ix->doc_len[d] = doc_len[src];
		for (uint64_t j = doc_off[src]; j < doc_off[src + 1]; j++) {
			const uint32_t t = pairs[2 * j] - 1;

			if (t >= n_terms)
				continue;
			ix->post_doc[cursor[t]] = d;
			ix->post_tf[cursor[t]] = pairs[2 * j + 1];
			cursor[t]++;
		}
	}
	free(order);

	/* Vocabulary copy + exact-match table. */
	ix->term_off = malloc(sizeof(uint32_t) * ((size_t)n_terms + 1));
	memcpy(ix->term_off, term_off, sizeof(uint32_t) * ((size_t)n_terms + 1));
	ix->blob = malloc((size_t)term_off[n_terms] + 1);
	memcpy(ix->blob, term_blob, term_off[n_terms]);
	ix->term_total = calloc((size_t)n_terms + 1, sizeof(uint64_t));
	if (term_total)
		memcpy(ix->term_total, term_total, sizeof(uint64_t) * n_terms);

	for (ix->hash_n = 64; ix->hash_n < (size_t)n_terms * 2; ix->hash_n <<= 1)
		;
	ix->hash_slots = calloc(ix->hash_n, sizeof(uint32_t));
	for (uint32_t t = 0; t < n_terms; t++) {
		size_t h = fnv1a(ix->blob + term_off[t], term_off[t + 1] - term_off[t])
		    & (ix->hash_n - 1);
		while (ix->hash_slots[h])
			h = (h + 1) & (ix->hash_n - 1);
		ix->hash_slots[h] = t + 1;
	}
	return ix;
}

static void
bk_free(bknode_t *n)
{
	if (!n)
		return;
	for (unsigned i = 0; i < n->nchild; i++)
		bk_free(n->child[i]);
	free(n->edge);
	free(n->child);
	free(n);
}

void
ora_index_free(ora_index_t *ix)
{
	if (!ix)
		return;
	bk_free(ix->bk_root);
	free(ix->doc_ids); free(ix->doc_len);
	free(ix->post_off); free(ix->post_doc); free(ix->post_tf);
	free(ix->blob); free(ix->term_off); free(ix->term_total);
	free(ix->hash_slots);
	free(ix);
}

uint32_t
ora_term_df(const ora_index_t *ix, uint32_t term_id)
{
	if (term_id == 0 || term_id > ix->n_terms)
		return 0;
	return ix->post_off[term_id] - ix->post_off[term_id - 1];
}

uint32_t
ora_term_lookup(ora_index_t *ix, const char *s, size_t len)
{
	size_t h = fnv1a(s, len) & (ix->hash_n - 1);

	while (ix->hash_slots[h]) {
		const uint32_t t = ix->hash_slots[h] - 1;

		if (ix->term_off[t + 1] - ix->term_off[t] == len &&
		    memcmp(ix->blob + ix->term_off[t], s, len) == 0)
			return t + 1;
		h = (h + 1) & (ix->hash_n - 1);
	}
	return 0;
}

/*
 * Ranking: ref algo/ranking.c.
 */

float
ora_score_pair(const ora_index_t *ix, int algo, uint32_t term_id,
    uint32_t tf_u, uint32_t doclen_u)
{
	/* ranking.c:72-78,144-150: the operand types matter below. */
	const int term_freq = (int)tf_u;
	const unsigned long doc_count = ix->doc_count;
	const unsigned long doc_freq = ora_term_df(ix, term_id);

	if (term_freq <= 0 || doc_count == 0)		/* :86,156 */
		return -1;

	if (algo == ORA_TFIDF) {
		/* ranking.c:90-96: float tf, float idf from a FLOAT division. */
		float tf = log(term_freq + 1);
		float idf = log((float)doc_count / doc_freq) + 1;
		return tf * idf;
	} else {
		/* ranking.c:141-142: the constants are float literals widened. */
		static const double k = 1.2f;
		static const double b = 0.75f;
		double tf, dl, adl, tf_bm25, idf_bm25;

		/* ranking.c:163: INTEGER division of u64 by unsigned long. */
		adl = ix->token_count / doc_count;
		if (adl < 1)
			return -1;
		tf = log(term_freq + 1);
		dl = (int)doclen_u;			/* idxdoc.c:78-94 */
		tf_bm25 = tf / (tf + k * (1 - b + b * dl / adl));
		idf_bm25 = log(((doc_count - doc_freq + 0.5) /
		    (doc_freq + 0.5)) + 1);
		return tf_bm25 * idf_bm25;
	}
}

/*
 * Top-N: a capped binary min-heap followed by an in-place heapsort
 * (ref algo/heap.c).  Tie behaviour is an artefact of exactly these sift
 * rules, so they are restated one for one: sift-up stops at "not smaller
 * than the parent"; sift-down prefers the left child unless the right one is
 * strictly smaller; at the cap an item not strictly greater than the root is
 * dropped.
 */

typedef struct { float s; uint64_t p; } hitem_t;

static void
heap_sift_down(hitem_t *h, size_t n)
{
	size_t i = 0;

	for (;;) {
		const size_t l = 2 * i + 1, r = l + 1;
		size_t m = i;

		if (l >= n)
			break;
		if (h[l].s < h[i].s)
			m = l;
		if (r < n && h[r].s < h[m].s)
			m = r;
		if (m == i)
			break;
		hitem_t t = h[i]; h[i] = h[m]; h[m] = t;
		i = m;
	}
}

static hitem_t
heap_pop_min(hitem_t *h, size_t *n)
{
	const hitem_t top = h[0];

	if (--*n) {
		h[0] = h[*n];
		heap_sift_down(h, *n);
	}
	return top;
}

static void
heap_push(hitem_t *h, size_t *n, size_t cap, hitem_t it)
{
	size_t i;

	if (*n == cap) {
		if (!(it.s > h[0].s))		/* heap.c:68-75: cmp <= 0 drops */
			return;
		heap_pop_min(h, n);
	}
	i = (*n)++;
	h[i] = it;
	while (i) {
		const size_t p = (i - 1) / 2;

		if (!(h[i].s < h[p].s))
			break;
		hitem_t t = h[i]; h[i] = h[p]; h[p] = t;
		i = p;
	}
}

size_t
ora_heap_topn(size_t limit, size_t n, const float *scores,
    const uint64_t *payload, uint64_t *out_payload, float *out_scores)
{
	const size_t cap = limit < n ? limit : n;
	hitem_t *h = malloc(sizeof(hitem_t) * (cap + 1));
	size_t cnt = 0, total;

	if (cap == 0) {
		free(h);
		return 0;
	}
	for (size_t i = 0; i < n; i++)
		heap_push(h, &cnt, cap, (hitem_t){ scores[i], payload[i] });

	/* heap.c:196-221: repeatedly move the minimum to the shrinking tail. */
	total = cnt;
	while (cnt) {
		const size_t last = cnt - 1;
		const hitem_t m = heap_pop_min(h, &cnt);

		out_scores[last] = m.s;
		out_payload[last] = m.p;
	}
	free(h);
	return total;
}

/*
 * Boolean logic + scoring: ref query/search.c:118-278.
 */

typedef uint64_t word_t;
#define	WBITS	64

static word_t *
term_bitset(const ora_index_t *ix, uint32_t term_id, size_t nw)
{
	word_t *bs = calloc(nw ? nw : 1, sizeof(word_t));

	if (term_id && term_id <= ix->n_terms) {
		for (uint64_t j = ix->post_off[term_id - 1];
		    j < ix->post_off[term_id]; j++) {
			const uint32_t d = ix->post_doc[j];
			bs[d / WBITS] |= (word_t)1 << (d % WBITS);
		}
	}
	return bs;
}

/* get_expr_bitmap (search.c:118-174) on a postfix form of the same tree. */
static word_t *
eval_program(const ora_index_t *ix, uint32_t n_tokens,
    const uint32_t *token_terms, uint32_t n_prog, const int32_t *prog,
    size_t nw)
{
	word_t **stack = calloc((size_t)n_prog + 1, sizeof(word_t *));
	size_t sp = 0;
	word_t *res = NULL;

	for (uint32_t i = 0; i < n_prog; i++) {
		const int32_t op = prog[i];

		if (op >= 0) {
			if ((uint32_t)op >= n_tokens)
				goto bad;
			stack[sp++] = term_bitset(ix, token_terms[op], nw);
		} else if (op == ORA_OP_EMPTY) {
			stack[sp++] = term_bitset(ix, 0, nw);
		} else {
			word_t *a, *b;

			if (sp < 2)
				goto bad;
			b = stack[--sp];
			a = stack[sp - 1];
			for (size_t w = 0; w < nw; w++) {
				switch (op) {
				case ORA_OP_AND:	a[w] &= b[w]; break;
				case ORA_OP_OR:		a[w] |= b[w]; break;
				case ORA_OP_ANDNOT:	a[w] &= ~b[w]; break;
				default:		free(b); goto bad;
				}
			}
			free(b);
		}
	}
	if (sp == 1)
		res = stack[--sp];
bad:
	while (sp)
		free(stack[--sp]);
	free(stack);
	return res;
}

/*
 * Scores of all matching documents, ascending id (run_query_logic,
 * search.c:210-278, with nxs_resp_addresult's float accumulation,
 * results.c:128-151).  Returns arrays the caller frees.
 */
static int64_t
score_matches(const ora_index_t *ix, int algo, uint32_t n_tokens,
    const uint32_t *token_terms, uint32_t n_prog, const int32_t *prog,
    uint32_t **docs_out, float **scores_out)
{
	const size_t nw = ((size_t)ix->n_docs + WBITS - 1) / WBITS;
	uint64_t *cur = calloc((size_t)n_tokens + 1, sizeof(uint64_t));
	uint32_t *docs = NULL;
	float *scores = NULL;
	size_t n = 0, cap = 0;
	word_t *match;

	*docs_out = NULL;
	*scores_out = NULL;

	/* search.c:224-226: nothing resolved => empty result, no error. */
	if (n_prog == 0 || n_tokens == 0) {
		free(cur);
		return 0;
	}
	if ((match = eval_program(ix, n_tokens, token_terms, n_prog, prog, nw))
	    == NULL) {
		free(cur);
		return -1;
	}
	for (uint32_t t = 0; t < n_tokens; t++)
		cur[t] = ix->post_off[token_terms[t] - 1];

	for (size_t w = 0; w < nw; w++) {
		word_t bits = match[w];

		while (bits) {
			const uint32_t d = w * WBITS + __builtin_ctzll(bits);
			bool have = false;
			float acc = 0;

			bits &= bits - 1;
			/* EVERY resolved token, token-list order (search.c:239). */
			for (uint32_t t = 0; t < n_tokens; t++) {
				const uint32_t term = token_terms[t];
				const uint64_t end = ix->post_off[term];
				float s;

				while (cur[t] < end && ix->post_doc[cur[t]] < d)
					cur[t]++;
				if (cur[t] == end || ix->post_doc[cur[t]] != d)
					continue;	/* search.c:250-253 */
				s = ora_score_pair(ix, algo, term,
				    ix->post_tf[cur[t]], ix->doc_len[d]);
				if (s < 0)
					continue;	/* search.c:261-266 */
				if (!have) {
					acc = s;	/* results.c:141-147 */
					have = true;
				} else {
					acc += s;	/* results.c:135-137 */
				}
			}
			if (!have)
				continue;
			if (n == cap) {
				cap = cap ? cap * 2 : 1024;
				docs = realloc(docs, sizeof(uint32_t) * cap);
				scores = realloc(scores, sizeof(float) * cap);
			}
			docs[n] = d;
			scores[n] = acc;
			n++;
		}
	}
	free(match);
	free(cur);
	*docs_out = docs;
	*scores_out = scores;
	return n;
}

int64_t
ora_search_all(const ora_index_t *ix, int algo, uint32_t n_tokens,
    const uint32_t *token_terms, uint32_t n_prog, const int32_t *prog,
    uint64_t *out_ids, float *out_scores, size_t cap)
{
	uint32_t *docs;
	float *scores;
	const int64_t n = score_matches(ix, algo, n_tokens, token_terms,
	    n_prog, prog, &docs, &scores);

	for (int64_t i = 0; i < n && (size_t)i < cap; i++) {
		out_ids[i] = ix->doc_ids[docs[i]];
		out_scores[i] = scores[i];
	}
	free(docs);
	free(scores);
	return n;
}

int64_t
ora_search(const ora_index_t *ix, int algo, uint64_t limit, uint32_t n_tokens,
    const uint32_t *token_terms, uint32_t n_prog, const int32_t *prog,
    uint64_t *out_ids, float *out_scores, size_t cap)
{
	uint32_t *docs;
	float *scores, *fs, *os;
	uint64_t *fp, *op;
	size_t kept;
	const int64_t n = score_matches(ix, algo, n_tokens, token_terms,
	    n_prog, prog, &docs, &scores);

	if (n <= 0) {
		free(docs);
		free(scores);
		return n;
	}
	/*
	 * results.c:145-147 head-inserts each new entry, so nxs_resp_build
	 * (results.c:190-195) feeds the heap in DESCENDING document order.
	 */
	fs = malloc(sizeof(float) * n);
	fp = malloc(sizeof(uint64_t) * n);
	os = malloc(sizeof(float) * n);
	op = malloc(sizeof(uint64_t) * n);
	for (int64_t i = 0; i < n; i++) {
		fs[i] = scores[n - 1 - i];
		fp[i] = ix->doc_ids[docs[n - 1 - i]];
	}
	kept = ora_heap_topn(limit, n, fs, fp, op, os);
	for (size_t i = 0; i < kept && i < cap; i++) {
		out_ids[i] = op[i];
		out_scores[i] = os[i];
	}
	free(fs); free(fp); free(os); free(op);
	free(docs);
	free(scores);
	return kept;
}

/*
 * Levenshtein distance: ref algo/levdist.c:67-150 (Wagner-Fischer, one row,
 * bytes, unit costs).
 */
int
ora_levdist(const char *a, size_t n, const char *b, size_t m)
{
	uint16_t stackrow[128], *row = stackrow;
	int d;

	if (n < m) {
		const char *ts = a; a = b; b = ts;
		size_t tn = n; n = m; m = tn;
	}
	if (m == 0)
		return (int)n;
	if (m + 1 > sizeof(stackrow) / sizeof(stackrow[0]))
		row = malloc(sizeof(uint16_t) * (m + 1));

	for (size_t j = 0; j <= m; j++)
		row[j] = j;
	for (size_t i = 0; i < n; i++) {
		unsigned diag = i;	/* D[i][0] */

		row[0] = i + 1;
		for (size_t j = 1; j <= m; j++) {
			const unsigned up = row[j];
			unsigned best = diag + (a[i] != b[j - 1]);

			if (up + 1 < best)
				best = up + 1;
			if ((unsigned)row[j - 1] + 1 < best)
				best = row[j - 1] + 1;
			row[j] = best;
			diag = up;
		}
	}
	d = row[m];
	if (row != stackrow)
		free(row);
	return d;
}

/*
 * BK-tree: ref algo/bktree.c.  Insert order = term-id order
 * (index/terms.c:404-405, index/idxterm.c:171).
 */

static inline const char *
term_str(const ora_index_t *ix, uint32_t t, size_t *len)
{
	*len = ix->term_off[t + 1] - ix->term_off[t];
	return ix->blob + ix->term_off[t];
}

static bknode_t *
bk_child(const bknode_t *n, unsigned edge)
{
	for (unsigned i = 0; i < n->nchild; i++) {
		if (n->edge[i] == edge)
			return n->child[i];
	}
	return NULL;
}

static void
bk_attach(bknode_t *n, unsigned edge, bknode_t *c)
{
	unsigned pos = 0;

	n->edge = realloc(n->edge, n->nchild + 1);
	n->child = realloc(n->child, sizeof(bknode_t *) * (n->nchild + 1));
	while (pos < n->nchild && n->edge[pos] < edge)
		pos++;
	memmove(n->edge + pos + 1, n->edge + pos, n->nchild - pos);
	memmove(n->child + pos + 1, n->child + pos,
	    sizeof(bknode_t *) * (n->nchild - pos));
	n->edge[pos] = edge;
	n->child[pos] = c;
	n->nchild++;
}

static void
bk_build(ora_index_t *ix)
{
	for (uint32_t t = 0; t < ix->n_terms; t++) {
		bknode_t *nn = calloc(1, sizeof(*nn)), *cur = ix->bk_root;
		size_t tl, cl;
		const char *ts = term_str(ix, t, &tl);

		nn->term = t;
		if (!cur) {
			ix->bk_root = nn;
			continue;
		}
		for (;;) {
			const char *cs = term_str(ix, cur->term, &cl);
			int d = ora_levdist(ts, tl, cs, cl);
			bknode_t *next;

			if (d <= 0) {		/* bktree.c:183-190: duplicate */
				free(nn);
				break;
			}
			if (d > BK_EDGE_MAX)	/* bktree.c:196 */
				d = BK_EDGE_MAX;
			if ((next = bk_child(cur, d)) == NULL) {
				bk_attach(cur, d, nn);
				break;
			}
			cur = next;
		}
	}
}

uint32_t
ora_fuzzy(ora_index_t *ix, const char *q, size_t len, uint32_t *cands,
    uint32_t *dists, size_t cap, size_t *n_cands, size_t *n_visited)
{
	bknode_t **queue;
	size_t head = 0, tail = 0, nc = 0;
	uint32_t chosen = 0;

	if (!ix->bk_root && ix->n_terms)
		bk_build(ix);
	if (n_cands) *n_cands = 0;
	if (n_visited) *n_visited = 0;
	if (!ix->bk_root)
		return 0;

	queue = malloc(sizeof(bknode_t *) * ((size_t)ix->n_terms + 1));
	queue[tail++] = ix->bk_root;

	/* bktree.c:240-271: breadth-first, children by ascending edge label. */
	while (head < tail) {
		const bknode_t *n = queue[head++];
		size_t tl;
		const char *ts = term_str(ix, n->term, &tl);
		const int d = ora_levdist(q, len, ts, tl);
		int lo, hi;

		if (d <= LEV_TOLERANCE) {
			/*
			 * idxterm.c:238-242 pops the candidates back to front and
			 * keeps overwriting `term` whenever total > term_total,
			 * with term_total never updated from 0: the survivor is the
			 * EARLIEST pushed candidate whose on-disk total is > 0.
			 */
			if (!chosen && ix->term_total[n->term] > 0)
				chosen = n->term + 1;
			if (cands && nc < cap) {
				cands[nc] = n->term + 1;
				if (dists)
					dists[nc] = d;
			}
			nc++;
		}
		/*
		 * bktree.c:151-157,258-264: children with edge label in
		 * [max(d - tol, 0), min(d + tol, 63)) -- upper bound EXCLUSIVE.
		 */
		lo = d - LEV_TOLERANCE < 0 ? 0 : d - LEV_TOLERANCE;
		hi = d + LEV_TOLERANCE > BK_EDGE_MAX ? BK_EDGE_MAX : d + LEV_TOLERANCE;
		for (unsigned i = 0; i < n->nchild; i++) {
			if (n->edge[i] >= lo && n->edge[i] < hi)
				queue[tail++] = n->child[i];
		}
	}
	if (n_cands) *n_cands = nc;
	if (n_visited) *n_visited = head;
	free(queue);
	return chosen;
}

/*
 * Every vocabulary term within LEV_TOLERANCE of q, by brute force over the
 * whole vocabulary (no tree): the "true <= 2 set" SURVEY 8a F3 asks to be
 * reported next to the BK-tree's pruned candidate list (the half-open child
 * range of ref src/algo/bktree.c:151-157 makes that list a subset).  Terms in
 * id order; returns the count (entries past cap are counted, not stored).
 */
size_t
ora_fuzzy_true(ora_index_t *ix, const char *q, size_t len, uint32_t *terms,
    uint32_t *dists, size_t cap)
{
	size_t n = 0;

	for (uint32_t t = 0; t < ix->n_terms; t++) {
		size_t tl;
		const char *ts = term_str(ix, t, &tl);
		int d;

		if (tl > len + LEV_TOLERANCE || len > tl + LEV_TOLERANCE)
			continue;
		d = ora_levdist(q, len, ts, tl);
		if (d > LEV_TOLERANCE)
			continue;
		if (n < cap) {
			if (terms)
				terms[n] = t + 1;
			if (dists)
				dists[n] = (uint32_t)d;
		}
		n++;
	}
	return n;
}
