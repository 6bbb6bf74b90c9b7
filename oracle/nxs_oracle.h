/*
 * ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar CPU restatement of the reference's query-scoring path, used only
 * as the checker in tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  Nothing under nxsearch_b200/ may
 * include, link or call it.
 *
 * Parity of this restatement is PINNED: tests/test_oracle_*.py check it
 * against the reference's own golden vectors (src/tests/t_scoring.c,
 * t_levdist.c, t_bktree.c, t_heap.c, t_querylogic.c) and, where
 * oracle/_ref/libnxsearch_ref.so exists, against the reference itself run on
 * the same synthetic index.
 */
#ifndef NXS_ORACLE_H
#define NXS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_index ora_index_t;

enum { ORA_TFIDF = 0, ORA_BM25 = 1 };		/* ref index.h:30-34 */

/* Postfix boolean program over token slots (shared with the product). */
enum {
	ORA_OP_EMPTY	= -1,	/* push the empty set (unresolved leaf) */
	ORA_OP_AND	= -2,
	ORA_OP_OR	= -3,
	ORA_OP_ANDNOT	= -4,
	/* >= 0: push the document set of token slot i */
};

/*
 * Build from a document-major corpus (same arrays as nxsb_corpus_t).
 * doc_count / token_count are the dtmap header counters the reference's
 * ranking functions read (ranking.c:77,149,163).
 */
ora_index_t *	ora_index_build(uint32_t n_docs, const uint64_t *doc_ids,
		    const uint32_t *doc_len, const uint64_t *doc_off,
		    const uint32_t *pairs, uint32_t n_terms,
		    const char *term_blob, const uint32_t *term_off,
		    const uint64_t *term_total, uint64_t token_count,
		    uint32_t doc_count);
void		ora_index_free(ora_index_t *);
uint32_t	ora_term_df(const ora_index_t *, uint32_t term_id);

/*
 * One (term, doc) score, as ranking.c:41-176.  doc is a dense index in
 * ascending external-id order.  Returns < 0 for "skip".
 */
float		ora_score_pair(const ora_index_t *, int algo, uint32_t term_id,
		    uint32_t tf, uint32_t doclen);

/*
 * The whole of run_query_logic + nxs_resp_build (search.c:210-278,
 * results.c:128-220): evaluate the boolean program, score every resolved
 * token of the query (token-list order) on every matching document
 * (ascending id), then cap with the reference's min-heap and heapsort.
 * token_terms[i] is the 1-based term id of token slot i.  Returns the result
 * count (<= limit, <= cap) or -1 on a malformed program.
 */
int64_t		ora_search(const ora_index_t *, int algo, uint64_t limit,
		    uint32_t n_tokens, const uint32_t *token_terms,
		    uint32_t n_prog, const int32_t *prog,
		    uint64_t *out_ids, float *out_scores, size_t cap);

/*
 * As ora_search but returns EVERY matching document (ascending id) with its
 * accumulated score and no cap: the ground truth the tie-aware comparison
 * in tests needs (SURVEY 8a "F5 ties").  Returns the match count; fills at
 * most cap entries.
 */
int64_t		ora_search_all(const ora_index_t *, int algo,
		    uint32_t n_tokens, const uint32_t *token_terms,
		    uint32_t n_prog, const int32_t *prog,
		    uint64_t *out_ids, float *out_scores, size_t cap);

/* Capped min-heap + heapsort of heap.c:59-221 on (score, payload) items fed
 * in the given order; writes the final order.  Returns the kept count. */
size_t		ora_heap_topn(size_t limit, size_t n, const float *scores,
		    const uint64_t *payload, uint64_t *out_payload,
		    float *out_scores);

/* Byte-wise Levenshtein distance, unit costs (levdist.c:67-150). */
int		ora_levdist(const char *a, size_t n, const char *b, size_t m);

/*
 * Fuzzy resolution of one query token (idxterm.c:210-249 over the BK-tree of
 * bktree.c:160-275 built in term-id order, tolerance 2).  Returns the chosen
 * term id, or 0.  If cands != NULL the candidate term ids are written in the
 * order the reference's BFS pushes them (at most cap), *n_cands gets their
 * number and *n_visited the number of tree nodes whose distance was computed.
 */
uint32_t	ora_fuzzy(ora_index_t *, const char *q, size_t len,
		    uint32_t *cands, uint32_t *dists, size_t cap,
		    size_t *n_cands, size_t *n_visited);

/* Exact lookup of a term string: term id or 0 (idxterm.c:192-196). */
uint32_t	ora_term_lookup(ora_index_t *, const char *s, size_t len);

#ifdef __cplusplus
}
#endif

#endif
