/*
 * nxsearch-b200: bulk tooling around the index files (additive; the
 * reference has no counterpart beyond nxs_index_add()).
 *
 *  - the deterministic synthetic corpus / query generator of SURVEY.md
 *    section 8(d) (Zipf(1.0) vocabulary, seed 0x6E78735F42323030);
 *  - a bulk writer that emits `nxsterms` / `nxsdtmap` files in the
 *    reference's on-disk format (reference: src/index/storage.h:12-133;
 *    golden bytes: src/tests/t_index_terms.c:23-37,
 *    src/tests/t_index_dtmap.c:25-41), so that this engine and the
 *    reference open byte-identical files.
 *
 * Everything here is host-only integer work; no CUDA.
 */
#ifndef NXSB200_TOOLS_H
#define NXSB200_TOOLS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Everything declared here is exported; the rest of the library is hidden. */
#pragma GCC visibility push(default)

#define NXSB_CORPUS_SEED	UINT64_C(0x6E78735F42323030)	/* "nxs_B200" */

/*
 * A corpus in document-major form, host memory, native endianness.
 * This is the in-memory twin of an `nxsdtmap` file: for document i
 * (0-based, in file order) pairs[2*j], pairs[2*j+1] for j in
 * [doc_off[i], doc_off[i+1]) are (term id, count), ascending term id.
 */
typedef struct nxsb_corpus {
	uint32_t	n_docs;
	uint32_t	n_terms;	/* vocabulary size; term ids 1..n_terms */
	uint64_t	n_pairs;
	uint64_t	token_count;	/* sum of doc_len (file: header field) */
	uint32_t	doc_count;	/* N used for scoring (file: header field) */
	uint64_t *	doc_ids;	/* [n_docs] external ids (non-zero) */
	uint32_t *	doc_len;	/* [n_docs] tokens incl. repeats */
	uint64_t *	doc_off;	/* [n_docs + 1] */
	uint32_t *	pairs;		/* [2 * n_pairs] */
	/* Vocabulary: term i (id i+1) = term_blob[term_off[i]..term_off[i+1]). */
	char *		term_blob;
	uint32_t *	term_off;	/* [n_terms + 1] */
	uint64_t *	term_total;	/* [n_terms] occurrences in the corpus */
	uint32_t *	term_df;	/* [n_terms] documents containing it */
} nxsb_corpus_t;

/*
 * Generate the SURVEY section 8(d) corpus: V distinct [a-z0-9]{4,12} terms
 * (term id = Zipf rank), documents of 16..111 tokens drawn Zipf(1.0), ids
 * 1..n_docs (or, with sparse_ids, a strictly increasing pseudo-random 64-bit
 * sequence).  first_doc/n_docs select a slice of the same global corpus, so
 * shards can be generated independently; the vocabulary is always the whole.
 * nthreads <= 0 means "all online CPUs".  NULL on allocation failure.
 */
nxsb_corpus_t *	nxsb_corpus_generate(uint64_t seed, uint32_t n_terms,
		    uint64_t first_doc, uint32_t n_docs, int sparse_ids,
		    int nthreads);
void		nxsb_corpus_free(nxsb_corpus_t *);

/*
 * Synthetic query terms: fills term_ids[0..n) with Zipf(1.0) draws over the
 * vocabulary restricted to terms with df[t] >= 1 (SURVEY 8d "Queries");
 * df may be NULL (no restriction).  Deterministic in (seed, n_terms).
 */
void		nxsb_corpus_query_terms(uint64_t seed, uint32_t n_terms,
		    const uint32_t *df, uint32_t *term_ids, size_t n);

/*
 * Fuzzy workload (SURVEY 8d "C4"): n query strings, each a random
 * vocabulary term with 1-2 random byte edits (substitute / insert / delete
 * over [a-z0-9]), re-drawn if it equals a vocabulary term.  Strings are
 * written NUL-terminated into out (stride bytes apart, stride >= 16).
 */
void		nxsb_corpus_fuzzy_terms(uint64_t seed, const nxsb_corpus_t *,
		    char *out, size_t stride, size_t n);

/*
 * Write the corpus as reference-format index files (big-endian, 32 KB file
 * granularity as the reference's idx_db_map, src/index/idxmap.c:119-177).
 * Returns 0, or -1 with errno set.
 */
int		nxsb_write_terms_file(const char *path, const nxsb_corpus_t *);
int		nxsb_write_dtmap_file(const char *path, const nxsb_corpus_t *);

/*
 * Read reference-format index files back into a corpus (doc_ids in file
 * order; deleted blocks and deletion markers applied as the reference's
 * idx_dtmap_sync does, src/index/dtmap.c:357-384,440-544).  term_df is
 * recomputed; term_total is taken from the file.  NULL + errno on error
 * (EINVAL for a corrupt file).
 */
nxsb_corpus_t *	nxsb_read_index_files(const char *terms_path,
		    const char *dtmap_path);

/*
 * Flat mirror of the BK-tree the reference would build over a vocabulary by
 * inserting the terms in id order (ref src/algo/bktree.c:160-217): for term
 * index t, parent[t] (UINT32_MAX for the root), edge[t] = min(distance to
 * the parent, 63) and rank[t] = position in a full breadth-first walk with
 * children by ascending edge label.  This is the input of
 * nxsb_engine_load_vocab().  Returns 0, or -1 on allocation failure.
 */
int		nxsb_bkmirror_build(const char *term_blob, const uint32_t *term_off,
		    uint32_t n_terms, uint32_t *parent, uint8_t *edge,
		    uint32_t *rank);

/*
 * Query-language introspection (the parser itself is internal).  Used by the
 * tests that pin the lexer / grammar to the reference's golden cases
 * (ref src/tests/t_queryparser.c:27-115).
 *
 * nxsb_query_lex:     token kinds in order (1 OR, 2 AND, 3 NOT, 4 '(', 5 ')',
 *                     6 free-form string, 7 quoted string); returns the count.
 * nxsb_query_dump:    the parse tree as "(AND (OR `A` `B`) `C`)", malloc'ed;
 *                     NULL on a syntax error, *errmsg then holds a malloc'ed
 *                     "syntax error near L:C: ..." message.
 * nxsb_query_compile: what the search path hands to the GPU engine, without
 *                     filters or term resolution: the distinct leaf strings in
 *                     token-list order (NUL-separated into tokens_buf) and the
 *                     postfix program over their slots.  Returns 0, or -1 on a
 *                     syntax error / insufficient capacity.
 */
size_t		nxsb_query_lex(const char *query, int *kinds, size_t cap);
char *		nxsb_query_dump(const char *query, char **errmsg);
int		nxsb_query_compile(const char *query, char *tokens_buf,
		    size_t buf_len, uint32_t *n_tokens, int32_t *prog,
		    uint32_t prog_cap, uint32_t *n_prog);

/*
 * The text front end on its own (the tokenizer is internal): the distinct
 * words of `text` in first-seen order, NUL-separated into tokens_buf, with
 * their occurrence counts -- what nxs_index_add hands to the term table
 * (ref src/core/tokenizer.c:234-302, token sets :94-117).  Bit 0 of
 * `normalize' runs the "normalizer" filter over every word, bit 1 the
 * English "stemmer" after it.  Used by the tests that pin the segmentation
 * to the reference's golden cases (ref src/tests/t_tokenize.c:17-62,
 * src/tests/t_utf8.c:70-74) and the stemmer to the published vocabulary.
 * Returns 0, or -1 on insufficient capacity.
 */
int		nxsb_tokenize(const char *text, size_t len, int normalize,
		    char *tokens_buf, size_t buf_len, uint32_t *n_tokens,
		    uint32_t *counts, uint32_t counts_cap);

/*
 * Drain n responses of a batch the way a C caller would -- with the public
 * iterator nxs_resp_iter_reset() / nxs_resp_iter_result() (ref
 * src/core/results.c:222-247) -- into flat arrays: counts[i] results of
 * response i at ids/scores[i * stride ...].  NULL responses count as empty.
 * Returns the total number of results copied.  (Opaque pointers: nxs_resp_t.)
 */
uint64_t	nxsb_resp_collect(void *const *resps, size_t n, uint32_t stride,
		    uint32_t *counts, uint64_t *ids, float *scores);

/*
 * How the HBM image of an open index (opaque pointer: nxs_index_t) reached its
 * current state: out[0] full builds, out[1] delta segments added, out[2]
 * consolidations of the delta segments, out[3] delta segments present now,
 * out[4] removed documents noted against a segment, out[5] live documents,
 * out[6] documents appended but not on the GPU yet.  (Incremental refresh,
 * SURVEY 8f N1; the counterpart of the reference folding appended dtmap
 * blocks into its in-memory index, ref src/index/dtmap.c:357-441.)
 */
void		nxsb_index_image_stats(const void *index, uint64_t out[7]);

/*
 * Live documents containing term_id as the open index (opaque pointer:
 * nxs_index_t) counts them -- the cardinality of the term's document set in
 * the reference (ref src/algo/ranking.c:78,150).  0 for an unknown id.
 */
uint32_t	nxsb_index_term_df(const void *index, uint32_t term_id);

#pragma GCC visibility pop

#ifdef __cplusplus
}
#endif

#endif
