/*
 * Introspection of the query front end for the tests and benchmarks
 * (include/nxsb200_tools.h): built into libnxsb_tools.so, NOT into the product
 * library -- libnxsearch.so carries the search path only.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "nxs.h"
#include "nxsb200_gpu.h"
#include "nxsb200_tools.h"
#include "nxs_impl.h"
#include "query.h"
#include "tokenizer.h"

/*
 * The tokenizer reports filter errors through the instance's error slot
 * (nxs.c); there is no instance here and the tools run without filters.
 */
void
nxs_set_error(nxs_t *nxs, nxs_err_t code, const char *fmt, ...)
{
	(void)nxs; (void)code; (void)fmt;
}

void
nxs_set_syserror(nxs_t *nxs, nxs_err_t code, const char *fmt, ...)
{
	(void)nxs; (void)code; (void)fmt;
}

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
	tokenset_t *ts = NULL;
	int32_t *stack = NULL;
	qtree_t tree;
	size_t off = 0;
	int32_t sp = 0;
	int ret = -1;

	/* No engine limits here: the general token set, same walk as prepare_query(). */
	*n_tokens = *n_prog = 0;
	qtree_parse(&tree, query);
	if (tree.error || (ts = tokenset_create()) == NULL ||
	    (stack = malloc(sizeof(int32_t) * (tree.n_nodes + 2))) == NULL)
		goto out;
	if (tree.root >= 0)
		stack[sp++] = tree.root;
	while (sp) {
		qnode_t *n = &tree.nodes[stack[--sp]];

		if (n->type != QN_VALUE) {
			stack[sp++] = n->left;
			stack[sp++] = n->right;
		} else if (tokenize_value(&nofilters, ts, n->value, strlen(n->value),
		    &n->token) == -1) {
			goto out;
		}
	}
	for (uint32_t j = 0; j < ts->count; j++) {
		const token_t *t = &ts->list[j];

		if (off + t->len + 1 > buf_len)
			goto out;
		memcpy(tokens_buf + off, t->str, t->len + 1);
		off += t->len + 1;
	}
	if (tree.root >= 0) {
		if ((uint32_t)tree.n_nodes > prog_cap || tree.depth > NXS_QUERY_RLIMIT)
			goto out;
		qtree_emit_program(&tree, tree.root, prog, n_prog);
	}
	*n_tokens = ts->count;
	ret = 0;
out:
	free(stack);
	tokenset_destroy(ts);
	qtree_free(&tree);
	return ret;
}

NXS_API int
nxsb_tokenize(const char *text, size_t len, int normalize, char *tokens_buf,
    size_t buf_len, uint32_t *n_tokens, uint32_t *counts, uint32_t counts_cap)
{
	filter_pipeline_t fp = { 0 };
	tokenset_t *ts;
	size_t off = 0;
	int ret = -1;

	*n_tokens = 0;
	if (normalize & 1)
		fp.kinds[fp.count++] = FILT_NORMALIZER;
	if (normalize & 2) {
		fp.kinds[fp.count++] = FILT_STEMMER;
		fp.stem_english = true;
	}
	if ((ts = tokenize(&fp, text, len)) == NULL)
		return -1;
	for (uint32_t j = 0; j < ts->count; j++) {
		const token_t *t = &ts->list[j];

		if (off + t->len + 1 > buf_len || (counts && j >= counts_cap))
			goto out;
		memcpy(tokens_buf + off, t->str, t->len);
		tokens_buf[off + t->len] = '\0';
		off += t->len + 1;
		if (counts)
			counts[j] = t->count;
	}
	*n_tokens = ts->count;
	ret = 0;
out:
	tokenset_destroy(ts);
	return ret;
}
