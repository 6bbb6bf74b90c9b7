/*
 * ORACLE SHIM (test infrastructure, not product code).
 *
 * The reference generates its lexer with re2c (query/scan.re) and its
 * parser with lemon (query/grammar.y); neither tool nor the generated C
 * files exist in this image.  This file supplies hand-written lex_init(),
 * lex() and query_parse() with the behaviour those two sources specify,
 * against the reference's own query_t / expr_t (query/query.h, expr.h):
 *
 *  - lexer: longest match, earlier rule wins a tie (scan.re:64-119):
 *    whitespace, '&'|AND, '|'|OR, NOT (case-insensitive), parentheses,
 *    quoted strings with backslash escapes kept verbatim, free-form strings
 *    = any run of bytes other than NUL, whitespace, '(' and ')'.
 *  - grammar (grammar.y:62-110): OR < AND < NOT, all left-associative;
 *    "AND NOT" is one binary operator; juxtaposition at the TOP LEVEL only
 *    is an implicit OR; parentheses hold a single expr.
 *
 * Pinned by the cases of tests/t_queryparser.c:27-115 (see
 * tests/test_oracle_ref.py).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#define __NXSLIB_PRIVATE
#define __NXS_PARSER_PRIVATE
#include "nxs_impl.h"
#include "expr.h"
#include "query.h"
#include "grammar.h"

static int
is_sp(unsigned char c)
{
	return c == ' ' || c == '\t' || c == '\v' || c == '\f' ||
	    c == '\r' || c == '\n';
}

static int
is_ff(unsigned char c)
{
	return c != 0 && !is_sp(c) && c != '(' && c != ')';
}

void
lex_init(lexer_t *ctx, const char *s)
{
	ctx->cursor = s;
	ctx->cur_line = s;
	ctx->line = 1;
}

/* Length of a well-formed quoted string starting at s, or 0. */
static size_t
quoted_len(const char *s)
{
	const char q = s[0];
	size_t i = 1;

	for (;;) {
		const char c = s[i];

		if (c == '\0')
			return 0;
		if (c == '\\') {
			if (s[i + 1] == '\0')
				return 0;
			i += 2;
			continue;
		}
		i++;
		if (c == q)
			return i;
	}
}

int
lex(query_t *q)
{
	lexer_t *ctx = &q->lexer;
	lexval_t *lval = &q->lval;

	for (;;) {
		const char *s = ctx->cursor;
		const unsigned char c = (unsigned char)*s;
		size_t ff, oplen = 0, qlen = 0;
		int op = 0;

		ctx->token = s;
		if (c == '\0') {
			return 0;
		}
		if (is_sp(c)) {
			size_t n = 1;

			while (is_sp((unsigned char)s[n]))
				n++;
			if (c == '\n' && n == 1) {
				/* EOL rule wins the tie with WSP. */
				ctx->cur_line = ctx->token;
				ctx->line++;
			}
			ctx->cursor = s + n;
			continue;
		}
		if (c == '(') {
			ctx->cursor = s + 1;
			return TOKEN_BR_OPEN;
		}
		if (c == ')') {
			ctx->cursor = s + 1;
			return TOKEN_BR_CLOSE;
		}

		/* Free-form run length; every other candidate competes with it. */
		for (ff = 1; is_ff((unsigned char)s[ff]); ff++)
			;

		if (c == '&') {
			op = TOKEN_AND, oplen = 1;
		} else if (c == '|') {
			op = TOKEN_OR, oplen = 1;
		} else if (strncasecmp(s, "AND", 3) == 0) {
			op = TOKEN_AND, oplen = 3;
		} else if (strncasecmp(s, "NOT", 3) == 0) {
			op = TOKEN_NOT, oplen = 3;
		} else if (strncasecmp(s, "OR", 2) == 0) {
			op = TOKEN_OR, oplen = 2;
		}
		if (op && oplen >= ff) {
			ctx->cursor = s + oplen;
			return op;
		}
		if (c == '\'' || c == '"') {
			qlen = quoted_len(s);
		}
		if (qlen && qlen >= ff) {
			ctx->cursor = s + qlen;
			lval->len = qlen;
			lval->str = strndup(s + 1, qlen - 2);
			return TOKEN_QUOTED_STRING;
		}
		ctx->cursor = s + ff;
		lval->len = ff;
		lval->str = strndup(s, ff);
		return TOKEN_FF_STRING;
	}
}

/*
 * Recursive-descent parser with one token of look-ahead.
 */

typedef struct {
	query_t *	q;
	int		tok;
	char *		str;	// owned string of a pending value token
	unsigned	depth;
} pstate_t;

static void
p_advance(pstate_t *p)
{
	p->tok = lex(p->q);
	p->str = NULL;
	if (p->tok == TOKEN_FF_STRING || p->tok == TOKEN_QUOTED_STRING) {
		p->str = p->q->lval.str;
	}
}

static expr_t *
p_fail(pstate_t *p, expr_t *e1, expr_t *e2)
{
	if (!p->q->error) {
		query_set_error(p->q);
	}
	if (e1) expr_destroy(e1);
	if (e2) expr_destroy(e2);
	return NULL;
}

static expr_t *p_or(pstate_t *);

static expr_t *
p_primary(pstate_t *p)
{
	expr_t *e;

	if (p->tok == TOKEN_FF_STRING || p->tok == TOKEN_QUOTED_STRING) {
		e = expr_create_token(p->str);
		p_advance(p);
		return e;
	}
	if (p->tok == TOKEN_BR_OPEN) {
		p_advance(p);
		if ((e = p_or(p)) == NULL) {
			return NULL;
		}
		if (p->tok != TOKEN_BR_CLOSE) {
			return p_fail(p, e, NULL);
		}
		p_advance(p);
		return e;
	}
	return p_fail(p, NULL, NULL);
}

static expr_t *
p_and(pstate_t *p)
{
	expr_t *l, *r;

	if ((l = p_primary(p)) == NULL) {
		return NULL;
	}
	while (p->tok == TOKEN_AND) {
		expr_type_t type = EXPR_OP_AND;

		p_advance(p);
		if (p->tok == TOKEN_NOT) {
			type = EXPR_OP_NOT;
			p_advance(p);
		}
		if ((r = p_primary(p)) == NULL) {
			expr_destroy(l);
			return NULL;
		}
		l = expr_create_operator(type, l, r);
	}
	return l;
}

static expr_t *
p_or(pstate_t *p)
{
	expr_t *l, *r;

	if ((l = p_and(p)) == NULL) {
		return NULL;
	}
	while (p->tok == TOKEN_OR) {
		p_advance(p);
		if ((r = p_and(p)) == NULL) {
			expr_destroy(l);
			return NULL;
		}
		l = expr_create_operator(EXPR_OP_OR, l, r);
	}
	return l;
}

int
query_parse(query_t *q, const char *query)
{
	pstate_t p = { .q = q };
	expr_t *root, *r;

	lex_init(&q->lexer, query);
	p_advance(&p);

	if ((root = p_or(&p)) == NULL) {
		goto fail;
	}
	/* expr_list ::= expr_list expr -- implicit OR, top level only. */
	while (p.tok == TOKEN_FF_STRING || p.tok == TOKEN_QUOTED_STRING ||
	    p.tok == TOKEN_BR_OPEN) {
		if ((r = p_or(&p)) == NULL) {
			expr_destroy(root);
			goto fail;
		}
		root = expr_create_operator(EXPR_OP_OR, root, r);
	}
	if (p.tok != 0) {
		p_fail(&p, root, NULL);
		goto fail;
	}
	q->root = root;
	return 0;
fail:
	/* A value token that was lexed but never consumed. */
	free(p.str);
	q->root = NULL;
	return 0;
}
