/*
 * Query lexer and parser; see query.h.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include "query.h"
#include "nxsb200_gpu.h"

#define PARSE_DEPTH_MAX		2000	/* parenthesis recursion guard */

/*
 * Lexer.  One candidate set per position, the longest wins and the earlier
 * rule wins a tie -- which, for this token set, reduces to: an operator
 * keyword or a quoted string is recognised only if it is at least as long as
 * the free-form run starting at the same place.
 */

static inline bool
is_space(unsigned char c)
{
	return c == ' ' || (c >= '\t' && c <= '\r');
}

static inline bool
is_freeform(unsigned char c)
{
	return c && !is_space(c) && c != '(' && c != ')';
}

void
qlex_init(qlexer_t *lx, const char *s)
{
	memset(lx, 0, sizeof(*lx));
	lx->cursor = lx->cur_line = s;
	lx->line = 1;
}

static size_t
quoted_span(const char *s)
{
	const char quote = *s;

	for (size_t i = 1; s[i]; i++) {
		if (s[i] == '\\') {
			if (!s[++i])
				break;
		} else if (s[i] == quote) {
			return i + 1;
		}
	}
	return 0;	/* unterminated */
}

qtok_t
qlex_next(qlexer_t *lx)
{
	for (;;) {
		const char *s = lx->cursor;
		const unsigned char c = *s;
		size_t run = 0, kw = 0, qs;
		qtok_t kwtok = QTOK_EOF;

		lx->token = s;
		if (!c)
			return QTOK_EOF;
		if (is_space(c)) {
			while (is_space(s[run]))
				run++;
			if (c == '\n' && run == 1) {
				lx->cur_line = s;
				lx->line++;
			}
			lx->cursor += run;
			continue;
		}
		if (c == '(' || c == ')') {
			lx->cursor++;
			return c == '(' ? QTOK_BR_OPEN : QTOK_BR_CLOSE;
		}
		while (is_freeform(s[run]))
			run++;

		switch (c) {
		case '&':
			kw = 1, kwtok = QTOK_AND;
			break;
		case '|':
			kw = 1, kwtok = QTOK_OR;
			break;
		case 'a': case 'A':
			if (strncasecmp(s, "and", 3) == 0)
				kw = 3, kwtok = QTOK_AND;
			break;
		case 'n': case 'N':
			if (strncasecmp(s, "not", 3) == 0)
				kw = 3, kwtok = QTOK_NOT;
			break;
		case 'o': case 'O':
			if (strncasecmp(s, "or", 2) == 0)
				kw = 2, kwtok = QTOK_OR;
			break;
		}
		if (kw && kw >= run) {
			lx->cursor += kw;
			return kwtok;
		}
		if ((c == '\'' || c == '"') && (qs = quoted_span(s)) >= run && qs) {
			lx->cursor += qs;
			lx->len = qs;
			lx->str = strndup(s + 1, qs - 2);	/* escapes kept */
			return QTOK_QUOTED_STRING;
		}
		lx->cursor += run;
		lx->len = run;
		lx->str = strndup(s, run);
		return QTOK_FF_STRING;
	}
}

/*
 * Parser: precedence climbing over a flat node array.
 */

typedef struct {
	qtree_t *	tree;
	qlexer_t	lx;
	qtok_t		tok;
	unsigned	nesting;
} qparser_t;

static void
parser_advance(qparser_t *ps)
{
	ps->tok = qlex_next(&ps->lx);
}

static int32_t
syntax_error(qparser_t *ps)
{
	qtree_t *t = ps->tree;

	if (!t->error) {
		const unsigned col = (unsigned)(ps->lx.token - ps->lx.cur_line);

		t->error = true;
		if (asprintf(&t->errmsg, "syntax error near %u:%u: \"%.50s ...\"",
		    ps->lx.line, col, ps->lx.token) == -1)
			t->errmsg = NULL;
	}
	return -1;
}

static int32_t
new_node(qtree_t *t, qnode_type_t type, int32_t l, int32_t r, char *value)
{
	if (t->n_nodes == t->cap) {
		const int32_t ncap = t->cap * 2;
		qnode_t *nn = t->nodes == t->inl ? malloc(sizeof(qnode_t) * ncap)
		    : realloc(t->nodes, sizeof(qnode_t) * ncap);

		if (!nn) {
			free(value);
			t->error = true;
			return -1;
		}
		if (t->nodes == t->inl)
			memcpy(nn, t->inl, sizeof(qnode_t) * t->n_nodes);
		t->nodes = nn;
		t->cap = ncap;
	}
	t->nodes[t->n_nodes] = (qnode_t){
		.type = type, .left = l, .right = r, .value = value, .token = -1
	};
	return t->n_nodes++;
}

static int32_t parse_or(qparser_t *);

static int32_t
parse_primary(qparser_t *ps)
{
	int32_t e;

	if (ps->tok == QTOK_FF_STRING || ps->tok == QTOK_QUOTED_STRING) {
		char *v = ps->lx.str;

		ps->lx.str = NULL;
		e = new_node(ps->tree, QN_VALUE, -1, -1, v);
		parser_advance(ps);
		return e;
	}
	if (ps->tok == QTOK_BR_OPEN) {
		if (++ps->nesting > PARSE_DEPTH_MAX)
			return syntax_error(ps);
		parser_advance(ps);
		if ((e = parse_or(ps)) < 0)
			return -1;
		if (ps->tok != QTOK_BR_CLOSE)
			return syntax_error(ps);
		ps->nesting--;
		parser_advance(ps);
		return e;
	}
	return syntax_error(ps);
}

static int32_t
parse_and(qparser_t *ps)
{
	int32_t l = parse_primary(ps), r;

	while (l >= 0 && ps->tok == QTOK_AND) {
		qnode_type_t type = QN_AND;

		parser_advance(ps);
		if (ps->tok == QTOK_NOT) {
			type = QN_NOT;
			parser_advance(ps);
		}
		if ((r = parse_primary(ps)) < 0)
			return -1;
		l = new_node(ps->tree, type, l, r, NULL);
	}
	return l;
}

static int32_t
parse_or(qparser_t *ps)
{
	int32_t l = parse_and(ps), r;

	while (l >= 0 && ps->tok == QTOK_OR) {
		parser_advance(ps);
		if ((r = parse_and(ps)) < 0)
			return -1;
		l = new_node(ps->tree, QN_OR, l, r, NULL);
	}
	return l;
}

static unsigned
tree_depth(const qtree_t *t)
{
	/* Children always precede their parent in the node array. */
	unsigned small[QTREE_INLINE_NODES] = { 0 };
	unsigned *d = t->n_nodes <= QTREE_INLINE_NODES ? small
	    : calloc(t->n_nodes, sizeof(unsigned));
	unsigned max = 0;

	if (!d)
		return 0;
	for (int32_t i = t->n_nodes - 1; i >= 0; i--) {
		const qnode_t *n = &t->nodes[i];

		if (d[i] > max)
			max = d[i];
		if (n->type != QN_VALUE) {
			d[n->left] = d[n->right] = d[i] + 1;
		}
	}
	if (d != small)
		free(d);
	return max;
}

int
qtree_parse(qtree_t *t, const char *query)
{
	qparser_t ps = { .tree = t };
	int32_t root, r;

	t->nodes = t->inl;
	t->n_nodes = 0;
	t->cap = QTREE_INLINE_NODES;
	t->depth = 0;
	t->error = false;
	t->errmsg = NULL;
	t->root = -1;
	qlex_init(&ps.lx, query);
	parser_advance(&ps);

	root = parse_or(&ps);
	while (root >= 0 && (ps.tok == QTOK_FF_STRING ||
	    ps.tok == QTOK_QUOTED_STRING || ps.tok == QTOK_BR_OPEN)) {
		if ((r = parse_or(&ps)) < 0) {
			root = -1;
			break;
		}
		root = new_node(t, QN_OR, root, r, NULL);
	}
	if (root >= 0 && ps.tok != QTOK_EOF)
		root = syntax_error(&ps);
	free(ps.lx.str);
	if (root < 0 || t->error) {
		t->error = true;
		t->root = -1;
		return 0;
	}
	t->root = root;
	/* Only the nodes reachable from the root count; all are, by construction. */
	t->depth = tree_depth(t);
	return 0;
}

void
qtree_free(qtree_t *t)
{
	for (int32_t i = 0; i < t->n_nodes; i++)
		free(t->nodes[i].value);
	if (t->nodes != t->inl)
		free(t->nodes);
	free(t->errmsg);
	t->nodes = NULL;
	t->n_nodes = t->cap = 0;
	t->errmsg = NULL;
	t->root = -1;
}

static char *
dump_node(const qtree_t *t, int32_t i)
{
	static const char *names[] = {
		[QN_AND] = "AND", [QN_OR] = "OR", [QN_NOT] = "NOT"
	};
	const qnode_t *n = &t->nodes[i];
	char *out = NULL, *l, *r;

	if (n->type == QN_VALUE) {
		if (asprintf(&out, "`%s`", n->value) == -1)
			return NULL;
		return out;
	}
	l = dump_node(t, n->left);
	r = dump_node(t, n->right);
	if (!l || !r || asprintf(&out, "(%s %s %s)", names[n->type], l, r) == -1)
		out = NULL;
	free(l);
	free(r);
	return out;
}

char *
qtree_dump(const qtree_t *t)
{
	return t->root >= 0 ? dump_node(t, t->root) : NULL;
}

/* Post-order emission of the boolean program (depth is bounded by then). */
void
qtree_emit_program(const qtree_t *t, int32_t node, int32_t *prog, uint32_t *n)
{
	const qnode_t *nd = &t->nodes[node];

	if (nd->type == QN_VALUE) {
		/*
		 * A leaf a filter discarded is the empty set (ref search.c:
		 * 133-141).  A leaf without a term keeps its slot with term id
		 * 0, which the engine treats as an empty list -- the reference
		 * reads freed memory there; the empty set is the
		 * defined behaviour here.
		 */
		prog[(*n)++] = nd->token >= 0 ? nd->token : NXSB_OP_EMPTY;
		return;
	}
	qtree_emit_program(t, nd->left, prog, n);
	qtree_emit_program(t, nd->right, prog, n);
	prog[(*n)++] = nd->type == QN_AND ? NXSB_OP_AND :
	    nd->type == QN_OR ? NXSB_OP_OR : NXSB_OP_ANDNOT;
}
