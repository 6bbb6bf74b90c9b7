/*
 * Query language: lexer, parser and the intermediate representation.
 *
 * Same language as the reference (lexer: ref src/query/scan.re:43-119,
 * grammar: ref src/query/grammar.y:62-110):
 *
 *   query     := or_expr { or_expr }            juxtaposition = implicit OR,
 *                                               TOP LEVEL ONLY
 *   or_expr   := and_expr { OR and_expr }       OR  is '|' or 'or'
 *   and_expr  := primary { AND [NOT] primary }  AND is '&' or 'and';
 *                                               "AND NOT" is one operator
 *   primary   := value | '(' or_expr ')'
 *   value     := free-form string | 'quoted' | "quoted"
 *
 * all operators left-associative, keywords case-insensitive; a keyword is
 * only a keyword when followed by a separator (re2c longest match).
 */
#ifndef NXSB_QUERY_H
#define NXSB_QUERY_H

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

typedef enum {
	QTOK_EOF = 0,
	QTOK_OR, QTOK_AND, QTOK_NOT, QTOK_BR_OPEN, QTOK_BR_CLOSE,
	QTOK_FF_STRING, QTOK_QUOTED_STRING,
} qtok_t;

typedef enum {				/* ref expr.h:14-21 */
	QN_VALUE, QN_AND, QN_OR, QN_NOT,
} qnode_type_t;

typedef struct {
	qnode_type_t	type;
	int32_t		left, right;	/* node indexes (operators) */
	char *		value;		/* QN_VALUE: the leaf string (owned) */
	int32_t		token;		/* QN_VALUE: token slot, -1 = none */
} qnode_t;

typedef struct {
	const char *	cursor;
	const char *	token;		/* start of the last lexed token */
	const char *	cur_line;
	unsigned	line;
	char *		str;		/* value of a string token (owned) */
	size_t		len;		/* its source length incl. quotes */
} qlexer_t;

#define QTREE_INLINE_NODES	12	/* a typical query never allocates its node array */

typedef struct {
	qnode_t *	nodes;
	qnode_t		inl[QTREE_INLINE_NODES];
	int32_t		n_nodes, cap;
	int32_t		root;		/* -1 when empty / failed */
	unsigned	depth;		/* deepest node, root = 0 */
	bool		error;
	char *		errmsg;		/* "syntax error near L:C: ..." */
} qtree_t;

void		qlex_init(qlexer_t *, const char *);
qtok_t		qlex_next(qlexer_t *);

/* Returns 0; syntax errors are reported through tree->error / errmsg. */
int		qtree_parse(qtree_t *, const char *query);
void		qtree_free(qtree_t *);
/* "(AND (OR `A` `B`) `C`)"-style dump (ref tests/t_queryparser.c:139-162). */
char *		qtree_dump(const qtree_t *);

/* Post-order emission of the boolean program: token slots and NXSB_OP_* codes. */
void		qtree_emit_program(const qtree_t *, int32_t node, int32_t *prog,
		    uint32_t *n);

#endif
