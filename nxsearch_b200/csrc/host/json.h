/*
 * Minimal JSON document model for parameters and responses.
 *
 * The reference keeps nxs_params_t and the response body as yyjson mutable
 * documents (ref src/core/params.c:19-22, src/core/results.c:37-42).  Only
 * what that usage needs is implemented: objects (insertion-ordered, lookup
 * returns the FIRST match and adding never replaces -- the behaviour of
 * yyjson_mut_obj_add / yyjson_mut_obj_get the reference relies on), arrays,
 * strings, unsigned/signed integers, reals, booleans, null; a strict parser;
 * a writer that prints like yyjson (compact, or 4-space pretty; reals in the
 * shortest round-trip ECMAScript form with a kept ".0").
 */
#ifndef NXSB_JSON_H
#define NXSB_JSON_H

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

typedef enum {
	J_NULL, J_BOOL, J_UINT, J_SINT, J_REAL, J_STR, J_ARR, J_OBJ
} jtype_t;

typedef struct jval {
	jtype_t		type;
	union {
		bool		b;
		uint64_t	u;
		int64_t		i;
		double		d;
		char *		s;
		struct {
			struct jval **	items;
			char **		keys;	// J_OBJ only
			size_t		n, cap;
		} c;
	};
} jval_t;

jval_t *	json_new(jtype_t);
jval_t *	json_new_str(const char *);
jval_t *	json_new_uint(uint64_t);
jval_t *	json_new_real(double);
jval_t *	json_new_bool(bool);
void		json_free(jval_t *);

int		json_arr_append(jval_t *arr, jval_t *val);
int		json_obj_add(jval_t *obj, const char *key, jval_t *val);
jval_t *	json_obj_get(const jval_t *obj, const char *key);

/* Parse exactly one JSON value (surrounding whitespace allowed). */
jval_t *	json_parse(const char *s, size_t len, char *err, size_t errlen);
/* Serialise; the result is malloc'ed and NUL-terminated. */
char *		json_write(const jval_t *, bool pretty, size_t *len);

/* Shortest round-trip text of a double, yyjson style; buf >= 40 bytes. */
size_t		json_format_real(double, char *buf);

#endif
