/*
 * Minimal JSON document model; see json.h.
 */
#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "json.h"

jval_t *
json_new(jtype_t type)
{
	jval_t *v = calloc(1, sizeof(*v));

	if (v)
		v->type = type;
	return v;
}

jval_t *
json_new_str(const char *s)
{
	jval_t *v = json_new(J_STR);

	if (v && (v->s = strdup(s)) == NULL) {
		free(v);
		v = NULL;
	}
	return v;
}

jval_t *
json_new_uint(uint64_t u)
{
	jval_t *v = json_new(J_UINT);

	if (v)
		v->u = u;
	return v;
}

jval_t *
json_new_real(double d)
{
	jval_t *v = json_new(J_REAL);

	if (v)
		v->d = d;
	return v;
}

jval_t *
json_new_bool(bool b)
{
	jval_t *v = json_new(J_BOOL);

	if (v)
		v->b = b;
	return v;
}

void
json_free(jval_t *v)
{
	if (!v)
		return;
	if (v->type == J_STR) {
		free(v->s);
	} else if (v->type == J_ARR || v->type == J_OBJ) {
		for (size_t i = 0; i < v->c.n; i++) {
			json_free(v->c.items[i]);
			if (v->type == J_OBJ)
				free(v->c.keys[i]);
		}
		free(v->c.items);
		free(v->c.keys);
	}
	free(v);
}

static int
container_grow(jval_t *v)
{
	if (v->c.n < v->c.cap)
		return 0;
	const size_t ncap = v->c.cap ? v->c.cap * 2 : 8;
	jval_t **ni = realloc(v->c.items, ncap * sizeof(jval_t *));

	if (!ni)
		return -1;
	v->c.items = ni;
	if (v->type == J_OBJ) {
		char **nk = realloc(v->c.keys, ncap * sizeof(char *));

		if (!nk)
			return -1;
		v->c.keys = nk;
	}
	v->c.cap = ncap;
	return 0;
}

int
json_arr_append(jval_t *arr, jval_t *val)
{
	if (!arr || arr->type != J_ARR || !val || container_grow(arr) == -1)
		return -1;
	arr->c.items[arr->c.n++] = val;
	return 0;
}

int
json_obj_add(jval_t *obj, const char *key, jval_t *val)
{
	char *k;

	if (!obj || obj->type != J_OBJ || !val || container_grow(obj) == -1)
		return -1;
	if ((k = strdup(key)) == NULL)
		return -1;
	obj->c.keys[obj->c.n] = k;
	obj->c.items[obj->c.n++] = val;
	return 0;
}

jval_t *
json_obj_get(const jval_t *obj, const char *key)
{
	if (!obj || obj->type != J_OBJ)
		return NULL;
	for (size_t i = 0; i < obj->c.n; i++) {
		if (strcmp(obj->c.keys[i], key) == 0)
			return obj->c.items[i];
	}
	return NULL;
}

/*
 * Parser.
 */

typedef struct {
	const char *	p;
	const char *	end;
	const char *	start;
	const char *	errmsg;
	unsigned	depth;
} jparse_t;

static void
skip_ws(jparse_t *jp)
{
	while (jp->p < jp->end && (*jp->p == ' ' || *jp->p == '\t' ||
	    *jp->p == '\n' || *jp->p == '\r'))
		jp->p++;
}

static int
put_utf8(char *out, uint32_t cp)
{
	if (cp < 0x80) {
		out[0] = cp;
		return 1;
	}
	if (cp < 0x800) {
		out[0] = 0xc0 | (cp >> 6);
		out[1] = 0x80 | (cp & 0x3f);
		return 2;
	}
	if (cp < 0x10000) {
		out[0] = 0xe0 | (cp >> 12);
		out[1] = 0x80 | ((cp >> 6) & 0x3f);
		out[2] = 0x80 | (cp & 0x3f);
		return 3;
	}
	out[0] = 0xf0 | (cp >> 18);
	out[1] = 0x80 | ((cp >> 12) & 0x3f);
	out[2] = 0x80 | ((cp >> 6) & 0x3f);
	out[3] = 0x80 | (cp & 0x3f);
	return 4;
}

static int
hex4(const char *p, uint32_t *out)
{
	uint32_t v = 0;

	for (int i = 0; i < 4; i++) {
		const char c = p[i];

		v <<= 4;
		if (c >= '0' && c <= '9') v |= c - '0';
		else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
		else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
		else return -1;
	}
	*out = v;
	return 0;
}

static char *
parse_string(jparse_t *jp)
{
	char *out = malloc((size_t)(jp->end - jp->p) + 1);
	size_t n = 0;

	if (!out)
		return NULL;
	jp->p++;	// opening quote
	while (jp->p < jp->end) {
		const unsigned char c = *jp->p++;

		if (c == '"') {
			out[n] = '\0';
			return out;
		}
		if (c < 0x20)
			break;
		if (c != '\\') {
			out[n++] = c;
			continue;
		}
		if (jp->p >= jp->end)
			break;
		switch (*jp->p++) {
		case '"':  out[n++] = '"'; break;
		case '\\': out[n++] = '\\'; break;
		case '/':  out[n++] = '/'; break;
		case 'b':  out[n++] = '\b'; break;
		case 'f':  out[n++] = '\f'; break;
		case 'n':  out[n++] = '\n'; break;
		case 'r':  out[n++] = '\r'; break;
		case 't':  out[n++] = '\t'; break;
		case 'u': {
			uint32_t cp, lo;

			if (jp->end - jp->p < 4 || hex4(jp->p, &cp) == -1)
				goto bad;
			jp->p += 4;
			if (cp >= 0xd800 && cp < 0xdc00) {
				if (jp->end - jp->p < 6 || jp->p[0] != '\\' ||
				    jp->p[1] != 'u' || hex4(jp->p + 2, &lo) == -1 ||
				    lo < 0xdc00 || lo > 0xdfff)
					goto bad;
				jp->p += 6;
				cp = 0x10000 + ((cp - 0xd800) << 10) + (lo - 0xdc00);
			}
			n += put_utf8(out + n, cp);
			break;
		}
		default:
			goto bad;
		}
	}
bad:
	free(out);
	jp->errmsg = "invalid string";
	return NULL;
}

static jval_t *parse_value(jparse_t *);

static jval_t *
parse_number(jparse_t *jp)
{
	const char *s = jp->p;
	bool neg = false, real = false;
	char buf[64];
	size_t n;

	if (jp->p < jp->end && *jp->p == '-') {
		neg = true;
		jp->p++;
	}
	if (jp->p >= jp->end || !isdigit((unsigned char)*jp->p))
		goto bad;
	if (*jp->p == '0' && jp->p + 1 < jp->end && isdigit((unsigned char)jp->p[1]))
		goto bad;
	while (jp->p < jp->end && isdigit((unsigned char)*jp->p))
		jp->p++;
	if (jp->p < jp->end && *jp->p == '.') {
		real = true;
		jp->p++;
		if (jp->p >= jp->end || !isdigit((unsigned char)*jp->p))
			goto bad;
		while (jp->p < jp->end && isdigit((unsigned char)*jp->p))
			jp->p++;
	}
	if (jp->p < jp->end && (*jp->p == 'e' || *jp->p == 'E')) {
		real = true;
		jp->p++;
		if (jp->p < jp->end && (*jp->p == '+' || *jp->p == '-'))
			jp->p++;
		if (jp->p >= jp->end || !isdigit((unsigned char)*jp->p))
			goto bad;
		while (jp->p < jp->end && isdigit((unsigned char)*jp->p))
			jp->p++;
	}
	n = jp->p - s;
	if (n >= sizeof(buf))
		goto bad;
	memcpy(buf, s, n);
	buf[n] = '\0';

	if (!real) {
		errno = 0;
		if (neg) {
			const long long v = strtoll(buf, NULL, 10);
			if (errno == 0) {
				jval_t *j = json_new(J_SINT);
				if (j)
					j->i = v;
				return j;
			}
		} else {
			const unsigned long long v = strtoull(buf, NULL, 10);
			if (errno == 0)
				return json_new_uint(v);
		}
		/* Out of 64-bit range: falls back to a real, like yyjson. */
	}
	return json_new_real(strtod(buf, NULL));
bad:
	jp->errmsg = "invalid number";
	return NULL;
}

static jval_t *
parse_container(jparse_t *jp, bool is_obj)
{
	jval_t *c = json_new(is_obj ? J_OBJ : J_ARR);
	const char close = is_obj ? '}' : ']';

	if (!c)
		return NULL;
	jp->p++;
	skip_ws(jp);
	if (jp->p < jp->end && *jp->p == close) {
		jp->p++;
		return c;
	}
	for (;;) {
		char *key = NULL;
		jval_t *v;

		skip_ws(jp);
		if (is_obj) {
			if (jp->p >= jp->end || *jp->p != '"') {
				jp->errmsg = "object key expected";
				goto bad;
			}
			if ((key = parse_string(jp)) == NULL)
				goto bad;
			skip_ws(jp);
			if (jp->p >= jp->end || *jp->p != ':') {
				free(key);
				jp->errmsg = "':' expected";
				goto bad;
			}
			jp->p++;
		}
		if ((v = parse_value(jp)) == NULL) {
			free(key);
			goto bad;
		}
		if ((is_obj ? json_obj_add(c, key, v) : json_arr_append(c, v)) == -1) {
			free(key);
			json_free(v);
			goto bad;
		}
		free(key);
		skip_ws(jp);
		if (jp->p < jp->end && *jp->p == ',') {
			jp->p++;
			continue;
		}
		if (jp->p < jp->end && *jp->p == close) {
			jp->p++;
			return c;
		}
		jp->errmsg = "',' or closing bracket expected";
		goto bad;
	}
bad:
	json_free(c);
	return NULL;
}

static jval_t *
parse_value(jparse_t *jp)
{
	jval_t *v = NULL;

	skip_ws(jp);
	if (jp->p >= jp->end) {
		jp->errmsg = "unexpected end of input";
		return NULL;
	}
	if (++jp->depth > 256) {
		jp->errmsg = "nesting too deep";
		return NULL;
	}
	switch (*jp->p) {
	case '{':
		v = parse_container(jp, true);
		break;
	case '[':
		v = parse_container(jp, false);
		break;
	case '"': {
		char *s = parse_string(jp);

		if (s && (v = json_new(J_STR)) != NULL)
			v->s = s;
		else
			free(s);
		break;
	}
	case 't':
		if (jp->end - jp->p >= 4 && memcmp(jp->p, "true", 4) == 0) {
			jp->p += 4;
			v = json_new_bool(true);
		}
		break;
	case 'f':
		if (jp->end - jp->p >= 5 && memcmp(jp->p, "false", 5) == 0) {
			jp->p += 5;
			v = json_new_bool(false);
		}
		break;
	case 'n':
		if (jp->end - jp->p >= 4 && memcmp(jp->p, "null", 4) == 0) {
			jp->p += 4;
			v = json_new(J_NULL);
		}
		break;
	default:
		v = parse_number(jp);
		break;
	}
	if (!v && !jp->errmsg)
		jp->errmsg = "unexpected character";
	jp->depth--;
	return v;
}

jval_t *
json_parse(const char *s, size_t len, char *err, size_t errlen)
{
	jparse_t jp = { .p = s, .end = s + len, .start = s };
	jval_t *v = parse_value(&jp);

	if (v) {
		skip_ws(&jp);
		if (jp.p != jp.end) {
			json_free(v);
			v = NULL;
			jp.errmsg = "unexpected content after document";
		}
	}
	if (!v && err && errlen) {
		snprintf(err, errlen, "%s at %zu",
		    jp.errmsg ? jp.errmsg : "parse error", (size_t)(jp.p - jp.start));
	}
	return v;
}

/*
 * Writer.
 */

typedef struct {
	char *	buf;
	size_t	len, cap;
	bool	oom;
} jout_t;

static void
out_bytes(jout_t *o, const char *s, size_t n)
{
	if (o->oom)
		return;
	if (o->len + n + 1 > o->cap) {
		size_t ncap = o->cap ? o->cap * 2 : 256;
		char *nb;

		while (ncap < o->len + n + 1)
			ncap *= 2;
		if ((nb = realloc(o->buf, ncap)) == NULL) {
			o->oom = true;
			return;
		}
		o->buf = nb;
		o->cap = ncap;
	}
	memcpy(o->buf + o->len, s, n);
	o->len += n;
}

static void
out_str(jout_t *o, const char *s)
{
	out_bytes(o, s, strlen(s));
}

static void
out_quoted(jout_t *o, const char *s)
{
	out_bytes(o, "\"", 1);
	for (; *s; s++) {
		const unsigned char c = *s;
		char esc[8];

		switch (c) {
		case '"':  out_bytes(o, "\\\"", 2); break;
		case '\\': out_bytes(o, "\\\\", 2); break;
		case '\b': out_bytes(o, "\\b", 2); break;
		case '\f': out_bytes(o, "\\f", 2); break;
		case '\n': out_bytes(o, "\\n", 2); break;
		case '\r': out_bytes(o, "\\r", 2); break;
		case '\t': out_bytes(o, "\\t", 2); break;
		default:
			if (c < 0x20) {
				snprintf(esc, sizeof(esc), "\\u%04X", c);
				out_bytes(o, esc, 6);
			} else {
				out_bytes(o, (const char *)&c, 1);
			}
		}
	}
	out_bytes(o, "\"", 1);
}

size_t
json_format_real(double d, char *buf)
{
	char digits[32], tmp[48];
	int ndig = 0, e10, dot_pos;
	size_t n = 0;

	if (isnan(d) || isinf(d)) {
		memcpy(buf, "null", 5);
		return 4;
	}
	if (signbit(d)) {
		buf[n++] = '-';
		d = -d;
	}
	if (d == 0) {
		memcpy(buf + n, "0.0", 4);
		return n + 3;
	}
	/* Shortest digit string that reads back to the same double. */
	for (int prec = 0; prec <= 16; prec++) {
		snprintf(tmp, sizeof(tmp), "%.*e", prec, d);
		if (strtod(tmp, NULL) == d || prec == 16)
			break;
	}
	{
		const char *p = tmp;
		char *e = strchr(tmp, 'e');

		for (; p < e; p++) {
			if (*p != '.')
				digits[ndig++] = *p;
		}
		while (ndig > 1 && digits[ndig - 1] == '0')
			ndig--;
		e10 = atoi(e + 1);
	}
	dot_pos = e10 + 1;	// digits before the decimal point

	if (-6 < dot_pos && dot_pos <= 21) {
		if (dot_pos <= 0) {
			buf[n++] = '0';
			buf[n++] = '.';
			for (int i = 0; i < -dot_pos; i++)
				buf[n++] = '0';
			memcpy(buf + n, digits, ndig);
			n += ndig;
		} else if (ndig <= dot_pos) {
			memcpy(buf + n, digits, ndig);
			n += ndig;
			for (int i = ndig; i < dot_pos; i++)
				buf[n++] = '0';
			buf[n++] = '.';
			buf[n++] = '0';
		} else {
			memcpy(buf + n, digits, dot_pos);
			n += dot_pos;
			buf[n++] = '.';
			memcpy(buf + n, digits + dot_pos, ndig - dot_pos);
			n += ndig - dot_pos;
		}
	} else {
		buf[n++] = digits[0];
		if (ndig > 1) {
			buf[n++] = '.';
			memcpy(buf + n, digits + 1, ndig - 1);
			n += ndig - 1;
		}
		n += sprintf(buf + n, "e%d", e10);
	}
	buf[n] = '\0';
	return n;
}

static void
out_indent(jout_t *o, unsigned depth)
{
	out_bytes(o, "\n", 1);
	for (unsigned i = 0; i < depth; i++)
		out_bytes(o, "    ", 4);
}

static void
write_value(jout_t *o, const jval_t *v, bool pretty, unsigned depth)
{
	char num[48];

	switch (v->type) {
	case J_NULL:
		out_str(o, "null");
		break;
	case J_BOOL:
		out_str(o, v->b ? "true" : "false");
		break;
	case J_UINT:
		snprintf(num, sizeof(num), "%llu", (unsigned long long)v->u);
		out_str(o, num);
		break;
	case J_SINT:
		snprintf(num, sizeof(num), "%lld", (long long)v->i);
		out_str(o, num);
		break;
	case J_REAL:
		out_bytes(o, num, json_format_real(v->d, num));
		break;
	case J_STR:
		out_quoted(o, v->s);
		break;
	case J_ARR:
	case J_OBJ: {
		const bool obj = v->type == J_OBJ;

		out_bytes(o, obj ? "{" : "[", 1);
		for (size_t i = 0; i < v->c.n; i++) {
			if (i)
				out_bytes(o, ",", 1);
			if (pretty)
				out_indent(o, depth + 1);
			if (obj) {
				out_quoted(o, v->c.keys[i]);
				out_bytes(o, pretty ? ": " : ":", pretty ? 2 : 1);
			}
			write_value(o, v->c.items[i], pretty, depth + 1);
		}
		if (pretty && v->c.n)
			out_indent(o, depth);
		out_bytes(o, obj ? "}" : "]", 1);
		break;
	}
	}
}

char *
json_write(const jval_t *v, bool pretty, size_t *len)
{
	jout_t o = { 0 };

	write_value(&o, v, pretty, 0);
	if (o.oom || !o.buf) {
		free(o.buf);
		return NULL;
	}
	o.buf[o.len] = '\0';
	if (len)
		*len = o.len;
	return o.buf;
}
