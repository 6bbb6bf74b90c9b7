/*
 * Parameters: a JSON object behind an opaque handle.
 * Mirrors ref src/core/params.c (same keys, same getter semantics: a getter
 * fails unless the value has exactly the asked-for JSON type).
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nxs_impl.h"

nxs_params_t *
nxs_params_wrap(jval_t *root)
{
	nxs_params_t *p;

	if (!root)
		return NULL;
	if ((p = calloc(1, sizeof(*p))) == NULL) {
		json_free(root);
		return NULL;
	}
	p->root = root;
	return p;
}

NXS_API nxs_params_t *
nxs_params_create(void)
{
	return nxs_params_wrap(json_new(J_OBJ));
}

NXS_API void
nxs_params_release(nxs_params_t *p)
{
	if (p) {
		json_free(p->root);
		free(p);
	}
}

static int
params_add(nxs_params_t *p, const char *key, jval_t *val)
{
	if (!val)
		return -1;
	if (json_obj_add(p->root, key, val) == -1) {
		json_free(val);
		return -1;
	}
	return 0;
}

NXS_API int
nxs_params_set_strlist(nxs_params_t *p, const char *key, const char **vals,
    size_t count)
{
	jval_t *arr = json_new(J_ARR);

	for (size_t i = 0; arr && i < count; i++) {
		jval_t *s = json_new_str(vals[i]);

		if (!s || json_arr_append(arr, s) == -1) {
			json_free(s);
			json_free(arr);
			return -1;
		}
	}
	return params_add(p, key, arr);
}

NXS_API int
nxs_params_set_str(nxs_params_t *p, const char *key, const char *val)
{
	return params_add(p, key, json_new_str(val));
}

NXS_API int
nxs_params_set_uint(nxs_params_t *p, const char *key, uint64_t val)
{
	return params_add(p, key, json_new_uint(val));
}

NXS_API int
nxs_params_set_bool(nxs_params_t *p, const char *key, bool val)
{
	return params_add(p, key, json_new_bool(val));
}

/* The caller free(3)s the returned array (not the strings); ref params.c:108. */
const char **
nxs_params_get_strlist(nxs_params_t *p, const char *key, size_t *count)
{
	const jval_t *arr = json_obj_get(p->root, key);
	const char **out;
	size_t n = 0;

	if (!arr || arr->type != J_ARR || arr->c.n == 0)
		return NULL;
	if ((out = calloc(arr->c.n, sizeof(char *))) == NULL)
		return NULL;
	for (size_t i = 0; i < arr->c.n; i++) {
		if (arr->c.items[i]->type == J_STR)
			out[n++] = arr->c.items[i]->s;
	}
	*count = n;
	return out;
}

const char *
nxs_params_get_str(nxs_params_t *p, const char *key)
{
	const jval_t *v = json_obj_get(p->root, key);
	return v && v->type == J_STR ? v->s : NULL;
}

int
nxs_params_get_uint(nxs_params_t *p, const char *key, uint64_t *val)
{
	const jval_t *v = json_obj_get(p->root, key);

	if (!v || v->type != J_UINT)
		return -1;
	*val = v->u;
	return 0;
}

int
nxs_params_get_bool(nxs_params_t *p, const char *key, bool *val)
{
	const jval_t *v = json_obj_get(p->root, key);

	if (!v || v->type != J_BOOL)
		return -1;
	*val = v->b;
	return 0;
}

NXS_API char *
nxs_params_tojson(const nxs_params_t *p, size_t *len)
{
	return json_write(p->root, true, len);
}

NXS_API nxs_params_t *
nxs_params_fromjson(nxs_t *nxs, const char *json, size_t len)
{
	char err[128];
	jval_t *root = json_parse(json, len, err, sizeof(err));

	if (!root || root->type != J_OBJ) {
		if (nxs)
			nxs_set_error(nxs, NXS_ERR_SYSTEM, "params parsing failed: %s",
			    root ? "not a JSON object" : err);
		json_free(root);
		return NULL;
	}
	return nxs_params_wrap(root);
}

int
nxs_params_serialize(nxs_t *nxs, const nxs_params_t *p, const char *path)
{
	size_t len;
	char *s = nxs_params_tojson(p, &len);
	FILE *fp;

	if (!s || (fp = fopen(path, "w")) == NULL) {
		nxs_set_syserror(nxs, NXS_ERR_SYSTEM, "params serialize failed");
		free(s);
		return -1;
	}
	if (fwrite(s, 1, len, fp) != len || fclose(fp) != 0) {
		nxs_set_syserror(nxs, NXS_ERR_SYSTEM, "params serialize failed");
		free(s);
		return -1;
	}
	free(s);
	return 0;
}

nxs_params_t *
nxs_params_unserialize(nxs_t *nxs, const char *path)
{
	FILE *fp = fopen(path, "r");
	nxs_params_t *p = NULL;
	char *buf = NULL;
	long len;

	if (!fp) {
		nxs_set_syserror(nxs, NXS_ERR_SYSTEM, "params parsing failed: %s", path);
		return NULL;
	}
	if (fseek(fp, 0, SEEK_END) == 0 && (len = ftell(fp)) >= 0 &&
	    fseek(fp, 0, SEEK_SET) == 0 && (buf = malloc(len + 1)) != NULL &&
	    fread(buf, 1, len, fp) == (size_t)len) {
		p = nxs_params_fromjson(nxs, buf, len);
	} else {
		nxs_set_syserror(nxs, NXS_ERR_SYSTEM, "params parsing failed: %s", path);
	}
	free(buf);
	fclose(fp);
	return p;
}
