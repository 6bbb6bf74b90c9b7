/*
 * Instance, error slot and index lifecycle: the public entry points of
 * ref src/core/nxs.c with the same directory layout
 * ($basedir/data/<index>/{params.db,nxsterms,nxsdtmap}), defaults and error
 * codes.  The search entry points live in search.c.
 */
#define _GNU_SOURCE
#include <sys/stat.h>

#include <ctype.h>
#include <errno.h>
#include <inttypes.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <unistd.h>

#include "index.h"

/*
 * ref nxs.c:249-268 adds "stemmer" as well; this build has no Snowball and
 * refuses the filter rather than pass terms through silently (tokenizer.c).
 */
static const char *default_filters[] = { "normalizer", "stopwords" };

int
str_isalnumdu(const char *s)
{
	for (; *s; s++) {
		if (!isalnum((unsigned char)*s) && *s != '-' && *s != '_')
			return -1;
	}
	return 0;
}

void
nxs_clear_error(nxs_t *nxs)
{
	free(nxs->errmsg);
	nxs->errmsg = NULL;
	nxs->errcode = NXS_ERR_SUCCESS;
}

static void
set_error_v(nxs_t *nxs, nxs_err_t code, int sys_errno, const char *fmt, va_list ap)
{
	char *s = NULL, *msg = NULL;

	if (vasprintf(&s, fmt, ap) == -1)
		s = NULL;
	if (sys_errno >= 0 && s) {
		if (asprintf(&msg, "%s: %s", s, strerror(sys_errno)) == -1)
			msg = NULL;
		free(s);
	} else {
		msg = s;
	}
	if (!nxs) {
		free(msg);
		return;
	}
	free(nxs->errmsg);
	nxs->errmsg = msg;
	nxs->errcode = code;
}

void
nxs_set_error(nxs_t *nxs, nxs_err_t code, const char *fmt, ...)
{
	va_list ap;

	va_start(ap, fmt);
	set_error_v(nxs, code, -1, fmt, ap);
	va_end(ap);
}

void
nxs_set_syserror(nxs_t *nxs, nxs_err_t code, const char *fmt, ...)
{
	const int e = errno;
	va_list ap;

	va_start(ap, fmt);
	set_error_v(nxs, code, e, fmt, ap);
	va_end(ap);
}

void
nxs_error_checkpoint(nxs_t *nxs)
{
	/* Error paths that declared nothing get a generic fatal (nxs.c:162). */
	if (!nxs->errcode)
		nxs_set_syserror(nxs, NXS_ERR_FATAL, "internal error; last system errno");
}

NXS_API nxs_err_t
nxs_get_error(const nxs_t *nxs, const char **msg)
{
	if (msg)
		*msg = nxs->errmsg;
	return nxs->errcode;
}

/* "0-3", "0,2,5", "1": anything else leaves the single device of NXS_GPU_DEVICE. */
static void
parse_device_list(nxs_t *nxs, const char *s)
{
	int n = 0;

	while (*s && n < NXS_MAX_GPU_DEVICES) {
		char *end;
		long a = strtol(s, &end, 10), b;

		if (end == s || a < 0)
			return;
		b = a;
		if (*end == '-') {
			s = end + 1;
			b = strtol(s, &end, 10);
			if (end == s || b < a)
				return;
		}
		for (long d = a; d <= b && n < NXS_MAX_GPU_DEVICES; d++)
			nxs->devices[n++] = (int)d;
		if (*end != ',' && *end != '\0')
			return;
		s = *end ? end + 1 : end;
	}
	if (n >= 1) {
		nxs->n_devices = n;
		nxs->device = nxs->devices[0];
	}
}

NXS_API nxs_t *
nxs_open(const char *basedir)
{
	nxs_t *nxs = calloc(1, sizeof(*nxs));
	const char *s;
	char *path = NULL;

	if (!nxs)
		return NULL;
	s = basedir ? basedir : getenv("NXS_BASEDIR");
	if (!s || (nxs->basedir = realpath(s, NULL)) == NULL)
		goto err;
	if (asprintf(&path, "%s/data", nxs->basedir) == -1)
		goto err;
	if (mkdir(path, 0755) == -1 && errno != EEXIST)
		goto err;
	free(path);
	path = NULL;
	if ((s = getenv("NXS_GPU_DEVICE")) != NULL)
		nxs->device = atoi(s);
	if ((s = getenv("NXS_GPU_DEVICES")) != NULL)
		parse_device_list(nxs, s);
	if ((s = getenv("NXS_GPU_LAYOUT")) != NULL) {
		if (strcmp(s, "shards") == 0) {
			nxs->shards = true;
		} else if (strcmp(s, "replicas") != 0) {
			errno = EINVAL;	/* a typo must not fall back to the other layout */
			goto err;
		}
	}
	return nxs;
err:
	free(path);
	free(nxs->basedir);
	free(nxs);
	return NULL;
}

NXS_API void
nxs_close(nxs_t *nxs)
{
	while (nxs->indexes)
		nxs_index_close(nxs->indexes);
	free(nxs->basedir);
	free(nxs->errmsg);
	free(nxs);
}

NXS_API int
nxs_luafilter_load(nxs_t *nxs, const char *name, const char *code)
{
	(void)name; (void)code;
	nxs_clear_error(nxs);
	nxs_set_error(nxs, NXS_ERR_INVALID,
	    "Lua filters are not supported by the nxsearch-b200 build");
	return -1;
}

static nxs_index_t *
find_open_index(nxs_t *nxs, const char *name)
{
	for (nxs_index_t *i = nxs->indexes; i; i = i->next) {
		if (strcmp(i->name, name) == 0)
			return i;
	}
	return NULL;
}

/*
 * $basedir/data/<index>[/<file>]: every path of an index comes from here.
 * NULL (with the error slot set) when memory runs out.
 */
static char *
index_path(nxs_t *nxs, const char *index, const char *file)
{
	char *path = NULL;

	if (asprintf(&path, file ? "%s/data/%s/%s" : "%s/data/%s", nxs->basedir, index, file) == -1) {
		nxs_set_error(nxs, NXS_ERR_SYSTEM, "out of memory");
		return NULL;
	}
	return path;
}

/* Index names are path components: [A-Za-z0-9_-]+ only (ref nxs.c:227). */
static bool
index_name_ok(nxs_t *nxs, const char *name)
{
	if (str_isalnumdu(name) == 0)
		return true;
	nxs_set_error(nxs, NXS_ERR_INVALID, "invalid characters in index name");
	return false;
}

/* What an index created without the key gets (ref nxs.c:249-268). */
static int
apply_index_defaults(nxs_params_t *params)
{
	static const struct { const char *key, *val; } strs[] = {
		{ "algo", NXS_DEFAULT_RANKING_ALGO },
		{ "lang", NXS_DEFAULT_LANGUAGE },
	};
	size_t n = 0;
	const char **list = nxs_params_get_strlist(params, "filters", &n);
	const bool has_filters = list != NULL;

	free(list);
	if (!has_filters && nxs_params_set_strlist(params, "filters", default_filters,
	    sizeof(default_filters) / sizeof(default_filters[0])) == -1)
		return -1;
	for (size_t i = 0; i < sizeof(strs) / sizeof(strs[0]); i++) {
		if (!nxs_params_get_str(params, strs[i].key) &&
		    nxs_params_set_str(params, strs[i].key, strs[i].val) == -1)
			return -1;
	}
	return 0;
}

NXS_API nxs_index_t *
nxs_index_create(nxs_t *nxs, const char *name, nxs_params_t *params)
{
	nxs_params_t *own = NULL;
	nxs_index_t *idx = NULL;
	char *dir, *db = NULL;

	nxs_clear_error(nxs);
	if (!index_name_ok(nxs, name) || (dir = index_path(nxs, name, NULL)) == NULL)
		return NULL;

	/* The directory is the claim on the name: whoever makes it owns the index. */
	if (mkdir(dir, 0755) == -1) {
		const bool taken = errno == EEXIST;

		nxs_set_syserror(nxs, taken ? NXS_ERR_EXISTS : NXS_ERR_SYSTEM,
		    taken ? "index `%s' already exists" : "could not create directory at %s",
		    taken ? name : dir);
		free(dir);
		nxs_error_checkpoint(nxs);
		return NULL;
	}
	free(dir);

	if (!params)
		params = own = nxs_params_create();
	if (params && apply_index_defaults(params) == 0 &&
	    (db = index_path(nxs, name, "params.db")) != NULL &&
	    nxs_params_serialize(nxs, params, db) == 0)
		idx = nxs_index_open(nxs, name);
	if (!idx)
		nxs_error_checkpoint(nxs);
	if (own)
		nxs_params_release(own);
	free(db);
	return idx;
}

NXS_API int
nxs_index_destroy(nxs_t *nxs, const char *name)
{
	/* The three files, then the directory that held them (file == NULL). */
	static const char *const parts[] = { "params.db", "nxsterms", "nxsdtmap", NULL };

	nxs_clear_error(nxs);
	if (!index_name_ok(nxs, name))
		return -1;
	for (size_t i = 0; i < sizeof(parts) / sizeof(parts[0]); i++) {
		char *path = index_path(nxs, name, parts[i]);
		const int rc = !path ? -1 : parts[i] ? unlink(path) : rmdir(path);

		if (path && rc == -1)
			nxs_set_syserror(nxs, NXS_ERR_SYSTEM, "could not remove `%s'", path);
		free(path);
		if (rc == -1) {
			nxs_error_checkpoint(nxs);
			return -1;
		}
	}
	return 0;
}

static int
algo_id(const char *name)
{
	if (strcasecmp(name, "TF-IDF") == 0)
		return NXSB_ALGO_TFIDF;
	if (strcasecmp(name, "BM25") == 0)
		return NXSB_ALGO_BM25;
	return -1;
}

/* params.db -> ranking algorithm and filter pipeline of the index. */
static int
open_params(nxs_t *nxs, nxs_index_t *idx, const char *name)
{
	char *db = index_path(nxs, name, "params.db");
	const char *algo;
	struct stat sb;

	if (!db)
		return -1;
	if (stat(db, &sb) == -1 && errno == ENOENT) {
		nxs_set_error(nxs, NXS_ERR_MISSING, "index `%s' does not exist", name);
		free(db);
		return -1;
	}
	idx->params = nxs_params_unserialize(nxs, db);
	free(db);
	if (!idx->params)
		return -1;
	if ((algo = nxs_params_get_str(idx->params, "algo")) == NULL) {
		nxs_set_error(nxs, NXS_ERR_FATAL, "corrupted index params");
		return -1;
	}
	idx->algo = algo_id(algo);
	idx->fp = filter_pipeline_create(nxs, idx->params);
	return idx->fp ? 0 : -1;
}

/* One of the two append-only files of the index through its opener. */
static int
open_file(nxs_t *nxs, nxs_index_t *idx, const char *name, const char *file,
    int (*opener)(nxs_index_t *, const char *))
{
	char *path = index_path(nxs, name, file);
	const int rc = path ? opener(idx, path) : -1;

	free(path);
	return rc;
}

NXS_API nxs_index_t *
nxs_index_open(nxs_t *nxs, const char *name)
{
	nxs_index_t *idx;

	nxs_clear_error(nxs);
	if (!index_name_ok(nxs, name))
		return NULL;
	if (find_open_index(nxs, name)) {
		nxs_set_error(nxs, NXS_ERR_EXISTS, "index `%s' is already open", name);
		return NULL;
	}
	if ((idx = calloc(1, sizeof(*idx))) == NULL) {
		nxs_error_checkpoint(nxs);
		return NULL;
	}
	idx->nxs = nxs;
	idx->algo = -1;
	if (open_params(nxs, idx, name) == -1 ||
	    open_file(nxs, idx, name, "nxsterms", idx_terms_open) == -1 ||
	    open_file(nxs, idx, name, "nxsdtmap", idx_dtmap_open) == -1 ||
	    (idx->name = strdup(name)) == NULL) {
		nxs_error_checkpoint(nxs);
		nxs_index_close(idx);	/* not on the list yet: name is unset or just set */
		return NULL;
	}
	/* Nothing is in HBM until the first search asks for it. */
	idx->image_dirty = idx->vocab_dirty = true;
	idx->next = nxs->indexes;
	nxs->indexes = idx;
	return idx;
}

NXS_API nxs_params_t *
nxs_index_get_params(nxs_index_t *idx)
{
	return idx->params;
}

NXS_API void
nxs_index_close(nxs_index_t *idx)
{
	nxs_t *nxs = idx->nxs;

	if (idx->name) {
		for (nxs_index_t **pp = &nxs->indexes; *pp; pp = &(*pp)->next) {
			if (*pp == idx) {
				*pp = idx->next;
				break;
			}
		}
		free(idx->name);
	}
	if (idx->engine)
		nxsb_engine_destroy(idx->engine);
	bkmirror_free(&idx->bk);
	if (idx->fp)
		filter_pipeline_destroy(idx->fp);
	if (idx->params)
		nxs_params_release(idx->params);
	if (idx->doc_map || idx->dfile.base)
		idx_dtmap_close(idx);
	if (idx->term_map || idx->tfile.base)
		idx_terms_close(idx);
	free(idx);
}

NXS_API int
nxs_index_add(nxs_index_t *idx, nxs_params_t *params, nxs_doc_id_t doc_id,
    const char *text, size_t len)
{
	tokenset_t *ts;
	int ret = -1;

	(void)params;
	nxs_clear_error(idx->nxs);
	if (doc_id == 0) {
		nxs_set_error(idx->nxs, NXS_ERR_INVALID, "document ID must be non-zero");
		return -1;
	}
	if (idx_docmap_ensure(idx) == -1) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "document map build failed");
		return -1;
	}
	if (u64map_get(idx->doc_map, doc_id, NULL)) {
		nxs_set_error(idx->nxs, NXS_ERR_EXISTS,
		    "document %" PRIu64 " is already indexed", doc_id);
		return -1;
	}
	if ((ts = tokenize(idx->fp, text, len)) == NULL) {
		nxs_set_error(idx->nxs, NXS_ERR_FATAL, "tokenizer failed");
		return -1;
	}
	if (ts->count == 0) {
		nxs_set_error(idx->nxs, NXS_ERR_MISSING,
		    "the text is empty or no meaningful tokens found");
		goto out;
	}
	for (uint32_t i = 0; i < ts->count; i++)
		ts->list[i].term_id = idx_term_lookup(idx, ts->list[i].str,
		    ts->list[i].len);
	if (idx_terms_add(idx, ts) == -1 || idx_dtmap_add(idx, doc_id, ts) == -1)
		goto out;
	ret = 0;
out:
	tokenset_destroy(ts);
	if (ret != 0)
		nxs_error_checkpoint(idx->nxs);
	return ret;
}

NXS_API int
nxs_index_remove(nxs_index_t *idx, nxs_doc_id_t doc_id)
{
	nxs_clear_error(idx->nxs);
	if (idx_dtmap_remove(idx, doc_id) == -1) {
		nxs_error_checkpoint(idx->nxs);
		return -1;
	}
	return 0;
}
