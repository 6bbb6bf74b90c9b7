/*
 * nxsearch-b200: public C API.
 *
 * This is the drop-in boundary.  Every declaration below has the name,
 * argument meaning, ownership rule and error behaviour of the reference's
 * public header (reference: src/core/nxs.h:26-101, docs/c-api.md), so a
 * caller of the reference library -- its Lua binding (src/core/lua.c:341),
 * its CLI (src/utils/benchmark.c:204) or its tests (src/tests/helpers.c:262)
 * -- links against libnxsearch.so from this tree unchanged.  The single hot
 * path behind nxs_index_search() runs on a B200 (sm_100a); see DESIGN.md.
 *
 * Conventions (reference: docs/c-api.md):
 *  - pointer-returning calls yield NULL on error, int-returning calls -1;
 *    nxs_get_error() then gives the code and a message owned by the library;
 *  - the caller releases nxs_resp_t / its own nxs_params_t, and free(3)s the
 *    strings returned by the *_tojson() calls;
 *  - one nxs_t per thread; concurrency is multi-process over the index files.
 *
 * nxs_index_search_batch() at the bottom is ADDITIVE (not in the reference).
 */
#ifndef NXSB200_NXS_H
#define NXSB200_NXS_H

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Everything declared here is exported; the rest of the library is hidden. */
#pragma GCC visibility push(default)

typedef uint64_t nxs_doc_id_t;			/* ref nxs.h:21; 0 is invalid */

typedef struct nxs nxs_t;
typedef struct nxs_params nxs_params_t;
typedef struct nxs_index nxs_index_t;
typedef struct nxs_resp nxs_resp_t;

/* Error codes: values are ABI-frozen (ref nxs.h:35-46). */
typedef enum {
	NXS_ERR_SUCCESS	= 0,
	NXS_ERR_FATAL	= 1,	/* unspecified fatal error */
	NXS_ERR_SYSTEM	= 2,	/* operating system (or CUDA runtime) error */
	NXS_ERR_INVALID	= 3,	/* invalid parameter or value */
	NXS_ERR_EXISTS	= 4,	/* resource already exists */
	NXS_ERR_MISSING	= 5,	/* resource is missing */
	NXS_ERR_LIMIT	= 6,	/* resource limit reached */
} nxs_err_t;

/* Instance (ref nxs.c:88-147).  basedir NULL => $NXS_BASEDIR. */
nxs_t *		nxs_open(const char *basedir);
void		nxs_close(nxs_t *);
nxs_err_t	nxs_get_error(const nxs_t *, const char **msg);

/* Lua filters are out of scope here: always fails with NXS_ERR_INVALID. */
int		nxs_luafilter_load(nxs_t *, const char *name, const char *code);

/* Parameters: a JSON object (ref params.c). */
nxs_params_t *	nxs_params_create(void);
nxs_params_t *	nxs_params_fromjson(nxs_t *, const char *json, size_t len);
int		nxs_params_set_strlist(nxs_params_t *, const char *key,
		    const char **vals, size_t count);
int		nxs_params_set_str(nxs_params_t *, const char *key,
		    const char *val);
int		nxs_params_set_uint(nxs_params_t *, const char *key,
		    uint64_t val);
int		nxs_params_set_bool(nxs_params_t *, const char *key, bool val);
char *		nxs_params_tojson(const nxs_params_t *, size_t *len);
void		nxs_params_release(nxs_params_t *);

/* Index lifecycle (ref nxs.c:219-560). */
nxs_index_t *	nxs_index_create(nxs_t *, const char *name, nxs_params_t *);
int		nxs_index_destroy(nxs_t *, const char *name);
nxs_params_t *	nxs_index_get_params(nxs_index_t *);
nxs_index_t *	nxs_index_open(nxs_t *, const char *name);
void		nxs_index_close(nxs_index_t *);
int		nxs_index_add(nxs_index_t *, nxs_params_t *, nxs_doc_id_t,
		    const char *text, size_t len);
int		nxs_index_remove(nxs_index_t *, nxs_doc_id_t);

/*
 * Search (ref search.c:285-342).  `len` is ignored, as in the reference: the
 * query must be NUL-terminated.  Params (all optional): "limit" (uint,
 * default 1000), "algo" ("BM25" | "TF-IDF"), "fuzzymatch" (bool, default on).
 */
nxs_resp_t *	nxs_index_search(nxs_index_t *, nxs_params_t *,
		    const char *query, size_t len);

/* Response (ref results.c:88-247). */
void		nxs_resp_iter_reset(nxs_resp_t *);
bool		nxs_resp_iter_result(nxs_resp_t *, nxs_doc_id_t *, float *);
unsigned	nxs_resp_resultcount(const nxs_resp_t *);
char *		nxs_resp_tojson(nxs_resp_t *, size_t *len);
void		nxs_resp_release(nxs_resp_t *);

/*
 * ADDITIVE: run n queries as one GPU batch with shared params.  Returns 0
 * and fills resps[0..n) (each to be released by the caller; an entry is NULL
 * when that query alone failed, e.g. a syntax error -- nxs_get_error() then
 * describes the LAST such failure), or -1 if the batch as a whole failed.
 */
int		nxs_index_search_batch(nxs_index_t *, nxs_params_t *,
		    const char *const *queries, size_t n, nxs_resp_t **resps);

/*
 * ADDITIVE: the same batch in two halves, so that the host work of the next
 * batch (parsing, term lookup) overlaps the GPU work of this one.  _begin
 * parses, resolves and submits the batch and returns at once; _end waits for
 * it, fills resps[0..n) as nxs_index_search_batch() does and frees the batch
 * handle (resps = NULL abandons the results).  Up to 4 batches may be in
 * flight per index; they complete in submission order.  Every batch must be
 * ended before nxs_index_close().  nxs_index_search_batch() is _begin
 * followed by _end.
 */
typedef struct nxs_batch nxs_batch_t;
nxs_batch_t *	nxs_index_search_batch_begin(nxs_index_t *, nxs_params_t *,
		    const char *const *queries, size_t n);
int		nxs_index_search_batch_end(nxs_batch_t *, nxs_resp_t **resps);

#pragma GCC visibility pop

#ifdef __cplusplus
}
#endif

#endif
