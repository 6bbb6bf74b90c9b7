/*
 * nxsearch-b200 host library: internal declarations.
 */
#ifndef NXSB_NXS_IMPL_H
#define NXSB_NXS_IMPL_H

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

#include "nxs.h"
#include "nxsb200_gpu.h"
#include "json.h"
#include "hashmap.h"

#define NXS_API		__attribute__((visibility("default")))

#define NXS_DEFAULT_RESULTS_LIMIT	1000		/* ref nxs_impl.h:39 */
#define NXS_DEFAULT_RANKING_ALGO	"BM25"		/* ref nxs_impl.h:40 */
#define NXS_DEFAULT_LANGUAGE		"en"		/* ref nxs_impl.h:41 */
#define NXS_MAX_GPU_DEVICES	16
#define NXS_QUERY_RLIMIT		100		/* ref search.c:70 */

struct nxs {
	char *		basedir;
	char *		errmsg;
	nxs_err_t	errcode;
	nxs_index_t *	indexes;	/* singly linked list of open indexes */
	int		device;		/* CUDA device ordinal ($NXS_GPU_DEVICE) */
	/*
	 * $NXS_GPU_DEVICES ("0-7", "0,2,3"): more than one device => every
	 * index of this instance keeps a replica of its image on each and a
	 * batch's queries are split between them (nxsb_engine_create_replicated).
	 */
	int		devices[NXS_MAX_GPU_DEVICES];
	int		n_devices;
	/*
	 * $NXS_GPU_LAYOUT=shards: the devices hold one range of the documents
	 * each instead (nxsb_engine_create_sharded) -- an index larger than one
	 * GPU; every device scores every query and the top-N lists are merged.
	 */
	bool		shards;
};

struct nxs_params {
	jval_t *	root;		/* always a J_OBJ */
};

/* Error slot (ref nxs.c:154-217). */
void		nxs_clear_error(nxs_t *);
void		nxs_set_error(nxs_t *, nxs_err_t, const char *fmt, ...)
		    __attribute__((format(printf, 3, 4)));
/* As nxs_set_error, appending ": strerror(errno)". */
void		nxs_set_syserror(nxs_t *, nxs_err_t, const char *fmt, ...)
		    __attribute__((format(printf, 3, 4)));
void		nxs_error_checkpoint(nxs_t *);

/* Params internals (ref nxs_impl.h:65-71). */
nxs_params_t *	nxs_params_wrap(jval_t *root);
int		nxs_params_serialize(nxs_t *, const nxs_params_t *, const char *path);
nxs_params_t *	nxs_params_unserialize(nxs_t *, const char *path);
const char **	nxs_params_get_strlist(nxs_params_t *, const char *, size_t *);
const char *	nxs_params_get_str(nxs_params_t *, const char *);
int		nxs_params_get_uint(nxs_params_t *, const char *, uint64_t *);
int		nxs_params_get_bool(nxs_params_t *, const char *, bool *);

/* Response construction. */
nxs_resp_t *	nxs_resp_from_arrays(const uint64_t *ids, const float *scores,
		    uint32_t n);

int		str_isalnumdu(const char *);

#endif
