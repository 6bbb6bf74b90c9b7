/*
 * Index: the two on-disk files, their in-memory view, and the GPU image.
 *
 * On-disk format and multi-process protocol are the reference's
 * (src/index/storage.h, terms.c, dtmap.c, idxmap.c): append-only files,
 * header data_len published with a release store after the payload, writers
 * under flock(LOCK_EX).  What differs is the in-memory side: instead of one
 * roaring64 bitmap per term (filled posting by posting) the host keeps only
 * flat per-document bookkeeping, and the reverse index is built on the GPU.
 */
#ifndef NXSB_INDEX_H
#define NXSB_INDEX_H

#include "nxs_impl.h"
#include "tokenizer.h"

#define IDX_SIZE_STEP		(32UL * 1024)	/* ref index.h:24 */
#define TERMS_HDR_LEN		16
#define DTMAP_HDR_LEN		32

typedef struct {
	int		fd;
	uint8_t *	base;
	size_t		mapped_len;
} idxfile_t;

typedef struct {
	uint32_t *	parent;		/* [n] term index, UINT32_MAX = root */
	uint8_t *	edge;		/* [n] label of the edge to the parent */
	uint32_t *	rank;		/* [n] BFS rank */
	uint32_t	n;		/* terms inserted so far */
	uint32_t	cap;
} bkmirror_t;

struct nxs_index {
	nxs_t *			nxs;
	nxs_index_t *		next;
	char *			name;
	nxs_params_t *		params;
	filter_pipeline_t *	fp;
	int			algo;		/* NXSB_ALGO_* or -1 */

	/* Terms (nxsterms). */
	idxfile_t		tfile;
	size_t			terms_consumed;
	uint32_t		n_terms;	/* == last term id */
	uint32_t		max_term_len;	/* longest vocabulary term, bytes */
	uint32_t		terms_cap;
	strmap_t *		term_map;	/* value -> term id */
	char *			term_blob;
	size_t			blob_len, blob_cap;
	uint32_t *		term_off;	/* [n_terms + 1] into term_blob */
	uint32_t *		term_total_off;	/* [n_terms] file offset of total */

	/* Documents (nxsdtmap). */
	idxfile_t		dfile;
	size_t			dt_consumed;
	u64map_t *		doc_map;	/* doc id -> slot */
	bool			doc_map_ready;	/* false: a bulk open left it for the first writer */
	uint64_t *		doc_ids;	/* per slot, file order */
	uint32_t *		doc_len;
	uint32_t *		doc_n;
	uint64_t *		doc_blk;	/* file offset of the block */
	uint8_t *		doc_dead;
	uint32_t		n_slots, slots_cap, n_live;
	uint32_t		slots_want;	/* bulk sync: capacity to grow to at once */
	/*
	 * Live documents per term -- what the cardinality of a term's roaring
	 * bitmap is to the reference (ranking.c:78,150) -- kept current as
	 * blocks and deletion markers are consumed.
	 */
	uint32_t *		df;
	uint32_t		df_cap;

	/*
	 * GPU side.  The image is a base segment (0) plus delta segments
	 * (1..n_segs) holding what was appended since; documents removed
	 * after their segment was built are listed per segment and dropped
	 * when the per-segment results are merged (index.c "GPU image").
	 */
	nxsb_engine_t *		engine;
	bool			image_dirty;	/* nothing usable on the GPU: build it all */
	bool			stats_dirty;	/* df[] / header counters moved */
	bool			vocab_dirty;
	bool			totals_dirty;	/* term totals moved since the vocabulary upload */
	uint8_t *		doc_seg;	/* per slot; valid below built_slots */
	uint32_t		built_slots;	/* slots the image knows of */
	uint32_t		n_pending;	/* live slots at or above built_slots */
	uint32_t		n_segs;
	uint32_t		seg_live[NXSB_MAX_SEGMENTS + 1];
	uint64_t *		seg_dead[NXSB_MAX_SEGMENTS + 1];
	uint32_t		seg_ndead[NXSB_MAX_SEGMENTS + 1];
	uint32_t		seg_dead_cap[NXSB_MAX_SEGMENTS + 1];
	bool			seg_dead_dirty[NXSB_MAX_SEGMENTS + 1];
	/* How the image got to its current state (nxsb_index_image_stats). */
	uint64_t		n_full_builds, n_delta_builds, n_consolidations;
	bkmirror_t		bk;
};

int		idx_terms_open(nxs_index_t *, const char *path);
int		idx_terms_sync(nxs_index_t *);
int		idx_terms_add(nxs_index_t *, tokenset_t *);
void		idx_terms_close(nxs_index_t *);
uint64_t	idx_term_total(const nxs_index_t *, uint32_t term_id);

int		idx_dtmap_open(nxs_index_t *, const char *path);
int		idx_dtmap_sync(nxs_index_t *, bool partial);
int		idx_dtmap_add(nxs_index_t *, nxs_doc_id_t, tokenset_t *);
int		idx_dtmap_remove(nxs_index_t *, nxs_doc_id_t);
void		idx_dtmap_close(nxs_index_t *);
/* The id -> slot map is only needed to add / remove: built on first use. */
int		idx_docmap_ensure(nxs_index_t *);
uint64_t	idx_get_token_count(const nxs_index_t *);
uint32_t	idx_get_doc_count(const nxs_index_t *);

/* Exact term lookup: id or 0 (ref idxterm.c:192-196). */
uint32_t	idx_term_lookup(const nxs_index_t *, const char *, size_t);

/* (Re)build the HBM image if the files moved on; creates the engine. */
int		idx_gpu_prepare(nxs_index_t *, bool need_vocab);

/* BK-tree mirror (bkmirror.c). */
int		bk_levdist(const char *a, size_t n, const char *b, size_t m);
int		bkmirror_update(bkmirror_t *, const char *blob,
		    const uint32_t *term_off, uint32_t n_terms);
void		bkmirror_free(bkmirror_t *);

#endif
