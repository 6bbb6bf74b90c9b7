/*
 * nxsearch-b200: the thin C ABI between the C11 host library and the
 * sm_100a CUDA engine.  Plain pointers and sizes only; no CUDA or torch
 * types.  This is the seam the reference's internals would bind if the GPU
 * engine were dropped into the reference tree:
 *
 *   nxsb_engine_load_shard   <- the in-memory reverse index built by
 *                               idx_terms_sync / idx_dtmap_sync
 *                               (ref src/index/terms.c:320-414,
 *                                src/index/dtmap.c:386-544)
 *   nxsb_engine_search       <- run_query_logic + nxs_resp_build
 *                               (ref src/query/search.c:118-278,
 *                                src/core/results.c:128-220) with the
 *                               ranking functions of src/algo/ranking.c
 *   nxsb_engine_fuzzy        <- idxterm_fuzzysearch
 *                               (ref src/index/idxterm.c:210-249 over
 *                                src/algo/bktree.c, src/algo/levdist.c)
 *
 * All functions return 0 on success and -1 on failure unless stated;
 * nxsb_engine_errmsg() / nxsb_last_error() describe the failure.  There is
 * NO CPU fallback: without a usable CUDA device nxsb_engine_create() fails.
 */
#ifndef NXSB200_GPU_H
#define NXSB200_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Everything declared here is exported; the rest of the library is hidden. */
#pragma GCC visibility push(default)

typedef struct nxsb_engine nxsb_engine_t;

enum { NXSB_ALGO_TFIDF = 0, NXSB_ALGO_BM25 = 1 };	/* ref index.h:30-34 */

/* Postfix boolean program over a query's token slots. */
enum {
	NXSB_OP_EMPTY	= -1,	/* push the empty set (unresolved leaf) */
	NXSB_OP_AND	= -2,
	NXSB_OP_OR	= -3,
	NXSB_OP_ANDNOT	= -4,
	/* >= 0: push the document set of token slot i */
};

#define NXSB_MAX_QUERY_TOKENS	32	/* resolved tokens per query */
#define NXSB_MAX_QUERY_PROG	128	/* postfix program length */
#ifndef NXSB_TILE_DOCS
#define NXSB_TILE_DOCS		16384	/* documents per scoring tile */
#endif

/* Number of CUDA devices visible; 0 when there is no driver / device. */
int		nxsb_gpu_device_count(void);
const char *	nxsb_last_error(void);

nxsb_engine_t *	nxsb_engine_create(int device);
/*
 * One engine object over n devices, a complete replica of the image on each:
 * load_shard / segment_add / set_dead / set_global_stats go to every replica
 * (side by side on host threads), a batch's queries are split in n contiguous
 * shares that run concurrently, and the results come back in query order --
 * the host library's NXS_GPU_DEVICES.  An index of 10M documents is ~6 GB per
 * device; when the index outgrows one GPU, shard instead (nxsearch_b200/dist.py).
 * Resident batches, external streams, _begin_dev and merge_topk are refused.
 */
nxsb_engine_t *	nxsb_engine_create_replicated(const int *devices, int n);
int		nxsb_engine_replica_count(const nxsb_engine_t *);
/*
 * One engine object over n devices, a contiguous range of the
 * documents on each -- for an index that outgrows one GPU, behind the same
 * calls (the host library's NXS_GPU_DEVICES with NXS_GPU_LAYOUT=shards):
 *   load_shard    cuts the documents (ascending ids) into n ranges of about
 *                 equal posting counts; every device builds its range with the
 *                 whole-index df / N / token count, as ranking needs
 *                 (ref ranking.c:77-78,149-150,163);
 *   segment_add   a delta segment goes whole to one device, round robin;
 *   set_dead      ids of the base image go to the device whose range holds
 *                 them, ids of a delta segment to the device that has it;
 *   search[_begin/_end]  every device scores the WHOLE batch on its documents
 *                 into device memory, the per-device top-k lists travel to
 *                 the first device (cudaMemcpyPeerAsync: NVLink where peers
 *                 can map each other) and one kernel merges them in
 *                 (score desc, id desc) order ahead of a single D2H copy.
 * Results are bit-identical to one engine over the whole index.  The same
 * calls are refused as on a replicated engine.
 */
nxsb_engine_t *	nxsb_engine_create_sharded(const int *devices, int n);
int		nxsb_engine_is_sharded(const nxsb_engine_t *);
void		nxsb_engine_destroy(nxsb_engine_t *);
const char *	nxsb_engine_errmsg(const nxsb_engine_t *);

/*
 * Run all engine work on an externally owned CUDA stream (a cudaStream_t
 * passed as void *; NULL restores the engine's own stream).
 */
int		nxsb_engine_set_stream(nxsb_engine_t *, void *cuda_stream);
/*
 * Without an external stream the engine runs searches on two streams of its
 * own ("lanes"): nxsb_engine_search_begin slots and nxsb_engine_batch_run
 * handles alternate between them, so that two batches in flight overlap on
 * the GPU (the image is shared, arenas exist per lane; NXSB_LANES=1 turns it
 * off).  lane_stream returns a lane's cudaStream_t (as void *) for callers
 * that time with their own events; lanes_join makes lane 0 wait for what the
 * other lanes have been given so far, so that an event recorded on lane 0
 * afterwards closes a region that spans both.
 */
void *		nxsb_engine_lane_stream(nxsb_engine_t *, int lane);
int		nxsb_engine_lanes_join(nxsb_engine_t *);

/*
 * One document shard in document-major form (host memory), documents in
 * ASCENDING external-id order.  pairs[2j], pairs[2j+1] = (term id 1-based,
 * count) for j in [doc_off[i], doc_off[i+1]).  doc_count / token_count / df
 * are WHOLE-INDEX statistics: BM25 and TF-IDF use the global N, df and
 * average length (ref ranking.c:77-78,149-150,163), so a shard must not
 * substitute its own.  df may be NULL when the shard is the whole index.
 *
 * Instead of pairs/doc_off the shard may be given as the bytes of the
 * `nxsdtmap` file itself: raw points at host memory holding
 * the file (or any part of it), and document i's raw_n[i] big-endian
 * (term id u32, count u32) pairs start at raw + raw_off[i] (ref
 * src/index/storage.h:80-84; raw_off[i] is a multiple of 8).  The bytes are
 * copied to the device as they are and decoded there.
 */
typedef struct nxsb_shard_desc {
	uint32_t		n_docs;
	uint32_t		n_terms;
	const uint64_t *	doc_ids;
	const uint32_t *	doc_len;
	const uint64_t *	doc_off;
	const uint32_t *	pairs;
	uint64_t		token_count;
	uint32_t		doc_count;
	const uint32_t *	df;
	const void *		raw;		/* NULL: use pairs / doc_off */
	const uint64_t *	raw_off;
	const uint32_t *	raw_n;
} nxsb_shard_desc_t;

/* Build (or rebuild) the HBM-resident CSR image of the shard. */
int		nxsb_engine_load_shard(nxsb_engine_t *, const nxsb_shard_desc_t *);

/*
 * Incremental refresh of the image (the reference applies every
 * appended dtmap block to its in-memory index on the next search,
 * ref src/index/dtmap.c:357-441, src/query/search.c:309-310).  The image is
 * the base shard plus up to NXSB_MAX_SEGMENTS delta segments on the same GPU:
 *
 * segment_add   builds the image of the documents appended since the last
 *               build (same descriptor as load_shard; df[] -- whole-index
 *               counts -- is required).  Returns the segment number >= 1.
 * set_dead      ids (strictly ascending) removed from segment `segment`
 *               (0 = the base shard) after it was built.  Searches ask every
 *               segment for limit + (longest dead list) results and drop the
 *               dead ones while merging, so results equal a full rebuild.
 * segments_drop forgets all delta segments and all dead lists;
 *               load_shard does the same.
 * set_global_stats (below) reaches every segment; n_terms may exceed the
 * vocabulary a segment was built with (term ids only grow).
 *
 * Searches on a segmented image go through nxsb_engine_search and
 * nxsb_engine_search_begin/_begin_dev/_end; resident batches
 * (batch_upload/_run) are refused.
 */
#define NXSB_MAX_SEGMENTS	8
int		nxsb_engine_segment_add(nxsb_engine_t *, const nxsb_shard_desc_t *);
int		nxsb_engine_segment_count(const nxsb_engine_t *);
int		nxsb_engine_segments_drop(nxsb_engine_t *);
int		nxsb_engine_set_dead(nxsb_engine_t *, uint32_t segment,
		    const uint64_t *ids, uint32_t n);

/* Shard-local df[t] for t in [0, n_terms) after a load (for the all-reduce). */
int		nxsb_engine_get_df(nxsb_engine_t *, uint32_t *df, uint32_t n_terms);
/* Replace the statistics the scores use (after a cross-shard all-reduce). */
int		nxsb_engine_set_global_stats(nxsb_engine_t *, const uint32_t *df,
		    uint32_t n_terms, uint64_t token_count, uint32_t doc_count);

typedef struct nxsb_query {
	uint32_t	tok_off;	/* first token slot in batch.tokens */
	uint32_t	n_tokens;	/* resolved tokens, token-list order */
	uint32_t	prog_off;	/* first op in batch.prog */
	uint32_t	n_prog;
} nxsb_query_t;

typedef struct nxsb_batch {
	int			algo;		/* NXSB_ALGO_* */
	uint32_t		limit;		/* top-N per query, >= 1 */
	uint32_t		n_queries;
	const nxsb_query_t *	queries;
	const uint32_t *	tokens;		/* 1-based term ids */
	uint32_t		n_tokens;
	const int32_t *		prog;
	uint32_t		n_prog;
} nxsb_batch_t;

/*
 * Score a batch.  Host buffers: counts[n_queries], ids/scores
 * [n_queries * limit] (query q's results at [q*limit, q*limit + counts[q]),
 * descending score, ties by descending document id).  Synchronous.
 */
int		nxsb_engine_search(nxsb_engine_t *, const nxsb_batch_t *,
		    uint32_t *counts, uint64_t *ids, float *scores);

/*
 * The same search in two halves: begin enqueues the descriptor copy, the
 * kernels and the result copy and returns a handle >= 0 without waiting; end
 * waits for that batch only and fills the host arrays (pass counts = NULL to
 * discard).  Up to 4 searches may be in flight, so a caller can prepare batch
 * i+1 on the host while batch i is on the device.
 */
int		nxsb_engine_search_begin(nxsb_engine_t *, const nxsb_batch_t *);
int		nxsb_engine_search_end(nxsb_engine_t *, int handle,
		    uint32_t *counts, uint64_t *ids, float *scores);
/*
 * begin for a batch some of whose tokens still need a fuzzy lookup (the
 * reference resolves them one by one before it scores, ref src/index/
 * idxterm.c:210-249 from src/core/tokenizer.c): miss i is the byte string
 * miss_blob[miss_off[i] .. miss_off[i+1]) and its answer -- the term
 * nxsb_engine_fuzzy would return, 0 = none = an empty list -- replaces
 * batch.tokens[miss_pos[i]] ON THE DEVICE, between the descriptor copy and
 * the scoring kernels of the same stream: the host never waits for the scan.
 * (Replicated engines and segmented images run the lookups first instead.)
 */
int		nxsb_engine_search_begin_fz(nxsb_engine_t *, const nxsb_batch_t *,
		    uint32_t n_miss, const char *miss_blob, const uint32_t *miss_off,
		    const uint32_t *miss_pos);
/*
 * begin for a sharded index: this shard's 16-byte records (see batch_run) are
 * written to d_recs, device memory owned by the caller -- the send buffer of
 * the all-gather -- and nothing is copied to the host.  End it with
 * counts = NULL.
 */
int		nxsb_engine_search_begin_dev(nxsb_engine_t *, const nxsb_batch_t *,
		    void *d_recs);

/*
 * The same in three steps, for callers that keep the batch resident in HBM
 * (the `value` leg of bench.py) or chain a collective on the device.
 * upload: copies the batch descriptors to the device, returns a handle >= 0.
 * run:    enqueues the scoring + top-k kernels on the engine's stream; the
 *         results stay in device memory: 16-byte records
 *         { u64 doc_id; f32 score; u32 valid } at recs[q*limit + r].
 *         If d_recs is NULL an engine-owned buffer is used.
 * fetch:  copies an engine-owned result buffer to host arrays.
 */
int		nxsb_engine_batch_upload(nxsb_engine_t *, const nxsb_batch_t *);
int		nxsb_engine_batch_run(nxsb_engine_t *, int handle, void *d_recs);
int		nxsb_engine_batch_fetch(nxsb_engine_t *, int handle,
		    uint32_t *counts, uint64_t *ids, float *scores);
int		nxsb_engine_batch_release(nxsb_engine_t *, int handle);
/* Algorithmic bytes (8 B x sum of df over resolved tokens) of a batch. */
uint64_t	nxsb_engine_batch_bytes(nxsb_engine_t *, int handle);
int		nxsb_engine_sync(nxsb_engine_t *);

/*
 * Merge per-shard top-k record lists (device memory, as produced by
 * batch_run on each rank and gathered rank-major: d_in[g][q][r]) into the
 * global top-k per query, same record format, on the engine's stream.
 */
int		nxsb_engine_merge_topk(nxsb_engine_t *, const void *d_in,
		    uint32_t n_shards, uint32_t n_queries, uint32_t limit,
		    void *d_out);

/*
 * Vocabulary image for fuzzy matching.  bk_* mirror the reference's BK-tree
 * (built host-side by replaying bktree_insert in term-id order):
 * parent term index (UINT32_MAX for the root), edge label to the parent,
 * and rank in a full breadth-first walk with children by ascending label.
 */
int		nxsb_engine_load_vocab(nxsb_engine_t *, uint32_t n_terms,
		    const char *blob, const uint32_t *term_off,
		    const uint64_t *term_total, const uint32_t *bk_parent,
		    const uint8_t *bk_edge, const uint32_t *bk_rank);

/*
 * The per-term occurrence totals moved (documents added or removed; the
 * vocabulary is the one of the last load_vocab): fuzzy matching only picks
 * terms whose total is non-zero (ref src/index/idxterm.c:239, which reads the
 * mapped counter at search time).
 */
int		nxsb_engine_update_term_totals(nxsb_engine_t *, uint32_t n_terms,
		    const uint64_t *term_total);

/*
 * Fuzzy-resolve n query strings (NUL-free byte strings q[i] = qblob +
 * qoff[i] .. qoff[i+1]) against the whole vocabulary: out_term[i] = the term
 * id the reference's idxterm_fuzzysearch would return (0 = none),
 * out_dist[i] its edit distance, and -- optionally -- out_true[i] = the
 * number of vocabulary terms within distance 2 (the true candidate count).
 */
int		nxsb_engine_fuzzy(nxsb_engine_t *, uint32_t n, const char *qblob,
		    const uint32_t *qoff, uint32_t *out_term, uint32_t *out_dist,
		    uint32_t *out_true);

/*
 * The same lookups with the candidate sets laid open ("report
 * both the reference-compatible candidate set and the true <= 2 set").  For
 * query i, out_n[i] = the number of vocabulary terms within distance 2, and
 * the first min(out_n[i], cap) of them, in the BK-tree's breadth-first order,
 * are at [i * cap ...]: cand_term (1-based id), cand_dist, and cand_flags --
 * bit 0: the reference's pruned walk visits the term, i.e. it is in the list
 * bktree_search returns (ref src/algo/bktree.c:219-275; the flagged entries,
 * in this order, ARE that list); bit 1: the term's total is non-zero (ref
 * src/index/idxterm.c:239).  out_term / out_dist (may be NULL) as in
 * nxsb_engine_fuzzy.  A diagnostic call: it allocates and sorts on the host.
 */
int		nxsb_engine_fuzzy_candidates(nxsb_engine_t *, uint32_t n,
		    const char *qblob, const uint32_t *qoff, uint32_t cap,
		    uint32_t *out_term, uint32_t *out_dist, uint32_t *out_n,
		    uint32_t *cand_term, uint8_t *cand_dist, uint8_t *cand_flags);

/* Milliseconds spent by the last batch_run / fuzzy call per kernel family
 * (CUDA events on the engine's stream); names[i] are static strings.
 * Returns the number of entries written (<= cap). */
int		nxsb_engine_last_timings(nxsb_engine_t *, const char **names,
		    float *ms, int cap);
/* The same, summed over the last `last_runs` (<= 256) batch_run calls, so a
 * caller can enqueue many runs back to back and read the per-kernel device
 * time afterwards without synchronising in between. */
int		nxsb_engine_timings(nxsb_engine_t *, uint32_t last_runs,
		    const char **names, float *ms, int cap);
/* Kernel launches issued by the engine since creation. */
uint64_t	nxsb_engine_launch_count(const nxsb_engine_t *);
/*
 * Device / pinned-host allocations and frees made by the library since it
 * was loaded.  cudaFree synchronises the device, so the count must not move
 * while batches of one shape are searched (tests/test_gpu_engine.py).
 */
uint64_t	nxsb_alloc_events(void);

/*
 * Exact top-k pruning (bmw.cuh).  OR queries with limit <= 128 skip the
 * blocks of documents whose score bound -- the sum of the per-(term, block)
 * score maxima kept in the image -- cannot enter the query's current top-k;
 * results are bit-identical to scoring everything, which is what the
 * reference does (ref src/query/search.c:235-272).  On by default
 * (environment NXSB_BMW=0 turns it off at engine creation); set_pruning
 * switches it for the batches staged afterwards and returns the old setting.
 * pruning_stats: out[0..3] = { (query, chunk) items, blocks scored, postings
 * scored, selection rounds } since creation or the last reset; out[4..15] are
 * per-phase cycle counters of a -DBMW_PROF build, else 0.
 */
int		nxsb_engine_set_pruning(nxsb_engine_t *, int on);
/*
 * Threshold priming: the image keeps, per term, the k-th largest
 * query-independent weight (BM25 tf-normalisation or the TF-IDF tf weight)
 * for k = 1, 2, 4, 10, 20, 50, 100, 128; a query's pruning threshold starts
 * at the largest (k-th weight x idf) over its terms instead of zero -- at
 * least k documents score that much.  term_kth copies the NXSB_KTH_STEPS
 * values of n terms (1-based ids) to out[n][NXSB_KTH_STEPS]; 0 = the list has
 * fewer postings than that step (or the id is not a term).
 */
/*
 * Diagnostic: the score of n (term count, document length, idf) triples through
 * the kernels' own arithmetic (st_score: fp32, log table, rcp.approx), with the
 * K0 / K1 of the loaded image -- what scripts/bm25_deviation.py compares with
 * the reference's fp64 evaluation (ref src/algo/ranking.c:135-176).
 * tf < 65536, dl < 65536 (the packed image's ranges).
 */
int		nxsb_engine_score_pairs(nxsb_engine_t *, int algo, uint32_t n,
		    const uint32_t *tf, const uint32_t *dl, const float *idf, float *out);

#define NXSB_KTH_STEPS	8
int		nxsb_engine_term_kth(nxsb_engine_t *, int algo, const uint32_t *term_ids,
		    uint32_t n, float *out);
int		nxsb_engine_pruning_stats(nxsb_engine_t *, uint64_t out[16], int reset);

#pragma GCC visibility pop

#ifdef __cplusplus
}
#endif

#endif
