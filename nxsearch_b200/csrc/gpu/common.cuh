/*
 * Shared device-side definitions of the sm_100a scoring engine.
 *
 * HBM layout of one shard image (see DESIGN.md "Data layout"):
 *
 *   post[P]        uint2 {doc, tfdl}: postings, term-major, ascending dense
 *                  doc index within a term.  PACKED mode: tfdl = tf | dl<<16
 *                  (both < 65535) so one 8-byte load carries everything BM25
 *                  needs; WIDE mode (a tf or doc length >= 65535 exists):
 *                  tfdl = tf and the length is gathered from doc_len[].
 *   term_off[V+1]  u64 CSR offsets into post[]
 *   df[V]          u32 GLOBAL document frequency (all shards)
 *   idf_bm25[V], idf_tfidf[V]  f32, computed on the host in double exactly
 *                  as ref ranking.c:91,172 does
 *   skip_row[V]    i32: row in skip[] for terms with df_local >= DF_LONG
 *   skip[R][T+1]   u32: first posting (relative to the term) whose doc lies
 *                  in tile >= j -- the per-tile slice boundaries
 *   doc_ids[N]     u64 external ids, ascending;  doc_len[N] u32
 *   dense[D][T*16384]  u32: for the D terms present in most documents
 *                  (df_local >= dense_min * N) the same list once more as a
 *                  column indexed by document -- the packed tf|dl word, 0 where
 *                  the term is absent.  The dense encoding of a dense list
 *                  (what roaring's bitmap containers are to the reference):
 *                  4 bytes per document instead of 8 per posting, and the
 *                  document index is implicit, so the scorer's accumulator
 *                  accesses are 16-byte vectors without bank conflicts.
 *                  PACKED mode only.  dense_col[V] i32 maps term -> column.
 */
#ifndef NXSB_GPU_COMMON_CUH
#define NXSB_GPU_COMMON_CUH

#include <cstdint>
#include <cuda_runtime.h>

#include "nxsb200_gpu.h"

#define TILE_DOCS	NXSB_TILE_DOCS		// 16384
#ifndef TILE_SHIFT
#define TILE_SHIFT	14
#endif
#define TILE_WORDS	(TILE_DOCS / 32)
#define DF_LONG		2048u			// permanent skip row threshold
#define MT_SHIFT	10			// mini-tiles: finer rows of the long lists
#define LOGTAB_N	256			// (float)log(c + 1) for c < 256
#define SMALL_K_MAX	2048u			// optimised top-k path limit
#define SORT_CAP	4096u			// keys sorted in shared memory

static_assert((1u << TILE_SHIFT) == TILE_DOCS, "tile size");

#define DENSE_NONE	(~0ull)

/* A resolved token instance of a batch: the result of the term lookup. */
struct DTok {
	unsigned long long	post_off;	// first posting of the term
	const uint32_t *	skip;		// [ntiles + 1] slice boundaries
	uint32_t		df_local;
	float			idf;
	unsigned long long	dense_off;	// word offset of its dense column, or DENSE_NONE
	/*
	 * Finer slice boundaries for the block scorer (bmw.cuh): per mini-tile of
	 * 2^MT_SHIFT documents for the long lists, else `skip` again (tiles).
	 */
	const uint32_t *	fine;
	uint32_t		fine_shift;
	uint32_t		bcol;		// row of its block arrays (bmw.cuh), or 0xffffffff
	float			wmax;		// the list's largest weight
	float			wk;		// its k-th largest, k = the batch's ladder step (bmw.cuh); 0 = none
	const uint8_t *		mtmax;		// long list without block arrays: bytes per mini-tile (bmw.cuh)
	const uint32_t *	mtbits;		// ... and which of the mini-tile's blocks hold a posting
};

/* 16-byte result record (also the NCCL all-gather payload). */
struct Rec {
	unsigned long long	doc_id;
	float			score;
	uint32_t		valid;
};

static_assert(sizeof(Rec) == 16, "record size");

__device__ __forceinline__ unsigned long long
make_key(float score, uint32_t doc)
{
	return ((unsigned long long)__float_as_uint(score) << 32) | doc;
}

#endif
