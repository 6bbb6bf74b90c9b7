/*
 * Index files and their in-memory view; see index.h.
 */
#define _GNU_SOURCE
#include <sys/file.h>
#include <sys/mman.h>
#include <sys/stat.h>

#include <errno.h>
#include <fcntl.h>
#include <inttypes.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "index.h"
#include "be.h"

/*
 * File mapping (ref src/index/idxmap.c): created with O_EXCL under an
 * exclusive lock, sized in 32 KB steps, fully re-mapped on growth.
 */

static int
idxfile_open(idxfile_t *f, const char *path, bool *created)
{
	struct stat st;
	int fd, retry = 10;

	*created = false;
again:
	fd = open(path, O_RDWR | O_CLOEXEC);
	if (fd == -1 && errno == ENOENT) {
		fd = open(path, O_RDWR | O_CREAT | O_EXCL | O_CLOEXEC, 0644);
		if (fd == -1 && errno == EEXIST)
			goto again;
		if (fd == -1)
			return -1;
		if (flock(fd, LOCK_EX) == -1 || ftruncate(fd, IDX_SIZE_STEP) == -1)
			goto err;
		*created = true;
	} else if (fd == -1) {
		return -1;
	} else if (flock(fd, LOCK_SH) == -1) {
		goto err;
	}
	if (fstat(fd, &st) == -1)
		goto err;
	if (st.st_size == 0) {
		/* Opened while the creator has not sized it yet: retry. */
		flock(fd, LOCK_UN);
		close(fd);
		if (--retry == 0) {
			errno = EIO;
			return -1;
		}
		usleep(1000);
		goto again;
	}
	f->fd = fd;
	return 0;
err:
	close(fd);
	return -1;
}

static uint8_t *
idxfile_map(idxfile_t *f, size_t target_len, bool extend)
{
	const size_t file_len = (target_len + IDX_SIZE_STEP - 1) /
	    IDX_SIZE_STEP * IDX_SIZE_STEP;
	struct stat st;
	void *addr;

	if (file_len <= f->mapped_len)
		return f->base;
	if (fstat(f->fd, &st) == -1)
		return NULL;
	if (file_len > (size_t)st.st_size) {
		if (!extend) {
			errno = EINVAL;
			return NULL;
		}
		if (ftruncate(f->fd, file_len) == -1)
			return NULL;
	}
	addr = mmap(NULL, file_len, PROT_READ | PROT_WRITE, MAP_SHARED, f->fd, 0);
	if (addr == MAP_FAILED)
		return NULL;
	if (f->base)
		munmap(f->base, f->mapped_len);
	f->base = addr;
	f->mapped_len = file_len;
	return f->base;
}

static void
idxfile_release(idxfile_t *f)
{
	if (f->base)
		munmap(f->base, f->mapped_len);
	if (f->fd > 0)
		close(f->fd);
	memset(f, 0, sizeof(*f));
}

static inline uint64_t
load_acquire64(const uint8_t *p)
{
	return be64toh(__atomic_load_n((const uint64_t *)(const void *)p,
	    __ATOMIC_ACQUIRE));
}

static inline uint32_t
load_acquire32(const uint8_t *p)
{
	return be32toh(__atomic_load_n((const uint32_t *)(const void *)p,
	    __ATOMIC_ACQUIRE));
}

static inline void
store_release64(uint8_t *p, uint64_t v)
{
	__atomic_store_n((uint64_t *)(void *)p, htobe64(v), __ATOMIC_RELEASE);
}

static inline void
store_release32(uint8_t *p, uint32_t v)
{
	__atomic_store_n((uint32_t *)(void *)p, htobe32(v), __ATOMIC_RELEASE);
}

/*
 * Terms (ref src/index/terms.c).  Block: { len u16, bytes, NUL, pad to 8,
 * total u64 }; the term id is the ordinal of the block, from 1.
 */

static inline size_t
term_block_len(size_t len)
{
	return ((2 + len + 1 + 7) & ~(size_t)7) + 8;
}

int
idx_terms_open(nxs_index_t *idx, const char *path)
{
	bool created;

	if (idxfile_open(&idx->tfile, path, &created) == -1) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "could not open terms index");
		return -1;
	}
	if (idxfile_map(&idx->tfile, IDX_SIZE_STEP, false) == NULL) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "terms mapping failed");
		goto err;
	}
	if (created) {
		uint8_t *hdr = idx->tfile.base;

		memset(hdr, 0, TERMS_HDR_LEN);
		memcpy(hdr, "NXS_T", 5);
		hdr[5] = 1;
		store_release32(hdr + 8, 0);
	} else if (memcmp(idx->tfile.base, "NXS_T", 5) != 0) {
		nxs_set_error(idx->nxs, NXS_ERR_FATAL, "corrupted terms index header");
		goto err;
	} else if (idx->tfile.base[5] != 1) {
		nxs_set_error(idx->nxs, NXS_ERR_FATAL,
		    "incompatible nxsearch index version");
		goto err;
	}
	if ((idx->term_map = strmap_create(1024)) == NULL)
		goto err;
	idx->terms_consumed = 0;
	idx->n_terms = 0;
	idx->max_term_len = 0;
	flock(idx->tfile.fd, LOCK_UN);
	return idx_terms_sync(idx);
err:
	flock(idx->tfile.fd, LOCK_UN);
	idxfile_release(&idx->tfile);
	return -1;
}

void
idx_terms_close(nxs_index_t *idx)
{
	strmap_destroy(idx->term_map);
	idx->term_map = NULL;
	free(idx->term_blob);
	free(idx->term_off);
	free(idx->term_total_off);
	idx->term_blob = NULL;
	idx->term_off = idx->term_total_off = NULL;
	idxfile_release(&idx->tfile);
}

/* Register the next term (id = n_terms + 1) in the in-memory tables. */
static int
term_register(nxs_index_t *idx, const char *val, size_t len, size_t total_off)
{
	if (idx->n_terms + 2 > idx->terms_cap) {
		const uint32_t ncap = idx->terms_cap ? idx->terms_cap * 2 : 1024;
		uint32_t *no = realloc(idx->term_off, sizeof(uint32_t) * ((size_t)ncap + 1));
		uint32_t *nt;

		if (!no)
			return -1;
		idx->term_off = no;
		nt = realloc(idx->term_total_off, sizeof(uint32_t) * ncap);
		if (!nt)
			return -1;
		idx->term_total_off = nt;
		idx->terms_cap = ncap;
	}
	if (idx->blob_len + len + 1 > idx->blob_cap) {
		size_t ncap = idx->blob_cap ? idx->blob_cap * 2 : 65536;
		char *nb;

		while (ncap < idx->blob_len + len + 1)
			ncap *= 2;
		if ((nb = realloc(idx->term_blob, ncap)) == NULL)
			return -1;
		idx->term_blob = nb;
		idx->blob_cap = ncap;
	}
	if (idx->n_terms == 0)
		idx->term_off[0] = 0;
	memcpy(idx->term_blob + idx->blob_len, val, len);
	idx->blob_len += len;
	idx->term_blob[idx->blob_len] = '\0';
	idx->term_total_off[idx->n_terms] = total_off;
	if (len > idx->max_term_len)
		idx->max_term_len = (uint32_t)len;
	idx->n_terms++;
	idx->term_off[idx->n_terms] = idx->blob_len;
	/*
	 * A duplicate value keeps its first id (ref idxterm.c:166-171) but
	 * still consumes an ordinal, as in the reference (terms.c:404-405).
	 */
	if (strmap_put(idx->term_map, val, len, idx->n_terms, NULL) == -1)
		return -1;
	idx->vocab_dirty = true;
	return 0;
}

int
idx_terms_sync(nxs_index_t *idx)
{
	idxfile_t *f = &idx->tfile;
	size_t seen = load_acquire32(f->base + 8), off, end;

	if (seen == idx->terms_consumed)
		return 0;
	if (idxfile_map(f, TERMS_HDR_LEN + seen, false) == NULL) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "terms mapping failed");
		return -1;
	}
	off = TERMS_HDR_LEN + idx->terms_consumed;
	end = TERMS_HDR_LEN + seen;
	/* A block is at least 16 bytes: size the map once for a large catch-up. */
	if (end - off > (1u << 20))
		(void)strmap_reserve(idx->term_map, strmap_count(idx->term_map) + (end - off) / 16);
	while (off < end) {
		size_t len, blk;

		if (off + 2 > end || (len = be_get16(f->base + off)) == 0 ||
		    off + (blk = term_block_len(len)) > end) {
			nxs_set_error(idx->nxs, NXS_ERR_FATAL, "corrupted terms index");
			return -1;
		}
		if (term_register(idx, (const char *)f->base + off + 2, len,
		    off + blk - 8) == -1) {
			nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "term allocation failed");
			return -1;
		}
		off += blk;
		idx->terms_consumed = off - TERMS_HDR_LEN;
	}
	return 0;
}

uint32_t
idx_term_lookup(const nxs_index_t *idx, const char *val, size_t len)
{
	uint32_t id;
	return strmap_get(idx->term_map, val, len, &id) ? id : 0;
}

uint64_t
idx_term_total(const nxs_index_t *idx, uint32_t term_id)
{
	return load_acquire64(idx->tfile.base + idx->term_total_off[term_id - 1]);
}

static void
term_total_add(nxs_index_t *idx, uint32_t term_id, int64_t delta)
{
	uint64_t *tc = (uint64_t *)(void *)(idx->tfile.base +
	    idx->term_total_off[term_id - 1]);
	uint64_t old = __atomic_load_n(tc, __ATOMIC_RELAXED), nv;

	do {
		const uint64_t cur = be64toh(old);

		if (delta < 0 && cur < (uint64_t)-delta)
			return;		/* never underflow (ref idxterm.c:289-296) */
		nv = htobe64(cur + delta);
	} while (!__atomic_compare_exchange_n(tc, &old, nv, true,
	    __ATOMIC_RELAXED, __ATOMIC_RELAXED));
	idx->totals_dirty = true;
}

/*
 * Append the unresolved tokens of the set as new terms (ref terms.c:155-314)
 * and resolve them.
 */
int
idx_terms_add(nxs_index_t *idx, tokenset_t *ts)
{
	idxfile_t *f = &idx->tfile;
	size_t data_len, append = 0, off;
	uint32_t staged = 0;
	int ret = -1;

	for (uint32_t i = 0; i < ts->count; i++)
		staged += ts->list[i].term_id == 0;
	if (!staged)
		return 0;
	if (flock(f->fd, LOCK_EX) == -1)
		return -1;

	/* Pick up terms other processes appended; some may now resolve. */
	if (idx_terms_sync(idx) == -1)
		goto out;
	data_len = load_acquire32(f->base + 8);
	for (uint32_t i = 0; i < ts->count; i++) {
		token_t *t = &ts->list[i];

		if (t->term_id == 0)
			t->term_id = idx_term_lookup(idx, t->str, t->len);
		if (t->term_id == 0) {
			if (t->len > UINT16_MAX) {
				nxs_set_error(idx->nxs, NXS_ERR_LIMIT,
				    "term too long (%u)", t->len);
				goto out;
			}
			append += term_block_len(t->len);
		}
	}
	if (data_len + append > UINT32_MAX) {
		nxs_set_error(idx->nxs, NXS_ERR_LIMIT, "terms index is full");
		goto out;
	}
	if (idxfile_map(f, TERMS_HDR_LEN + data_len + append, true) == NULL) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "terms mapping failed");
		goto out;
	}
	off = TERMS_HDR_LEN + data_len;
	for (uint32_t i = 0; i < ts->count; i++) {
		token_t *t = &ts->list[i];
		size_t blk;

		if (t->term_id)
			continue;
		if (idx->n_terms == UINT32_MAX) {
			nxs_set_error(idx->nxs, NXS_ERR_LIMIT,
			    "reached the term limit (%u)", UINT32_MAX);
			goto publish;
		}
		blk = term_block_len(t->len);
		memset(f->base + off, 0, blk);
		be_put16(f->base + off, t->len);
		memcpy(f->base + off + 2, t->str, t->len);
		/* Initial total = occurrences in this document (terms.c:259). */
		be_put64(f->base + off + blk - 8, t->count);
		if (term_register(idx, t->str, t->len, off + blk - 8) == -1) {
			nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "term allocation failed");
			goto publish;
		}
		t->term_id = idx->n_terms;
		off += blk;
	}
	ret = 0;
publish:
	idx->terms_consumed = off - TERMS_HDR_LEN;
	store_release32(f->base + 8, idx->terms_consumed);
out:
	flock(f->fd, LOCK_UN);
	return ret;
}

/*
 * Document-term map (ref src/index/dtmap.c).  Block: { doc_id u64,
 * doc_len u32, n u32, n x (term_id u32, count u32) } sorted by term id.
 */

int
idx_dtmap_open(nxs_index_t *idx, const char *path)
{
	bool created;

	if (idxfile_open(&idx->dfile, path, &created) == -1) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "could not open dtmap index");
		return -1;
	}
	if (idxfile_map(&idx->dfile, IDX_SIZE_STEP, false) == NULL) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "dtmap mapping failed");
		goto err;
	}
	if (created) {
		uint8_t *hdr = idx->dfile.base;

		memset(hdr, 0, DTMAP_HDR_LEN);
		memcpy(hdr, "NXS_D", 5);
		hdr[5] = 1;
		store_release64(hdr + 8, 0);
	} else if (memcmp(idx->dfile.base, "NXS_D", 5) != 0) {
		nxs_set_error(idx->nxs, NXS_ERR_FATAL, "corrupted dtmap index header");
		goto err;
	} else if (idx->dfile.base[5] != 1) {
		nxs_set_error(idx->nxs, NXS_ERR_FATAL,
		    "incompatible nxsearch index version");
		goto err;
	}
	if ((idx->doc_map = u64map_create(1024)) == NULL)
		goto err;
	idx->doc_map_ready = true;
	idx->dt_consumed = 0;
	flock(idx->dfile.fd, LOCK_UN);
	return idx_dtmap_sync(idx, true);
err:
	flock(idx->dfile.fd, LOCK_UN);
	idxfile_release(&idx->dfile);
	return -1;
}

void
idx_dtmap_close(nxs_index_t *idx)
{
	u64map_destroy(idx->doc_map);
	idx->doc_map = NULL;
	free(idx->doc_ids);
	free(idx->doc_len);
	free(idx->doc_n);
	free(idx->doc_blk);
	free(idx->doc_dead);
	free(idx->doc_seg);
	free(idx->df);
	idx->doc_seg = NULL;
	idx->df = NULL;
	idx->df_cap = 0;
	for (unsigned g = 0; g <= NXSB_MAX_SEGMENTS; g++) {
		free(idx->seg_dead[g]);
		idx->seg_dead[g] = NULL;
		idx->seg_ndead[g] = idx->seg_dead_cap[g] = 0;
	}
	idx->doc_ids = NULL;
	idx->doc_len = idx->doc_n = NULL;
	idx->doc_blk = NULL;
	idx->doc_dead = NULL;
	idxfile_release(&idx->dfile);
}

uint64_t
idx_get_token_count(const nxs_index_t *idx)
{
	return load_acquire64(idx->dfile.base + 16);
}

uint32_t
idx_get_doc_count(const nxs_index_t *idx)
{
	return load_acquire32(idx->dfile.base + 24);
}

/* df[] covers every known term (new terms start at zero). */
static int
df_reserve(nxs_index_t *idx)
{
	if (idx->n_terms > idx->df_cap) {
		uint32_t ncap = idx->df_cap ? idx->df_cap : 1024;
		uint32_t *p;

		while (ncap < idx->n_terms)
			ncap *= 2;
		if ((p = realloc(idx->df, sizeof(uint32_t) * ncap)) == NULL)
			return -1;
		memset(p + idx->df_cap, 0, sizeof(uint32_t) * (ncap - idx->df_cap));
		idx->df = p;
		idx->df_cap = ncap;
	}
	return 0;
}

/* Add (+1) or retire (-1) the terms of the block at file offset blk in df[]. */
static void
df_apply(nxs_index_t *idx, uint64_t blk, uint32_t n, int sign)
{
	const uint8_t *p = idx->dfile.base + blk + 16;

	for (uint32_t j = 0; j < n; j++) {
		const uint32_t t = be_get32(p + (size_t)j * 8);

		if (t >= 1 && t <= idx->n_terms)
			idx->df[t - 1] += sign;
	}
}

static int
doc_register_opt(nxs_index_t *idx, uint64_t id, uint32_t len, uint32_t n,
    uint64_t blk, bool count_df)
{
	uint32_t slot = idx->n_slots;

	if (slot == idx->slots_cap) {
		const uint32_t ncap = idx->slots_want > slot ? idx->slots_want :
		    idx->slots_cap ? idx->slots_cap * 2 : 1024;
		void *p;

#define GROW(field, type) \
		if ((p = realloc(idx->field, sizeof(type) * ncap)) == NULL) \
			return -1; \
		idx->field = p;
		GROW(doc_ids, uint64_t)
		GROW(doc_len, uint32_t)
		GROW(doc_n, uint32_t)
		GROW(doc_blk, uint64_t)
		GROW(doc_dead, uint8_t)
		GROW(doc_seg, uint8_t)
#undef GROW
		idx->slots_cap = ncap;
	}
	if (df_reserve(idx) == -1)
		return -1;
	if (idx->doc_map_ready && u64map_put(idx->doc_map, id, slot, NULL) != 1) {
		errno = EEXIST;
		return -1;
	}
	idx->doc_ids[slot] = id;
	idx->doc_len[slot] = len;
	idx->doc_n[slot] = n;
	idx->doc_blk[slot] = blk;
	idx->doc_dead[slot] = 0;
	idx->doc_seg[slot] = 0;
	idx->n_slots++;
	idx->n_live++;
	if (count_df)
		df_apply(idx, blk, n, +1);
	/* Not on the GPU yet: the next search adds a delta segment. */
	idx->n_pending++;
	idx->stats_dirty = true;
	idx->totals_dirty = true;
	return 0;
}

static int
doc_register(nxs_index_t *idx, uint64_t id, uint32_t len, uint32_t n,
    uint64_t blk)
{
	return doc_register_opt(idx, id, len, n, blk, true);
}

/*
 * A search-only process never looks a document up by id, so a bulk open
 * (dtmap_sync_bulk) does not fill the map; the first add / remove / deletion
 * marker does, from the per-slot arrays.
 */
int
idx_docmap_ensure(nxs_index_t *idx)
{
	if (idx->doc_map_ready)
		return 0;
	if (u64map_reserve(idx->doc_map, idx->n_live) == -1)
		return -1;
	for (uint32_t s = 0; s < idx->n_slots; s++) {
		if (idx->doc_dead[s])
			continue;
		if (s + 16 < idx->n_slots)
			u64map_prefetch(idx->doc_map, idx->doc_ids[s + 16]);
		if (u64map_put(idx->doc_map, idx->doc_ids[s], s, NULL) != 1) {
			errno = EEXIST;		/* one id in two live blocks */
			return -1;
		}
	}
	idx->doc_map_ready = true;
	return 0;
}

static void
doc_unregister(nxs_index_t *idx, uint64_t id)
{
	uint32_t slot;

	if (idx_docmap_ensure(idx) == -1) {
		idx->image_dirty = true;	/* cannot tell which slot: rebuild from the files' view */
		return;
	}
	if (!u64map_get(idx->doc_map, id, &slot))
		return;
	u64map_del(idx->doc_map, id);
	idx->doc_dead[slot] = 1;
	idx->n_live--;
	/* Only the id of a removed block is cleared; its terms stay readable. */
	df_apply(idx, idx->doc_blk[slot], idx->doc_n[slot], -1);
	idx->stats_dirty = true;
	idx->totals_dirty = true;
	if (slot >= idx->built_slots) {
		idx->n_pending--;
	} else {
		/* In the image: note it against the segment that holds it. */
		const uint32_t g = idx->doc_seg[slot];

		if (idx->seg_ndead[g] == idx->seg_dead_cap[g]) {
			const uint32_t ncap = idx->seg_dead_cap[g] ? idx->seg_dead_cap[g] * 2 : 16;
			uint64_t *p = realloc(idx->seg_dead[g], sizeof(uint64_t) * ncap);

			if (p == NULL) {
				idx->image_dirty = true;	/* fall back to a rebuild */
				return;
			}
			idx->seg_dead[g] = p;
			idx->seg_dead_cap[g] = ncap;
		}
		idx->seg_dead[g][idx->seg_ndead[g]++] = id;
		idx->seg_dead_dirty[g] = true;
		idx->seg_live[g]--;
	}
}

/*
 * Bulk catch-up (opening a large index): the per-posting work of a sync --
 * checking every term id and counting df[] -- is spread over host threads.
 * The block chain itself (each block's length is in its header) is walked
 * once, serially; documents are then registered in file order as always.
 * Returns 1 when the range was consumed, 0 when the caller should take the
 * block-by-block path instead (small range, unknown term, no memory), -1 on
 * a corrupt file.
 */
#define BULK_MIN_BYTES	(64u << 20)
#define BULK_MAX_THREADS 8

typedef struct { uint64_t id; uint32_t dl, n; } bulk_hdr_t;

typedef struct {
	const nxs_index_t *	idx;
	const uint64_t *	blk;		/* block offsets */
	bulk_hdr_t *		hdr;		/* decoded block headers (filled here) */
	size_t			lo, hi;		/* blocks [lo, hi) */
	uint32_t *		df;		/* private counts */
	bool			bad;		/* a term id outside the vocabulary */
} bulk_job_t;

static void *
bulk_worker(void *arg)
{
	bulk_job_t *job = arg;
	const uint8_t *base = job->idx->dfile.base;
	const uint32_t n_terms = job->idx->n_terms;

	for (size_t i = job->lo; i < job->hi; i++) {
		const uint8_t *p = base + job->blk[i];
		const uint32_t n = be_get32(p + 12);

		job->hdr[i] = (bulk_hdr_t){ be_get64(p), be_get32(p + 8), n };
		if (job->hdr[i].id == 0 || job->hdr[i].dl == 0)
			continue;		/* deleted in place / deletion marker */
		for (uint32_t j = 0; j < n; j++) {
			const uint32_t t = be_get32(p + 16 + (size_t)j * 8);

			if (t == 0 || t > n_terms) {
				job->bad = true;
				return NULL;
			}
			job->df[t - 1]++;
		}
	}
	return NULL;
}

static int
dtmap_sync_bulk(nxs_index_t *idx, size_t off, size_t end)
{
	idxfile_t *f = &idx->dfile;
	long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
	unsigned nthr = ncpu < 1 ? 1 : ncpu > BULK_MAX_THREADS ? BULK_MAX_THREADS : (unsigned)ncpu;
	bulk_job_t jobs[BULK_MAX_THREADS] = { { 0 } };
	pthread_t tids[BULK_MAX_THREADS];
	/* Test hook: the size from which a sync takes this path. */
	const char *min_env = getenv("NXSB_BULK_MIN_BYTES");
	const size_t min_bytes = min_env ? (size_t)strtoull(min_env, NULL, 10) : BULK_MIN_BYTES;
	uint64_t *blk = NULL;
	bulk_hdr_t *hdr = NULL;
	size_t nblk = 0, cap = (end - off) / 256 + 1024;
	int ret = 0;

	if (end - off < min_bytes || nthr < 2 || df_reserve(idx) == -1)
		return 0;
#ifdef MADV_POPULATE_READ
	/* Fault the range in with one call instead of once per page. */
	(void)madvise(f->base + (off & ~(size_t)4095), end - (off & ~(size_t)4095),
	    MADV_POPULATE_READ);
#endif
	struct timespec ts0, ts1, ts2, ts3;
	const bool prof = getenv("NXSB_OPEN_PROF") != NULL;

	clock_gettime(CLOCK_MONOTONIC, &ts0);
	if ((blk = malloc(sizeof(uint64_t) * cap)) == NULL)
		return 0;
	/*
	 * Each header is found through the previous one, a chain of dependent
	 * loads ~450 bytes apart that the hardware prefetcher does not follow:
	 * keep every cache line of the next few KB on its way.
	 */
	size_t pf = off;
	for (size_t o = off; o < end;) {
		uint32_t n;

		for (; pf < o + 4096 && pf < end; pf += 64)
			__builtin_prefetch(f->base + pf, 0, 0);
		if (o + 16 > end)
			goto corrupt;
		n = be_get32(f->base + o + 12);
		if (o + 16 + (size_t)n * 8 > end)
			goto corrupt;
		if (nblk == cap) {
			uint64_t *nb = realloc(blk, sizeof(uint64_t) * cap * 2);

			if (nb == NULL)
				goto out;
			blk = nb;
			cap *= 2;
		}
		blk[nblk++] = o;
		o += 16 + (size_t)n * 8;
	}

	clock_gettime(CLOCK_MONOTONIC, &ts1);
	if ((hdr = malloc(sizeof(bulk_hdr_t) * (nblk + 1))) == NULL)
		goto out;
	for (unsigned t = 0; t < nthr; t++) {
		jobs[t].idx = idx;
		jobs[t].hdr = hdr;
		jobs[t].blk = blk;
		jobs[t].lo = nblk * t / nthr;
		jobs[t].hi = nblk * (t + 1) / nthr;
		if ((jobs[t].df = calloc(idx->n_terms ? idx->n_terms : 1, sizeof(uint32_t))) == NULL)
			nthr = t;		/* fewer workers: re-split below */
	}
	if (nthr < 2)
		goto out;
	for (unsigned t = 0; t < nthr; t++) {
		jobs[t].lo = nblk * t / nthr;
		jobs[t].hi = nblk * (t + 1) / nthr;
	}
	{
		unsigned started = 0;
		bool bad = false;

		for (; started < nthr; started++)
			if (pthread_create(&tids[started], NULL, bulk_worker, &jobs[started]) != 0)
				break;
		for (unsigned t = started; t < nthr; t++)
			bulk_worker(&jobs[t]);		/* could not start: do it here */
		for (unsigned t = 0; t < started; t++)
			pthread_join(tids[t], NULL);
		for (unsigned t = 0; t < nthr; t++)
			bad |= jobs[t].bad;
		if (bad)
			goto out;	/* the serial path reports or defers it */
	}
	for (unsigned t = 0; t < nthr; t++)
		for (uint32_t i = 0; i < idx->n_terms; i++)
			idx->df[i] += jobs[t].df[i];
	clock_gettime(CLOCK_MONOTONIC, &ts2);

	/* The tables grow once, and the document map's slots are fetched ahead. */
	if ((uint64_t)idx->n_slots + nblk < UINT32_MAX)
		idx->slots_want = idx->n_slots + (uint32_t)nblk;
	if (idx->n_slots == 0)
		idx->doc_map_ready = false;	/* a fresh open: left to idx_docmap_ensure */
	else if (idx->doc_map_ready)
		(void)u64map_reserve(idx->doc_map, u64map_count(idx->doc_map) + nblk);
	for (size_t i = 0; i < nblk; i++) {
		const uint64_t id = hdr[i].id;
		const uint32_t dl = hdr[i].dl, n = hdr[i].n;

		if (idx->doc_map_ready && i + 16 < nblk)
			u64map_prefetch(idx->doc_map, hdr[i + 16].id);
		if (id == 0) {
			/* Deleted in place (dtmap.c:360-368): skip. */
		} else if (dl == 0) {
			doc_unregister(idx, id);
		} else if (doc_register_opt(idx, id, dl, n, blk[i], false) == -1) {
			/*
			 * df[] already holds the rest of the range: undo what
			 * was not registered, then fail as the serial path does.
			 */
			for (size_t k = i; k < nblk; k++)
				if (hdr[k].id != 0 && hdr[k].dl != 0)
					df_apply(idx, blk[k], hdr[k].n, -1);
			nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM,
			    "document registration failed");
			ret = -1;
			goto out;
		}
		idx->dt_consumed = blk[i] + 16 + (size_t)n * 8 - DTMAP_HDR_LEN;
	}
	ret = 1;
	idx->slots_want = 0;
	if (prof) {
		clock_gettime(CLOCK_MONOTONIC, &ts3);
#define DT(a, b) ((b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec))
		fprintf(stderr, "dtmap_sync_bulk: %zu blocks, %u threads: chain %.2fs, "
		    "terms+df %.2fs, register %.2fs\n", nblk, nthr, DT(ts0, ts1),
		    DT(ts1, ts2), DT(ts2, ts3));
#undef DT
	}
out:
	for (unsigned t = 0; t < BULK_MAX_THREADS; t++)
		free(jobs[t].df);
	free(blk);
	free(hdr);
	return ret;
corrupt:
	free(blk);
	nxs_set_error(idx->nxs, NXS_ERR_FATAL, "corrupted dtmap index");
	return -1;
}

int
idx_dtmap_sync(nxs_index_t *idx, bool partial)
{
	idxfile_t *f = &idx->dfile;
	const size_t seen = load_acquire64(f->base + 8);
	size_t off, end;

	if (seen == idx->dt_consumed)
		return 0;
	if (idxfile_map(f, DTMAP_HDR_LEN + seen, false) == NULL) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "dtmap mapping failed");
		return -1;
	}
	off = DTMAP_HDR_LEN + idx->dt_consumed;
	end = DTMAP_HDR_LEN + seen;
	{
		const int rc = dtmap_sync_bulk(idx, off, end);

		if (rc != 0)
			return rc == 1 ? 0 : -1;
	}
	while (off < end) {
		uint64_t id;
		uint32_t dl, n;

		if (off + 16 > end)
			goto corrupt;
		id = be_get64(f->base + off);
		dl = be_get32(f->base + off + 8);
		n = be_get32(f->base + off + 12);
		if (off + 16 + (size_t)n * 8 > end)
			goto corrupt;

		if (id == 0) {
			/* Deleted in place (dtmap.c:360-368): skip. */
		} else if (dl == 0) {
			/* Deletion marker (dtmap.c:370-381). */
			doc_unregister(idx, id);
		} else {
			/* Every referenced term must be known (dtmap.c:404-412). */
			for (uint32_t j = 0; j < n; j++) {
				const uint32_t t = be_get32(f->base + off + 16 + (size_t)j * 8);

				if (t == 0 || t > idx->n_terms) {
					if (partial)
						return 0;	/* retry on the next sync */
					nxs_set_error(idx->nxs, NXS_ERR_FATAL,
					    "dtmap refers to unknown term %u", t);
					return -1;
				}
			}
			if (doc_register(idx, id, dl, n, off) == -1) {
				nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM,
				    "document registration failed");
				return -1;
			}
		}
		off += 16 + (size_t)n * 8;
		idx->dt_consumed = off - DTMAP_HDR_LEN;
	}
	return 0;
corrupt:
	nxs_set_error(idx->nxs, NXS_ERR_FATAL, "corrupted dtmap index");
	return -1;
}

static int
pair_cmp(const void *a, const void *b)
{
	const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
	return (x > y) - (x < y);
}

int
idx_dtmap_add(nxs_index_t *idx, nxs_doc_id_t doc_id, tokenset_t *ts)
{
	idxfile_t *f = &idx->dfile;
	const size_t blk_len = 16 + (size_t)ts->count * 8;
	uint32_t *pairs = malloc(sizeof(uint32_t) * 2 * ts->count);
	size_t data_len;
	uint8_t *p;
	int ret = -1;

	if (!pairs)
		return -1;
	for (uint32_t i = 0; i < ts->count; i++) {
		pairs[2 * i] = ts->list[i].term_id;
		pairs[2 * i + 1] = ts->list[i].count;
		/* Totals grow as the block is built (dtmap.c:232). */
		term_total_add(idx, ts->list[i].term_id, ts->list[i].count);
	}
	qsort(pairs, ts->count, 8, pair_cmp);

	if (idx_dtmap_sync(idx, true) == -1 || flock(f->fd, LOCK_EX) == -1)
		goto revert;
	for (;;) {
		data_len = load_acquire64(f->base + 8);
		if (idx->dt_consumed >= data_len)
			break;
		if (idx_terms_sync(idx) == -1 || idx_dtmap_sync(idx, false) == -1)
			goto unlock;
	}
	if (idx_docmap_ensure(idx) == -1) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "document map build failed");
		goto unlock;
	}
	if (u64map_get(idx->doc_map, doc_id, NULL)) {
		nxs_set_error(idx->nxs, NXS_ERR_EXISTS,
		    "document %" PRIu64 " is already indexed", doc_id);
		goto unlock;
	}
	if (idxfile_map(f, DTMAP_HDR_LEN + data_len + blk_len, true) == NULL) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "dtmap mapping failed");
		goto unlock;
	}
	p = f->base + DTMAP_HDR_LEN + data_len;
	be_put64(p, doc_id);
	be_put32(p + 8, ts->seen);
	be_put32(p + 12, ts->count);
	for (uint32_t i = 0; i < ts->count; i++) {
		be_put32(p + 16 + (size_t)i * 8, pairs[2 * i]);
		be_put32(p + 20 + (size_t)i * 8, pairs[2 * i + 1]);
	}
	if (doc_register(idx, doc_id, ts->seen, ts->count,
	    DTMAP_HDR_LEN + data_len) == -1) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "document registration failed");
		goto unlock;
	}

	/* Counters first, then publish the new length (dtmap.c:331-337). */
	idx->dt_consumed = data_len + blk_len;
	be_put64(f->base + 16, idx_get_token_count(idx) + ts->seen);
	be_put32(f->base + 24, idx_get_doc_count(idx) + 1);
	store_release64(f->base + 8, idx->dt_consumed);
	ret = 0;
unlock:
	flock(f->fd, LOCK_UN);
revert:
	if (ret != 0) {
		for (uint32_t i = 0; i < ts->count; i++)
			term_total_add(idx, ts->list[i].term_id,
			    -(int64_t)ts->list[i].count);
	}
	free(pairs);
	return ret;
}

int
idx_dtmap_remove(nxs_index_t *idx, nxs_doc_id_t doc_id)
{
	idxfile_t *f = &idx->dfile;
	size_t data_len;
	uint32_t slot, seen, n;
	uint8_t *blk, *p;
	int ret = -1;

	if (flock(f->fd, LOCK_EX) == -1)
		return -1;
	if (idx_terms_sync(idx) == -1 || idx_dtmap_sync(idx, false) == -1)
		goto out;
	if (idx_docmap_ensure(idx) == -1) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "document map build failed");
		goto out;
	}
	if (!u64map_get(idx->doc_map, doc_id, &slot)) {
		nxs_set_error(idx->nxs, NXS_ERR_MISSING,
		    "document %" PRIu64 " not found", doc_id);
		goto out;
	}
	data_len = load_acquire64(f->base + 8);
	if (idxfile_map(f, DTMAP_HDR_LEN + data_len + 16, true) == NULL) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "dtmap mapping failed");
		goto out;
	}

	/* Invalidate the block for fresh readers (dtmap.c:611-616). */
	blk = f->base + idx->doc_blk[slot];
	store_release64(blk, 0);
	seen = be_get32(blk + 8);
	n = be_get32(blk + 12);
	for (uint32_t j = 0; j < n; j++) {
		const uint32_t t = be_get32(blk + 16 + (size_t)j * 8);
		const uint32_t c = be_get32(blk + 20 + (size_t)j * 8);

		if (t >= 1 && t <= idx->n_terms)
			term_total_add(idx, t, -(int64_t)c);
	}

	/* Marker for the active readers, then the counters (dtmap.c:636-652). */
	p = f->base + DTMAP_HDR_LEN + data_len;
	be_put64(p, doc_id);
	be_put64(p + 8, 0);
	doc_unregister(idx, doc_id);
	be_put32(f->base + 24, idx_get_doc_count(idx) - 1);
	be_put64(f->base + 16, idx_get_token_count(idx) - seen);
	idx->dt_consumed = data_len + 16;
	store_release64(f->base + 8, idx->dt_consumed);
	ret = 0;
out:
	flock(f->fd, LOCK_UN);
	return ret;
}

/*
 * GPU image.  Live documents are handed to the engine in ascending id order
 * -- the order the reference's roaring64 iteration visits them in
 * (search.c:235) and the basis of its tie behaviour.
 *
 * The reference folds every appended block into its in-memory index on the
 * next search (dtmap.c:357-441).  Here the index lives in HBM in term-major
 * form, which cannot be appended to in place, so the image is kept as a base
 * segment plus a few small delta segments:
 *
 *   - documents appended since the last build become a new delta segment;
 *   - a removed document stays in its segment and is listed as dead; every
 *     segment is then asked for limit + (dead documents) results and the
 *     dead ones are dropped while the per-segment lists are merged;
 *   - df[], N and the token count are whole-index values held on the host
 *     and re-sent when they move, so scores equal those of a rebuilt image;
 *   - when the delta segments run out (NXSB_MAX_SEGMENTS) or one collects
 *     too many dead documents they are consolidated into one; when the
 *     deltas outgrow a fraction of the base, or the base collects too many
 *     dead documents, everything is rebuilt.
 */

#define SEG_DEAD_MAX		64	/* limit + this stays on the fast kernel path */
#define DELTA_DOCS_MIN		65536	/* deltas smaller than this never force a rebuild */
#define DELTA_BASE_FRACTION	8	/* ... nor while below base / 8 */

typedef struct { uint64_t id; uint32_t slot; } idslot_t;

static int
idslot_cmp(const void *a, const void *b)
{
	const idslot_t *x = a, *y = b;
	return (x->id > y->id) - (x->id < y->id);
}

enum { PICK_ALL, PICK_DELTAS, PICK_PENDING };

/*
 * Hand the live documents selected by `pick` to the engine as segment `seg`
 * (0: nxsb_engine_load_shard, else nxsb_engine_segment_add).  Returns the
 * number of documents, or -1.
 */
static int64_t
build_segment(nxs_index_t *idx, int pick, uint32_t seg)
{
	const uint32_t first = pick == PICK_PENDING ? idx->built_slots : 0;
	idslot_t *order = NULL;
	uint64_t *ids = NULL, *offs = NULL;
	uint32_t *lens = NULL, *pairs = NULL, *raw_n = NULL;
	const bool host_pairs = getenv("NXSB_IMAGE_HOST_PAIRS") != NULL;
	uint64_t np = 0;
	uint32_t n = 0, k = 0;
	bool sorted = true;
	int64_t ret = -1;

#define PICKED(s) (!idx->doc_dead[s] && (pick != PICK_DELTAS || \
	(s) >= idx->built_slots || idx->doc_seg[s] != 0))
	for (uint32_t s = first; s < idx->n_slots; s++)
		n += PICKED(s);
	if (n == 0 && seg != 0)
		return 0;
	order = malloc(sizeof(idslot_t) * ((size_t)n + 1));
	ids = malloc(sizeof(uint64_t) * ((size_t)n + 1));
	lens = malloc(sizeof(uint32_t) * ((size_t)n + 1));
	offs = malloc(sizeof(uint64_t) * ((size_t)n + 1));
	if (!order || !ids || !lens || !offs)
		goto out;
	for (uint32_t s = first; s < idx->n_slots; s++) {
		if (!PICKED(s))
			continue;
		if (k && idx->doc_ids[s] < order[k - 1].id)
			sorted = false;
		order[k++] = (idslot_t){ idx->doc_ids[s], s };
		np += idx->doc_n[s];
	}
#undef PICKED
	if (!sorted)
		qsort(order, n, sizeof(idslot_t), idslot_cmp);
	/*
	 * The postings go to the device as the file holds them (big-endian
	 * blocks, decoded there: SURVEY 8f N2); offs[] / lens[] double as
	 * raw_off[] / raw_n[].  NXSB_IMAGE_HOST_PAIRS=1 keeps the older
	 * host-side conversion for A/B timing.
	 */
	if (host_pairs) {
		if ((pairs = malloc(np * 8 + 8)) == NULL)
			goto out;
	} else if ((raw_n = malloc(sizeof(uint32_t) * ((size_t)n + 1))) == NULL) {
		goto out;
	}
	np = 0;
	for (uint32_t i = 0; i < n; i++) {
		const uint32_t s = order[i].slot;
		const uint8_t *p = idx->dfile.base + idx->doc_blk[s] + 16;

		ids[i] = order[i].id;
		lens[i] = idx->doc_len[s];
		if (!host_pairs) {
			offs[i] = idx->doc_blk[s] + 16;
			raw_n[i] = idx->doc_n[s];
			continue;
		}
		offs[i] = np;
		for (uint32_t j = 0; j < idx->doc_n[s]; j++) {
			pairs[2 * np] = be_get32(p + (size_t)j * 8);
			pairs[2 * np + 1] = be_get32(p + (size_t)j * 8 + 4);
			np++;
		}
	}
	offs[n] = np;

	const nxsb_shard_desc_t sd = {
		.n_docs = n, .n_terms = idx->n_terms,
		.doc_ids = ids, .doc_len = lens,
		.doc_off = host_pairs ? offs : NULL, .pairs = pairs,
		/* The header counters are what ranking reads (ranking.c:77,163). */
		.token_count = idx_get_token_count(idx),
		.doc_count = idx_get_doc_count(idx),
		.df = idx->df,
		.raw = host_pairs ? NULL : idx->dfile.base,
		.raw_off = host_pairs ? NULL : offs,
		.raw_n = raw_n,
	};
	if ((seg == 0 ? nxsb_engine_load_shard(idx->engine, &sd) :
	    nxsb_engine_segment_add(idx->engine, &sd)) == -1) {
		nxs_set_error(idx->nxs, NXS_ERR_SYSTEM, "GPU index image build failed: %s",
		    nxsb_engine_errmsg(idx->engine));
		goto out;
	}
	for (uint32_t i = 0; i < n; i++)
		idx->doc_seg[order[i].slot] = (uint8_t)seg;
	idx->seg_live[seg] = n;
	ret = n;
out:
	free(order);
	free(ids);
	free(lens);
	free(offs);
	free(pairs);
	free(raw_n);
	return ret;
}

static void
seg_dead_reset(nxs_index_t *idx, uint32_t from)
{
	for (uint32_t g = from; g <= NXSB_MAX_SEGMENTS; g++) {
		idx->seg_ndead[g] = 0;
		idx->seg_dead_dirty[g] = false;
		if (g)
			idx->seg_live[g] = 0;
	}
}

static int
u64_cmp(const void *a, const void *b)
{
	const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
	return (x > y) - (x < y);
}

/* Bring the HBM image in line with what the files say now. */
static int
refresh_image(nxs_index_t *idx)
{
	uint64_t delta_docs = idx->n_pending;
	bool consolidate = false;

	if (df_reserve(idx) == -1) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "out of memory");
		return -1;
	}
	for (uint32_t g = 1; g <= idx->n_segs; g++) {
		delta_docs += idx->seg_live[g];
		consolidate |= idx->seg_ndead[g] > SEG_DEAD_MAX;
	}
	consolidate |= idx->n_pending && idx->n_segs == NXSB_MAX_SEGMENTS;
	if (idx->seg_ndead[0] > SEG_DEAD_MAX || (delta_docs > DELTA_DOCS_MIN &&
	    delta_docs > idx->seg_live[0] / DELTA_BASE_FRACTION) ||
	    getenv("NXSB_REFRESH_FULL") != NULL)
		idx->image_dirty |= idx->n_pending || idx->stats_dirty;

	if (idx->image_dirty) {
		seg_dead_reset(idx, 0);
		if (build_segment(idx, PICK_ALL, 0) == -1)
			return -1;
		idx->n_segs = 0;
		idx->n_full_builds++;
		idx->image_dirty = idx->stats_dirty = false;
	} else if (consolidate) {
		int64_t n;

		if (nxsb_engine_segments_drop(idx->engine) == -1)
			goto fail;
		seg_dead_reset(idx, 1);
		idx->seg_dead_dirty[0] = idx->seg_ndead[0] != 0;
		if ((n = build_segment(idx, PICK_DELTAS, 1)) == -1) {
			idx->image_dirty = true;
			return -1;
		}
		idx->n_segs = n ? 1 : 0;
		idx->n_consolidations++;
	} else if (idx->n_pending) {
		if (build_segment(idx, PICK_PENDING, idx->n_segs + 1) == -1)
			return -1;
		idx->n_segs++;
		idx->n_delta_builds++;
	}
	idx->built_slots = idx->n_slots;
	idx->n_pending = 0;

	for (uint32_t g = 0; g <= idx->n_segs; g++) {
		if (!idx->seg_dead_dirty[g])
			continue;
		qsort(idx->seg_dead[g], idx->seg_ndead[g], sizeof(uint64_t), u64_cmp);
		if (nxsb_engine_set_dead(idx->engine, g, idx->seg_dead[g],
		    idx->seg_ndead[g]) == -1)
			goto fail;
		idx->seg_dead_dirty[g] = false;
	}
	if (idx->stats_dirty) {
		if (nxsb_engine_set_global_stats(idx->engine, idx->df, idx->n_terms,
		    idx_get_token_count(idx), idx_get_doc_count(idx)) == -1)
			goto fail;
		idx->stats_dirty = false;
	}
	return 0;
fail:
	nxs_set_error(idx->nxs, NXS_ERR_SYSTEM, "GPU index image refresh failed: %s",
	    nxsb_engine_errmsg(idx->engine));
	idx->image_dirty = true;
	return -1;
}

/* include/nxsb200_tools.h */
NXS_API uint32_t
nxsb_index_term_df(const void *index, uint32_t term_id)
{
	const nxs_index_t *idx = index;

	return term_id >= 1 && term_id <= idx->n_terms && term_id <= idx->df_cap
	    ? idx->df[term_id - 1] : 0;
}

/* include/nxsb200_tools.h */
NXS_API void
nxsb_index_image_stats(const void *index, uint64_t out[7])
{
	const nxs_index_t *idx = index;
	uint64_t dead = 0;

	for (uint32_t g = 0; g <= idx->n_segs; g++)
		dead += idx->seg_ndead[g];
	out[0] = idx->n_full_builds;
	out[1] = idx->n_delta_builds;
	out[2] = idx->n_consolidations;
	out[3] = idx->n_segs;
	out[4] = dead;
	out[5] = idx->n_live;
	out[6] = idx->n_pending;
}

static int
build_vocab(nxs_index_t *idx)
{
	uint64_t *totals = malloc(sizeof(uint64_t) * ((size_t)idx->n_terms + 1));
	int ret = -1;

	if (!totals)
		return -1;
	if (bkmirror_update(&idx->bk, idx->term_blob, idx->term_off,
	    idx->n_terms) == -1) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "BK-tree mirror failed");
		goto out;
	}
	for (uint32_t t = 1; t <= idx->n_terms; t++)
		totals[t - 1] = idx_term_total(idx, t);
	if (nxsb_engine_load_vocab(idx->engine, idx->n_terms,
	    idx->term_blob ? idx->term_blob : "", idx->term_off, totals,
	    idx->bk.parent, idx->bk.edge, idx->bk.rank) == -1) {
		nxs_set_error(idx->nxs, NXS_ERR_SYSTEM, "GPU vocabulary upload failed: %s",
		    nxsb_engine_errmsg(idx->engine));
		goto out;
	}
	idx->vocab_dirty = idx->totals_dirty = false;
	ret = 0;
out:
	free(totals);
	return ret;
}

/*
 * The vocabulary on the GPU is current but documents came or went: the
 * "total > 0" flags of the fuzzy matcher follow the counters in nxsterms
 * (ref idxterm.c:239 reads them at search time).
 */
static int
refresh_term_totals(nxs_index_t *idx)
{
	uint64_t *totals = malloc(sizeof(uint64_t) * ((size_t)idx->n_terms + 1));
	int ret = -1;

	if (!totals) {
		nxs_set_syserror(idx->nxs, NXS_ERR_SYSTEM, "out of memory");
		return -1;
	}
	for (uint32_t t = 1; t <= idx->n_terms; t++)
		totals[t - 1] = idx_term_total(idx, t);
	if (nxsb_engine_update_term_totals(idx->engine, idx->n_terms, totals) == -1) {
		nxs_set_error(idx->nxs, NXS_ERR_SYSTEM, "GPU vocabulary refresh failed: %s",
		    nxsb_engine_errmsg(idx->engine));
	} else {
		idx->totals_dirty = false;
		ret = 0;
	}
	free(totals);
	return ret;
}

int
idx_gpu_prepare(nxs_index_t *idx, bool need_vocab)
{
	if (!idx->engine) {
		idx->engine = idx->nxs->n_devices <= 1 ? nxsb_engine_create(idx->nxs->device)
		    : idx->nxs->shards
		    ? nxsb_engine_create_sharded(idx->nxs->devices, idx->nxs->n_devices)
		    : nxsb_engine_create_replicated(idx->nxs->devices, idx->nxs->n_devices);
		if (!idx->engine) {
			nxs_set_error(idx->nxs, NXS_ERR_SYSTEM,
			    "GPU engine unavailable: %s", nxsb_last_error());
			return -1;
		}
		idx->image_dirty = true;
		idx->vocab_dirty = true;
	}
	if ((idx->image_dirty || idx->n_pending || idx->stats_dirty) &&
	    refresh_image(idx) == -1)
		return -1;
	if (need_vocab && idx->vocab_dirty && idx->n_terms && build_vocab(idx) == -1)
		return -1;
	if (need_vocab && idx->totals_dirty && idx->n_terms && refresh_term_totals(idx) == -1)
		return -1;
	return 0;
}
