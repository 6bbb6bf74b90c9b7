/*
 * Synthetic corpus generator and bulk index-file reader/writer.
 * See include/nxsb200_tools.h for the interface and SURVEY.md section 8(d)
 * for the workload definition.  On-disk layout follows the reference's
 * src/index/storage.h:12-133 (all integers big-endian):
 *
 *   nxsterms: 16-byte header { "NXS_T", ver=1, 2 reserved, data_len u32,
 *             reserved u32 } then per term
 *             { len u16, bytes, NUL, pad to 8, total u64 }
 *   nxsdtmap: 32-byte header { "NXS_D", ver=1, 2 reserved, data_len u64,
 *             token_count u64, doc_count u32, reserved u32 } then per doc
 *             { doc_id u64, doc_len u32, n u32, n x (term_id u32, count u32) }
 *
 * Everything is integer arithmetic except the Zipf alias table, which uses
 * only IEEE double + - * / (bit-reproducible on any x86-64 host).
 */
#include <sys/types.h>
#include <sys/stat.h>
#include <sys/mman.h>

#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <endian.h>

#include "nxsb200_tools.h"
#include "hashmap.h"

#define	FILE_STEP	(32UL * 1024)	// ref IDX_SIZE_STEP, index.h:24
#define	DOCLEN_MIN	16
#define	DOCLEN_SPAN	96		// doc length uniform in [16, 112)

static const char alphabet[36] = "abcdefghijklmnopqrstuvwxyz0123456789";

/*
 * PRNG: xoshiro256** seeded through splitmix64 (both public domain
 * algorithms by Blackman & Vigna).
 */

static inline uint64_t
splitmix64(uint64_t *x)
{
	uint64_t z = (*x += UINT64_C(0x9e3779b97f4a7c15));

	z = (z ^ (z >> 30)) * UINT64_C(0xbf58476d1ce4e5b9);
	z = (z ^ (z >> 27)) * UINT64_C(0x94d049bb133111eb);
	return z ^ (z >> 31);
}

typedef struct { uint64_t s[4]; } rng_t;

static inline void
rng_seed(rng_t *r, uint64_t seed, uint64_t stream)
{
	uint64_t x = seed ^ (stream * UINT64_C(0xd1342543de82ef95));

	for (int i = 0; i < 4; i++)
		r->s[i] = splitmix64(&x);
}

static inline uint64_t
rotl64(uint64_t x, int k)
{
	return (x << k) | (x >> (64 - k));
}

static inline uint64_t
rng_next(rng_t *r)
{
	uint64_t *s = r->s;
	const uint64_t result = rotl64(s[1] * 5, 7) * 9;
	const uint64_t t = s[1] << 17;

	s[2] ^= s[0];
	s[3] ^= s[1];
	s[1] ^= s[2];
	s[0] ^= s[3];
	s[2] ^= t;
	s[3] = rotl64(s[3], 45);
	return result;
}

/* Uniform integer in [0, n) by multiply-shift on the high 32 bits. */
static inline uint32_t
rng_below(rng_t *r, uint32_t n)
{
	return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32);
}

/*
 * Zipf(s = 1) over ranks 1..V through a Walker/Vose alias table, so a draw
 * is O(1): one 64-bit random, two table reads, no floating point.
 */

typedef struct {
	uint32_t	n;
	uint32_t *	thresh;		// P(keep slot) scaled to 2^32
	uint32_t *	alias;
} zipf_t;

static void
zipf_free(zipf_t *z)
{
	free(z->thresh);
	free(z->alias);
}

static int
zipf_init(zipf_t *z, uint32_t n)
{
	double *q = malloc(sizeof(double) * n), norm = 0;
	uint32_t *small = malloc(sizeof(uint32_t) * n);
	uint32_t *large = malloc(sizeof(uint32_t) * n);
	uint32_t ns = 0, nl = 0;

	z->n = n;
	z->thresh = malloc(sizeof(uint32_t) * n);
	z->alias = malloc(sizeof(uint32_t) * n);
	if (!q || !small || !large || !z->thresh || !z->alias) {
		free(q); free(small); free(large);
		zipf_free(z);
		return -1;
	}
	for (uint32_t i = 0; i < n; i++)
		norm += 1.0 / (double)(i + 1);
	for (uint32_t i = 0; i < n; i++) {
		q[i] = (1.0 / (double)(i + 1)) / norm * (double)n;
		z->alias[i] = i;
	}
	/* Descending index order keeps the work lists deterministic. */
	for (uint32_t i = n; i-- > 0;) {
		if (q[i] < 1.0)
			small[ns++] = i;
		else
			large[nl++] = i;
	}
	while (ns && nl) {
		const uint32_t s = small[--ns], l = large[nl - 1];

		z->alias[s] = l;
		q[l] = (q[l] + q[s]) - 1.0;
		if (q[l] < 1.0) {
			nl--;
			small[ns++] = l;
		}
	}
	for (uint32_t i = 0; i < n; i++) {
		const double t = q[i] * 4294967296.0;

		z->thresh[i] = (z->alias[i] == i || t >= 4294967295.0) ?
		    UINT32_MAX : (uint32_t)t;
	}
	free(q); free(small); free(large);
	return 0;
}

static inline uint32_t
zipf_draw(const zipf_t *z, rng_t *r)
{
	const uint64_t x = rng_next(r);
	const uint32_t slot = (uint32_t)(((x >> 32) * (uint64_t)z->n) >> 32);

	return ((uint32_t)x <= z->thresh[slot]) ? slot : z->alias[slot];
}

/*
 * Vocabulary: term i is the first string of its own stream that no earlier
 * term already uses.
 */
static int
vocab_generate(nxsb_corpus_t *c, uint64_t seed)
{
	const uint32_t n = c->n_terms;
	strmap_t *seen = strmap_create(n);
	size_t off = 0;

	c->term_off = malloc(sizeof(uint32_t) * ((size_t)n + 1));
	c->term_blob = malloc((size_t)n * 13 + 1);
	if (!seen || !c->term_off || !c->term_blob) {
		strmap_destroy(seen);
		return -1;
	}
	for (uint32_t i = 0; i < n; i++) {
		rng_t r;
		char buf[16];
		unsigned len;

		rng_seed(&r, seed ^ UINT64_C(0x766f636162756c61), i);
		do {
			len = 4 + rng_below(&r, 9);
			for (unsigned k = 0; k < len; k++)
				buf[k] = alphabet[rng_below(&r, 36)];
		} while (strmap_put(seen, buf, len, i, NULL) != 1);

		c->term_off[i] = off;
		memcpy(c->term_blob + off, buf, len);
		off += len;
	}
	c->term_off[n] = off;
	c->term_blob[off] = '\0';
	strmap_destroy(seen);
	return 0;
}

/*
 * Documents.  Each worker owns a contiguous slice and appends to its own
 * buffer; the slices are concatenated afterwards.
 */

typedef struct {
	const zipf_t *	zipf;
	uint64_t	seed;
	uint64_t	first_doc;	// global index of the corpus' doc 0
	uint32_t	lo, hi;		// local doc range [lo, hi)
	uint32_t *	doc_len;
	uint32_t *	doc_n;		// unique terms per doc
	uint32_t *	pairs;		// own buffer
	uint64_t	n_pairs, cap;
	int		error;
} genjob_t;

static void
sort_u32(uint32_t *a, unsigned n)
{
	/* Insertion sort: n < 112 and mostly small runs. */
	for (unsigned i = 1; i < n; i++) {
		const uint32_t v = a[i];
		unsigned j = i;

		while (j && a[j - 1] > v) {
			a[j] = a[j - 1];
			j--;
		}
		a[j] = v;
	}
}

static void *
gen_worker(void *arg)
{
	genjob_t *job = arg;
	uint32_t toks[DOCLEN_MIN + DOCLEN_SPAN];

	for (uint32_t d = job->lo; d < job->hi; d++) {
		rng_t r;
		unsigned len, n = 0;

		rng_seed(&r, job->seed, job->first_doc + d + 1);
		len = DOCLEN_MIN + rng_below(&r, DOCLEN_SPAN);
		for (unsigned k = 0; k < len; k++)
			toks[k] = zipf_draw(job->zipf, &r) + 1;
		sort_u32(toks, len);

		if (job->n_pairs + len > job->cap) {
			const uint64_t ncap = job->cap * 2 + 4096;
			uint32_t *np = realloc(job->pairs, ncap * 8);

			if (!np) {
				job->error = 1;
				return NULL;
			}
			job->pairs = np;
			job->cap = ncap;
		}
		for (unsigned k = 0; k < len;) {
			unsigned e = k + 1;

			while (e < len && toks[e] == toks[k])
				e++;
			job->pairs[2 * (job->n_pairs + n)] = toks[k];
			job->pairs[2 * (job->n_pairs + n) + 1] = e - k;
			n++;
			k = e;
		}
		job->n_pairs += n;
		job->doc_len[d] = len;
		job->doc_n[d] = n;
	}
	return NULL;
}

static int
corpus_stats(nxsb_corpus_t *c)
{
	c->term_total = calloc(c->n_terms ? c->n_terms : 1, sizeof(uint64_t));
	c->term_df = calloc(c->n_terms ? c->n_terms : 1, sizeof(uint32_t));
	if (!c->term_total || !c->term_df)
		return -1;
	for (uint64_t j = 0; j < c->n_pairs; j++) {
		const uint32_t t = c->pairs[2 * j] - 1;

		if (t < c->n_terms) {
			c->term_total[t] += c->pairs[2 * j + 1];
			c->term_df[t]++;
		}
	}
	return 0;
}

void
nxsb_corpus_free(nxsb_corpus_t *c)
{
	if (!c)
		return;
	free(c->doc_ids);
	free(c->doc_len);
	free(c->doc_off);
	free(c->pairs);
	free(c->term_blob);
	free(c->term_off);
	free(c->term_total);
	free(c->term_df);
	free(c);
}

nxsb_corpus_t *
nxsb_corpus_generate(uint64_t seed, uint32_t n_terms, uint64_t first_doc,
    uint32_t n_docs, int sparse_ids, int nthreads)
{
	nxsb_corpus_t *c = calloc(1, sizeof(*c));
	genjob_t *jobs = NULL;
	pthread_t *tids = NULL;
	uint32_t *doc_n = NULL;
	zipf_t zipf = { 0 };
	uint64_t off = 0;
	int ok = 0;

	if (!c || n_terms == 0)
		goto out;
	if (nthreads <= 0)
		nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
	if (nthreads < 1)
		nthreads = 1;
	if ((uint32_t)nthreads > n_docs / 1024 + 1)
		nthreads = n_docs / 1024 + 1;

	c->n_docs = n_docs;
	c->n_terms = n_terms;
	if (vocab_generate(c, seed) == -1 || zipf_init(&zipf, n_terms) == -1)
		goto out;

	c->doc_ids = malloc(sizeof(uint64_t) * ((size_t)n_docs + 1));
	c->doc_len = malloc(sizeof(uint32_t) * ((size_t)n_docs + 1));
	c->doc_off = malloc(sizeof(uint64_t) * ((size_t)n_docs + 1));
	doc_n = malloc(sizeof(uint32_t) * ((size_t)n_docs + 1));
	jobs = calloc(nthreads, sizeof(genjob_t));
	tids = calloc(nthreads, sizeof(pthread_t));
	if (!c->doc_ids || !c->doc_len || !c->doc_off || !doc_n || !jobs || !tids)
		goto out;

	for (int t = 0; t < nthreads; t++) {
		genjob_t *job = &jobs[t];

		job->zipf = &zipf;
		job->seed = seed;
		job->first_doc = first_doc;
		job->lo = (uint64_t)n_docs * t / nthreads;
		job->hi = (uint64_t)n_docs * (t + 1) / nthreads;
		job->doc_len = c->doc_len;
		job->doc_n = doc_n;
		if (pthread_create(&tids[t], NULL, gen_worker, job) != 0) {
			gen_worker(job);
			tids[t] = 0;
		}
	}
	for (int t = 0; t < nthreads; t++) {
		if (tids[t])
			pthread_join(tids[t], NULL);
		if (jobs[t].error)
			goto out;
		c->n_pairs += jobs[t].n_pairs;
	}
	c->pairs = malloc(c->n_pairs * 8 + 8);
	if (!c->pairs)
		goto out;
	for (int t = 0; t < nthreads; t++) {
		memcpy(c->pairs + 2 * off, jobs[t].pairs, jobs[t].n_pairs * 8);
		off += jobs[t].n_pairs;
		free(jobs[t].pairs);
		jobs[t].pairs = NULL;
	}
	off = 0;
	for (uint32_t d = 0; d < n_docs; d++) {
		const uint64_t g = first_doc + d;

		c->doc_off[d] = off;
		off += doc_n[d];
		c->token_count += c->doc_len[d];
		if (sparse_ids) {
			uint64_t x = seed ^ g;
			/* Strictly increasing: 2^20 stride plus a 20-bit jitter. */
			c->doc_ids[d] = ((g + 1) << 20) | (splitmix64(&x) & 0xfffff);
		} else {
			c->doc_ids[d] = g + 1;
		}
	}
	c->doc_off[n_docs] = off;
	c->doc_count = n_docs;
	if (corpus_stats(c) == -1)
		goto out;
	ok = 1;
out:
	if (jobs) {
		for (int t = 0; t < nthreads; t++)
			free(jobs[t].pairs);
	}
	free(jobs);
	free(tids);
	free(doc_n);
	zipf_free(&zipf);
	if (!ok) {
		nxsb_corpus_free(c);
		c = NULL;
	}
	return c;
}

void
nxsb_corpus_query_terms(uint64_t seed, uint32_t n_terms, const uint32_t *df,
    uint32_t *term_ids, size_t n)
{
	zipf_t zipf;
	rng_t r;

	if (zipf_init(&zipf, n_terms) == -1) {
		memset(term_ids, 0, sizeof(uint32_t) * n);
		return;
	}
	rng_seed(&r, seed, UINT64_C(0x7175657279));
	for (size_t i = 0; i < n; i++) {
		uint32_t t;

		do {
			t = zipf_draw(&zipf, &r);
		} while (df && df[t] == 0);
		term_ids[i] = t + 1;
	}
	zipf_free(&zipf);
}

void
nxsb_corpus_fuzzy_terms(uint64_t seed, const nxsb_corpus_t *c, char *out,
    size_t stride, size_t n)
{
	strmap_t *vocab = strmap_create(c->n_terms);
	rng_t r;

	for (uint32_t i = 0; vocab && i < c->n_terms; i++) {
		strmap_put(vocab, c->term_blob + c->term_off[i],
		    c->term_off[i + 1] - c->term_off[i], i, NULL);
	}
	rng_seed(&r, seed, UINT64_C(0x66757a7a79));
	for (size_t i = 0; i < n; i++) {
		char buf[32];
		unsigned len;

		do {
			const uint32_t t = rng_below(&r, c->n_terms);
			const unsigned nedits = 1 + rng_below(&r, 2);

			len = c->term_off[t + 1] - c->term_off[t];
			if (len > 20)
				len = 20;
			memcpy(buf, c->term_blob + c->term_off[t], len);
			for (unsigned e = 0; e < nedits; e++) {
				const unsigned op = rng_below(&r, 3);
				const char ch = alphabet[rng_below(&r, 36)];
				unsigned pos;

				if (op == 0 && len > 0) {		// substitute
					pos = rng_below(&r, len);
					buf[pos] = ch;
				} else if (op == 1 && len < 24) {	// insert
					pos = rng_below(&r, len + 1);
					memmove(buf + pos + 1, buf + pos, len - pos);
					buf[pos] = ch;
					len++;
				} else if (len > 1) {			// delete
					pos = rng_below(&r, len);
					memmove(buf + pos, buf + pos + 1, len - pos - 1);
					len--;
				}
			}
		} while (len == 0 || len >= stride ||
		    (vocab && strmap_get(vocab, buf, len, NULL)));
		memcpy(out + i * stride, buf, len);
		out[i * stride + len] = '\0';
	}
	strmap_destroy(vocab);
}

/*
 * File writers.  The file is sized up front, mapped, filled, and the header
 * (with data_len) written last, mirroring the reference's publish order
 * (src/index/terms.c:304-305, src/index/dtmap.c:333-337).
 */

static void *
map_new_file(const char *path, size_t len, int *fdp)
{
	const size_t flen = (len + FILE_STEP - 1) / FILE_STEP * FILE_STEP;
	int fd = open(path, O_RDWR | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
	void *p;

	if (fd == -1)
		return NULL;
	if (ftruncate(fd, flen) == -1 ||
	    (p = mmap(NULL, flen, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0))
	    == MAP_FAILED) {
		close(fd);
		return NULL;
	}
	*fdp = fd;
	return p;
}

static inline void
put16(uint8_t *p, uint16_t v) { v = htobe16(v); memcpy(p, &v, 2); }
static inline void
put32(uint8_t *p, uint32_t v) { v = htobe32(v); memcpy(p, &v, 4); }
static inline void
put64(uint8_t *p, uint64_t v) { v = htobe64(v); memcpy(p, &v, 8); }
static inline uint16_t
get16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return be16toh(v); }
static inline uint32_t
get32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return be32toh(v); }
static inline uint64_t
get64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return be64toh(v); }

static inline size_t
term_block_len(size_t len)
{
	/* len u16 + bytes + NUL, padded to 8, then the u64 total. */
	return ((2 + len + 1 + 7) & ~(size_t)7) + 8;
}

int
nxsb_write_terms_file(const char *path, const nxsb_corpus_t *c)
{
	size_t data_len = 0, off = 16;
	uint8_t *base;
	int fd;

	for (uint32_t i = 0; i < c->n_terms; i++)
		data_len += term_block_len(c->term_off[i + 1] - c->term_off[i]);
	if (data_len > UINT32_MAX) {
		errno = EFBIG;
		return -1;
	}
	if ((base = map_new_file(path, 16 + data_len, &fd)) == NULL)
		return -1;

	for (uint32_t i = 0; i < c->n_terms; i++) {
		const size_t len = c->term_off[i + 1] - c->term_off[i];
		const size_t blk = term_block_len(len);

		memset(base + off, 0, blk);
		put16(base + off, len);
		memcpy(base + off + 2, c->term_blob + c->term_off[i], len);
		put64(base + off + blk - 8, c->term_total ? c->term_total[i] : 0);
		off += blk;
	}
	memset(base, 0, 16);
	memcpy(base, "NXS_T", 5);
	base[5] = 1;
	put32(base + 8, data_len);

	munmap(base, (16 + data_len + FILE_STEP - 1) / FILE_STEP * FILE_STEP);
	close(fd);
	return 0;
}

int
nxsb_write_dtmap_file(const char *path, const nxsb_corpus_t *c)
{
	const size_t data_len = (size_t)c->n_docs * 16 + c->n_pairs * 8;
	size_t off = 32;
	uint8_t *base;
	int fd;

	if ((base = map_new_file(path, 32 + data_len, &fd)) == NULL)
		return -1;

	for (uint32_t d = 0; d < c->n_docs; d++) {
		const uint64_t s = c->doc_off[d], e = c->doc_off[d + 1];

		put64(base + off, c->doc_ids[d]);
		put32(base + off + 8, c->doc_len[d]);
		put32(base + off + 12, e - s);
		off += 16;
		for (uint64_t j = s; j < e; j++) {
			put32(base + off, c->pairs[2 * j]);
			put32(base + off + 4, c->pairs[2 * j + 1]);
			off += 8;
		}
	}
	memset(base, 0, 32);
	memcpy(base, "NXS_D", 5);
	base[5] = 1;
	put64(base + 8, data_len);
	put64(base + 16, c->token_count);
	put32(base + 24, c->doc_count);

	munmap(base, (32 + data_len + FILE_STEP - 1) / FILE_STEP * FILE_STEP);
	close(fd);
	return 0;
}

/*
 * Reader.
 */

static const uint8_t *
map_file_ro(const char *path, size_t *lenp)
{
	struct stat st;
	void *p;
	int fd;

	if ((fd = open(path, O_RDONLY | O_CLOEXEC)) == -1)
		return NULL;
	if (fstat(fd, &st) == -1 || st.st_size == 0) {
		close(fd);
		errno = EINVAL;
		return NULL;
	}
	p = mmap(NULL, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if (p == MAP_FAILED)
		return NULL;
	*lenp = st.st_size;
	return p;
}

nxsb_corpus_t *
nxsb_read_index_files(const char *terms_path, const char *dtmap_path)
{
	nxsb_corpus_t *c = calloc(1, sizeof(*c));
	const uint8_t *tb = NULL, *db = NULL;
	size_t tlen = 0, dlen = 0, off, end, blob = 0;
	uint64_t npairs = 0;
	uint32_t nterms = 0, ndocs = 0, live = 0;
	u64map_t *byid = NULL;
	uint8_t *dead = NULL;
	int ok = 0;

	if (!c)
		return NULL;
	if ((tb = map_file_ro(terms_path, &tlen)) == NULL ||
	    (db = map_file_ro(dtmap_path, &dlen)) == NULL)
		goto out;
	if (tlen < 16 || memcmp(tb, "NXS_T", 5) != 0 || tb[5] != 1 ||
	    dlen < 32 || memcmp(db, "NXS_D", 5) != 0 || db[5] != 1) {
		errno = EINVAL;
		goto out;
	}

	/* Terms: two passes (count, then fill). */
	end = 16 + (size_t)get32(tb + 8);
	if (end > tlen) {
		errno = EINVAL;
		goto out;
	}
	for (off = 16; off < end;) {
		const size_t len = (off + 2 <= end) ? get16(tb + off) : 0;

		if (len == 0 || off + term_block_len(len) > end) {
			errno = EINVAL;
			goto out;
		}
		blob += len;
		nterms++;
		off += term_block_len(len);
	}
	c->n_terms = nterms;
	c->term_off = malloc(sizeof(uint32_t) * ((size_t)nterms + 1));
	c->term_blob = malloc(blob + 1);
	c->term_total = calloc(nterms ? nterms : 1, sizeof(uint64_t));
	if (!c->term_off || !c->term_blob || !c->term_total)
		goto out;
	blob = 0, nterms = 0;
	for (off = 16; off < end;) {
		const size_t len = get16(tb + off);
		const size_t blk = term_block_len(len);

		c->term_off[nterms] = blob;
		memcpy(c->term_blob + blob, tb + off + 2, len);
		c->term_total[nterms] = get64(tb + off + blk - 8);
		blob += len;
		nterms++;
		off += blk;
	}
	c->term_off[nterms] = blob;
	c->term_blob[blob] = '\0';

	/*
	 * Documents.  Pass 1 counts blocks and applies deletions: a block
	 * with doc_id 0 is skipped; a {doc_id, doc_len 0} marker removes the
	 * earlier live block of that id (ref dtmap.c:357-384).
	 */
	end = 32 + (size_t)get64(db + 8);
	if (end > dlen) {
		errno = EINVAL;
		goto out;
	}
	for (off = 32; off < end;) {
		uint32_t n;

		if (off + 16 > end) {
			errno = EINVAL;
			goto out;
		}
		n = get32(db + off + 12);
		if (off + 16 + (size_t)n * 8 > end) {
			errno = EINVAL;
			goto out;
		}
		ndocs++;
		off += 16 + (size_t)n * 8;
	}
	dead = calloc(ndocs ? ndocs : 1, 1);
	byid = u64map_create(ndocs);
	if (!dead || !byid)
		goto out;
	ndocs = 0;
	for (off = 32; off < end;) {
		const uint64_t id = get64(db + off);
		const uint32_t dl = get32(db + off + 8);
		const uint32_t n = get32(db + off + 12);

		if (id == 0) {
			dead[ndocs] = 1;
		} else if (dl == 0) {
			uint32_t prev;

			dead[ndocs] = 1;
			if (u64map_get(byid, id, &prev)) {
				dead[prev] = 1;
				u64map_del(byid, id);
			}
		} else {
			u64map_put(byid, id, ndocs, NULL);
		}
		ndocs++;
		off += 16 + (size_t)n * 8;
	}
	ndocs = 0;
	for (off = 32; off < end;) {
		const uint32_t n = get32(db + off + 12);

		if (!dead[ndocs]) {
			live++;
			npairs += n;
		}
		ndocs++;
		off += 16 + (size_t)n * 8;
	}

	c->n_docs = live;
	c->n_pairs = npairs;
	c->doc_ids = malloc(sizeof(uint64_t) * ((size_t)live + 1));
	c->doc_len = malloc(sizeof(uint32_t) * ((size_t)live + 1));
	c->doc_off = malloc(sizeof(uint64_t) * ((size_t)live + 1));
	c->pairs = malloc(npairs * 8 + 8);
	if (!c->doc_ids || !c->doc_len || !c->doc_off || !c->pairs)
		goto out;
	ndocs = 0, live = 0, npairs = 0;
	for (off = 32; off < end;) {
		const uint32_t n = get32(db + off + 12);

		if (!dead[ndocs]) {
			c->doc_ids[live] = get64(db + off);
			c->doc_len[live] = get32(db + off + 8);
			c->doc_off[live] = npairs;
			for (uint32_t j = 0; j < n; j++) {
				const uint8_t *p = db + off + 16 + (size_t)j * 8;

				c->pairs[2 * npairs] = get32(p);
				c->pairs[2 * npairs + 1] = get32(p + 4);
				npairs++;
			}
			live++;
		}
		ndocs++;
		off += 16 + (size_t)n * 8;
	}
	c->doc_off[live] = npairs;
	/* Header counters are authoritative for scoring (ref ranking.c:77,163). */
	c->token_count = get64(db + 16);
	c->doc_count = get32(db + 24);

	{
		uint64_t *totals = c->term_total;

		c->term_total = NULL;
		if (corpus_stats(c) == -1) {
			c->term_total = totals;
			goto out;
		}
		free(c->term_total);
		c->term_total = totals;
	}
	ok = 1;
out:
	if (tb)
		munmap((void *)(uintptr_t)tb, tlen);
	if (db)
		munmap((void *)(uintptr_t)db, dlen);
	free(dead);
	u64map_destroy(byid);
	if (!ok) {
		nxsb_corpus_free(c);
		c = NULL;
	}
	return c;
}
