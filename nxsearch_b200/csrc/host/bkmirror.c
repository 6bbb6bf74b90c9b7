/*
 * Host-side mirror of the reference's BK-tree.
 *
 * The reference answers a fuzzy lookup by a pruned breadth-first search of
 * a BK-tree keyed by Levenshtein distance (ref src/algo/bktree.c:160-275),
 * built by inserting terms in term-id order (ref src/index/terms.c:404-405,
 * src/index/idxterm.c:171).  The GPU scans the whole vocabulary instead, but
 * must return what that search would have returned -- including the matches
 * its half-open child range misses and its first-in-BFS-order pick (SURVEY
 * section 8a, F3).  So the tree SHAPE is replayed here, once per new term,
 * and exported as three flat arrays: parent[], edge[] (label of the edge to
 * the parent, capped at 63 as bktree.c:196) and rank[] (position in a full
 * breadth-first walk that enqueues children by ascending label, the order of
 * the ffs64 loop at bktree.c:265-269).  A pruned BFS visits a subsequence of
 * the full BFS, so the least rank among the reachable candidates is the
 * reference's answer.
 */
#include <stdlib.h>
#include <string.h>

#include "index.h"
#include "nxsb200_tools.h"

#define BK_EDGE_MAX	63

/* Byte-wise Levenshtein distance, two-row dynamic programme. */
int
bk_levdist(const char *a, size_t n, const char *b, size_t m)
{
	unsigned short stack_rows[2][72], *prev = stack_rows[0], *cur = stack_rows[1];
	unsigned short *heap = NULL;
	int d;

	if (m > n) {
		const char *ts = a; a = b; b = ts;
		const size_t tn = n; n = m; m = tn;
	}
	if (m == 0)
		return (int)n;
	if (m + 1 > 72) {
		if ((heap = malloc(sizeof(unsigned short) * 2 * (m + 1))) == NULL)
			return -1;
		prev = heap;
		cur = heap + m + 1;
	}
	for (size_t j = 0; j <= m; j++)
		prev[j] = j;
	for (size_t i = 1; i <= n; i++) {
		unsigned short *t;

		cur[0] = i;
		for (size_t j = 1; j <= m; j++) {
			unsigned v = prev[j - 1] + (a[i - 1] != b[j - 1]);

			if (prev[j] + 1u < v)
				v = prev[j] + 1u;
			if (cur[j - 1] + 1u < v)
				v = cur[j - 1] + 1u;
			cur[j] = v;
		}
		t = prev, prev = cur, cur = t;
	}
	d = prev[m];
	free(heap);
	return d;
}

void
bkmirror_free(bkmirror_t *bk)
{
	free(bk->parent);
	free(bk->edge);
	free(bk->rank);
	memset(bk, 0, sizeof(*bk));
}

/*
 * children[] grouped by parent and ordered by edge label -> BFS ranks.
 */
static int
bk_rank(bkmirror_t *bk)
{
	const uint32_t n = bk->n;
	uint32_t *start = calloc((size_t)n + 2, sizeof(uint32_t));
	uint32_t *kids = malloc(sizeof(uint32_t) * ((size_t)n + 1));
	uint32_t *queue = malloc(sizeof(uint32_t) * ((size_t)n + 1));
	uint32_t head = 0, tail = 0;

	if (!start || !kids || !queue) {
		free(start); free(kids); free(queue);
		return -1;
	}
	for (uint32_t t = 1; t < n; t++)
		start[bk->parent[t] + 2]++;
	for (uint32_t p = 0; p < n; p++)
		start[p + 2] += start[p + 1];
	for (uint32_t t = 1; t < n; t++)
		kids[start[bk->parent[t] + 1]++] = t;
	/* start[p] .. start[p + 1] now bounds p's children (insertion order). */
	for (uint32_t p = 0; p < n; p++) {
		/* Edge labels are unique per parent: insertion sort by label. */
		for (uint32_t i = start[p] + 1; i < start[p + 1]; i++) {
			const uint32_t v = kids[i];
			uint32_t j = i;

			while (j > start[p] && bk->edge[kids[j - 1]] > bk->edge[v]) {
				kids[j] = kids[j - 1];
				j--;
			}
			kids[j] = v;
		}
	}
	if (n)
		queue[tail++] = 0;
	while (head < tail) {
		const uint32_t p = queue[head];

		bk->rank[p] = head++;
		for (uint32_t i = start[p]; i < start[p + 1]; i++)
			queue[tail++] = kids[i];
	}
	free(start); free(kids); free(queue);
	return 0;
}

int
bkmirror_update(bkmirror_t *bk, const char *blob, const uint32_t *off,
    uint32_t n_terms)
{
	u64map_t *slots;

	if (n_terms == bk->n)
		return 0;
	if (n_terms > bk->cap) {
		const uint32_t ncap = n_terms + n_terms / 4 + 16;
		void *p;

		if ((p = realloc(bk->parent, sizeof(uint32_t) * ncap)) == NULL)
			return -1;
		bk->parent = p;
		if ((p = realloc(bk->edge, ncap)) == NULL)
			return -1;
		bk->edge = p;
		if ((p = realloc(bk->rank, sizeof(uint32_t) * ncap)) == NULL)
			return -1;
		bk->rank = p;
		bk->cap = ncap;
	}

	/* (parent, label) -> child, rebuilt from the arrays, then extended. */
	if ((slots = u64map_create(n_terms)) == NULL)
		return -1;
	for (uint32_t t = 1; t < bk->n; t++)
		u64map_put(slots, ((uint64_t)bk->parent[t] << 8) | bk->edge[t], t, NULL);

	for (uint32_t t = bk->n; t < n_terms; t++) {
		const char *ts = blob + off[t];
		const size_t tl = off[t + 1] - off[t];
		uint32_t cur = 0;

		bk->rank[t] = UINT32_MAX;
		if (t == 0) {
			bk->parent[0] = UINT32_MAX;
			bk->edge[0] = 0;
			continue;
		}
		for (;;) {
			int d = bk_levdist(ts, tl, blob + off[cur], off[cur + 1] - off[cur]);
			uint32_t next;

			if (d <= 0) {
				/*
				 * Duplicate value (or OOM): the reference's insert
				 * fails and the term never enters the tree
				 * (bktree.c:183-190).  Park it as an unreachable node.
				 */
				bk->parent[t] = t;
				bk->edge[t] = 255;
				break;
			}
			if (d > BK_EDGE_MAX)
				d = BK_EDGE_MAX;
			if (!u64map_get(slots, ((uint64_t)cur << 8) | (unsigned)d, &next)) {
				bk->parent[t] = cur;
				bk->edge[t] = d;
				u64map_put(slots, ((uint64_t)cur << 8) | (unsigned)d, t, NULL);
				break;
			}
			cur = next;
		}
	}
	u64map_destroy(slots);
	bk->n = n_terms;
	return bk_rank(bk);
}

int
nxsb_bkmirror_build(const char *blob, const uint32_t *off, uint32_t n_terms,
    uint32_t *parent, uint8_t *edge, uint32_t *rank)
{
	bkmirror_t bk = { 0 };

	if (bkmirror_update(&bk, blob, off, n_terms) == -1) {
		bkmirror_free(&bk);
		return -1;
	}
	memcpy(parent, bk.parent, sizeof(uint32_t) * n_terms);
	memcpy(edge, bk.edge, n_terms);
	memcpy(rank, bk.rank, sizeof(uint32_t) * n_terms);
	bkmirror_free(&bk);
	return 0;
}
