/*
 * Oracle shim (TEST INFRASTRUCTURE, never linked into the product): lets a
 * test read the candidate list the reference's OWN bktree_search() returns
 * for a query string -- idxterm_fuzzysearch() (ref src/index/idxterm.c:
 * 210-249) keeps that list private and only returns its pick.  Nothing here
 * restates an algorithm; it calls the reference's functions.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define __NXSLIB_PRIVATE
#include "nxs_impl.h"
#include "index.h"
#include "bktree.h"
#include "deque.h"
#include "levdist.h"

size_t
nxsb_ref_fuzzy_list(nxs_index_t *idx, const char *value, size_t len,
    uint32_t *ids, size_t cap)
{
	const unsigned total_len = offsetof(idxterm_t, value[(unsigned)len + 1]);
	idxterm_t *probe = calloc(1, total_len), *it;
	deque_t *results = deque_create(0, 0);
	size_t n = 0;

	if (!probe || !results)
		goto out;
	memcpy(probe->value, value, len);
	probe->value_len = len;
	if (bktree_search(idx->term_bkt, LEVDIST_TOLERANCE, probe, results) == -1)
		goto out;
	while ((it = deque_pop_front(results)) != NULL) {
		if (n < cap)
			ids[n] = it->id;
		n++;
	}
out:
	if (results)
		deque_destroy(results);
	free(probe);
	return n;
}
