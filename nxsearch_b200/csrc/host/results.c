/*
 * Response object: the top-N (doc id, score) list of one query, in the
 * order the engine produced it (descending score), with the iterator and
 * JSON rendering of ref src/core/results.c:88-247.
 *
 * The reference materialises a yyjson document per result eagerly
 * (results.c:153-160); at GPU query rates that would dominate, so the JSON
 * text is rendered only if nxs_resp_tojson() is called.  The
 * text is what the reference prints: {"results":[{"doc_id":N,"score":X},
 * ...],"count":K} with X the float score widened to double and printed in
 * yyjson's shortest round-trip form (golden: ref tests/t_misc.c:115-117).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nxs_impl.h"
#include "nxsb200_tools.h"

struct nxs_resp {
	uint32_t	count;
	uint32_t	iter;
	uint64_t *	ids;		/* both arrays live behind the header */
	float *		scores;
};

nxs_resp_t *
nxs_resp_from_arrays(const uint64_t *ids, const float *scores, uint32_t n)
{
	/* One allocation per response: header | ids[n] | scores[n]. */
	nxs_resp_t *r = malloc(sizeof(*r) + (size_t)n * (sizeof(uint64_t) + sizeof(float)));

	if (!r)
		return NULL;
	r->ids = (uint64_t *)(r + 1);
	r->scores = (float *)(r->ids + n);
	if (n) {
		memcpy(r->ids, ids, sizeof(uint64_t) * n);
		memcpy(r->scores, scores, sizeof(float) * n);
	}
	r->count = n;
	r->iter = 0;
	return r;
}

NXS_API void
nxs_resp_release(nxs_resp_t *r)
{
	free(r);
}

NXS_API void
nxs_resp_iter_reset(nxs_resp_t *r)
{
	r->iter = 0;
}

NXS_API bool
nxs_resp_iter_result(nxs_resp_t *r, nxs_doc_id_t *doc_id, float *score)
{
	if (r->iter >= r->count)
		return false;
	*doc_id = r->ids[r->iter];
	*score = r->scores[r->iter];
	r->iter++;
	return true;
}

NXS_API unsigned
nxs_resp_resultcount(const nxs_resp_t *r)
{
	return r->count;
}

NXS_API char *
nxs_resp_tojson(nxs_resp_t *r, size_t *len)
{
	/* {"doc_id":18446744073709551615,"score":-1.7976931348623157e308}, */
	const size_t cap = 64 + (size_t)r->count * 80;
	char *buf = malloc(cap);
	size_t n = 0;

	if (!buf)
		return NULL;
	n += sprintf(buf + n, "{\"results\":[");
	for (uint32_t i = 0; i < r->count; i++) {
		n += sprintf(buf + n, "%s{\"doc_id\":%llu,\"score\":", i ? "," : "",
		    (unsigned long long)r->ids[i]);
		n += json_format_real((double)r->scores[i], buf + n);
		buf[n++] = '}';
	}
	n += sprintf(buf + n, "],\"count\":%u}", r->count);
	if (len)
		*len = n;
	return buf;
}

/* include/nxsb200_tools.h: drain a batch of responses through the iterator. */
NXS_API uint64_t
nxsb_resp_collect(void *const *resps, size_t n, uint32_t stride,
    uint32_t *counts, uint64_t *ids, float *scores)
{
	uint64_t total = 0;

	for (size_t i = 0; i < n; i++) {
		nxs_resp_t *r = resps[i];
		uint32_t c = 0;

		if (r) {
			nxs_resp_iter_reset(r);
			while (c < stride && nxs_resp_iter_result(r,
			    &ids[i * (size_t)stride + c], &scores[i * (size_t)stride + c]))
				c++;
		}
		counts[i] = c;
		total += c;
	}
	return total;
}
