/*
 * ORACLE SHIM (test infrastructure, not product code).
 *
 * Stand-in for core/filters_builtin.c, which needs ICU and libstemmer.
 * Registers the three builtin filter names through the reference's own
 * nxs_filter_register() (core/filters.c:93), the same way the reference's
 * tests/t_filters.c:37-60 registers a test filter:
 *   normalizer -> ASCII lower-casing,  stopwords/stemmer -> pass-through.
 * Synthetic corpora are lowercase ASCII alphanumerics, for which the real
 * pipeline (NFKC casefold, diacritics, Snowball on non-words) is not
 * exercised by the parity harness.
 */
#include <stdlib.h>

#define __NXSLIB_PRIVATE
#include "nxs_impl.h"
#include "filters.h"
#include "strbuf.h"
#include "utils.h"

static filter_action_t
lower_filter(void *arg __unused, strbuf_t *buf)
{
	for (unsigned i = 0; i < buf->length; i++) {
		char c = buf->value[i];
		if (c >= 'A' && c <= 'Z') {
			buf->value[i] = c - 'A' + 'a';
		}
	}
	return FILT_MUTATION;
}

static filter_action_t
nop_filter(void *arg __unused, strbuf_t *buf __unused)
{
	return FILT_MUTATION;
}

static const filter_ops_t lower_ops = { .filter = lower_filter };
static const filter_ops_t nop_ops = { .filter = nop_filter };

int
filters_builtin_sysinit(nxs_t *nxs)
{
	nxs_filter_register(nxs, "normalizer", &lower_ops, NULL);
	nxs_filter_register(nxs, "stopwords", &nop_ops, NULL);
	nxs_filter_register(nxs, "stemmer", &nop_ops, NULL);
	return 0;
}
