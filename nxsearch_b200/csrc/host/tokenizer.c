/*
 * Text front end; see tokenizer.h.
 */
#define _GNU_SOURCE
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tokenizer.h"

static const char *filter_names[] = {
	[FILT_NORMALIZER] = "normalizer",
	[FILT_STOPWORDS] = "stopwords",
	[FILT_STEMMER] = "stemmer",
};

static strmap_t *
load_stopwords(nxs_t *nxs, const char *lang)
{
	char *path = NULL, *line = NULL;
	size_t cap = 0;
	ssize_t len;
	strmap_t *map;
	FILE *fp;

	if (!lang || asprintf(&path, "%s/filters/stopwords/%s",
	    nxs->basedir, lang) == -1)
		return NULL;
	fp = fopen(path, "r");
	free(path);
	if (!fp)
		return NULL;	/* no stop words: not an error */
	if ((map = strmap_create(256)) != NULL) {
		while ((len = getline(&line, &cap, fp)) > 0) {
			if (len <= 1)
				continue;
			line[--len] = '\0';
			strmap_put(map, line, len, 1, NULL);
		}
	}
	free(line);
	fclose(fp);
	return map;
}

filter_pipeline_t *
filter_pipeline_create(nxs_t *nxs, nxs_params_t *params)
{
	filter_pipeline_t *fp = calloc(1, sizeof(*fp));
	const char **names;
	size_t n = 0;

	if (!fp)
		return NULL;
	names = nxs_params_get_strlist(params, "filters", &n);
	for (size_t i = 0; i < n; i++) {
		int kind = -1;

		for (unsigned k = 0; k < 3; k++) {
			if (strcmp(names[i], filter_names[k]) == 0)
				kind = k;
		}
		if (kind < 0 || fp->count == 8) {
			nxs_set_error(nxs, NXS_ERR_INVALID,
			    "filter `%s' not found", names[i]);
			free(names);
			free(fp);
			return NULL;
		}
		/*
		 * No Snowball in this build: a pipeline that names the stemmer
		 * would index and query UNSTEMMED terms under parameters that
		 * promise stemmed ones (and miss the terms of an index the
		 * reference built).  That is an error, not a silent pass-through;
		 * NXSB_STEMMER_PASSTHROUGH=1 accepts it knowingly (identity
		 * stemmer -- what the tests' compiled reference is built with).
		 */
		if (kind == FILT_STEMMER) {
			const char *ok = getenv("NXSB_STEMMER_PASSTHROUGH");
			static bool warned;

			if (!ok || strcmp(ok, "1") != 0) {
				nxs_set_error(nxs, NXS_ERR_INVALID,
				    "filter `stemmer' is not available in this build (no "
				    "Snowball); drop it from `filters' or set "
				    "NXSB_STEMMER_PASSTHROUGH=1 to index unstemmed terms");
				free(names);
				strmap_destroy(fp->stopwords);
				free(fp);
				return NULL;
			}
			if (!warned) {
				warned = true;
				fprintf(stderr, "nxsearch-b200: warning: the `stemmer' filter "
				    "passes terms through unchanged (NXSB_STEMMER_PASSTHROUGH=1)\n");
			}
		}
		fp->kinds[fp->count++] = kind;
		if (kind == FILT_STOPWORDS && !fp->stopwords)
			fp->stopwords = load_stopwords(nxs,
			    nxs_params_get_str(params, "lang"));
	}
	free(names);
	return fp;
}

void
filter_pipeline_destroy(filter_pipeline_t *fp)
{
	if (fp) {
		strmap_destroy(fp->stopwords);
		free(fp);
	}
}

tokenset_t *
tokenset_create(void)
{
	tokenset_t *ts = calloc(1, sizeof(*ts));

	if (ts && (ts->map = strmap_create(16)) == NULL) {
		free(ts);
		ts = NULL;
	}
	return ts;
}

void
tokenset_destroy(tokenset_t *ts)
{
	if (!ts)
		return;
	for (uint32_t i = 0; i < ts->count; i++)
		free(ts->list[i].str);
	free(ts->list);
	strmap_destroy(ts->map);
	free(ts);
}

/*
 * The filter pipeline over one word, in place.  Returns 1 to keep the token,
 * 0 if a filter discarded it (or left it empty: ref filters.c:206-208).
 */
int
filter_apply(const filter_pipeline_t *fp, char *buf, size_t *lenp)
{
	const size_t len = *lenp;

	for (unsigned i = 0; i < fp->count; i++) {
		switch (fp->kinds[i]) {
		case FILT_NORMALIZER:
			for (size_t k = 0; k < len; k++) {
				if (buf[k] >= 'A' && buf[k] <= 'Z')
					buf[k] += 'a' - 'A';
			}
			break;
		case FILT_STOPWORDS:
			if (fp->stopwords && strmap_get(fp->stopwords, buf, len, NULL))
				return 0;
			break;
		case FILT_STEMMER:
			break;	/* identity: only with NXSB_STEMMER_PASSTHROUGH=1 */
		}
		if (len == 0)
			return 0;
	}
	return len != 0;
}

int
tokenize_value(filter_pipeline_t *fp, tokenset_t *ts, const char *val,
    size_t len, int32_t *slot)
{
	char stackbuf[128], *buf = stackbuf;
	uint32_t idx;
	int ret = -1;

	*slot = -1;
	if (len >= sizeof(stackbuf) && (buf = malloc(len + 1)) == NULL)
		return -1;
	memcpy(buf, val, len);
	buf[len] = '\0';

	if (!filter_apply(fp, buf, &len)) {
		ret = 0;	/* discarded */
		goto out;
	}

	if (strmap_get(ts->map, buf, len, &idx)) {
		ts->list[idx].count++;
		ts->seen++;
		*slot = idx;
		ret = 0;
		goto out;
	}
	if (ts->count == ts->cap) {
		const uint32_t ncap = ts->cap ? ts->cap * 2 : 16;
		token_t *nl = realloc(ts->list, sizeof(token_t) * ncap);

		if (!nl)
			goto out;
		ts->list = nl;
		ts->cap = ncap;
	}
	if ((ts->list[ts->count].str = strndup(buf, len)) == NULL)
		goto out;
	ts->list[ts->count].len = len;
	ts->list[ts->count].count = 1;
	ts->list[ts->count].term_id = 0;
	if (strmap_put(ts->map, buf, len, ts->count, NULL) == -1) {
		free(ts->list[ts->count].str);
		goto out;
	}
	*slot = ts->count++;
	ts->seen++;
	ts->data_len += len;
	ret = 0;
out:
	if (buf != stackbuf)
		free(buf);
	return ret;
}

static inline bool
is_word_byte(unsigned char c)
{
	return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') ||
	    (c >= 'A' && c <= 'Z') || c >= 0x80;
}

tokenset_t *
tokenize(filter_pipeline_t *fp, const char *text, size_t len)
{
	tokenset_t *ts = tokenset_create();
	size_t i = 0;

	if (!ts)
		return NULL;
	while (i < len && text[i]) {
		size_t s;
		int32_t slot;

		while (i < len && text[i] && !is_word_byte(text[i]))
			i++;
		s = i;
		while (i < len && is_word_byte(text[i]))
			i++;
		if (i > s && tokenize_value(fp, ts, text + s, i - s, &slot) == -1) {
			tokenset_destroy(ts);
			return NULL;
		}
	}
	return ts;
}
