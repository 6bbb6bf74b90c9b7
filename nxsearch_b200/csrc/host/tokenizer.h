/*
 * Text front end: word segmentation, the filter pipeline and token sets.
 *
 * OUT OF THE ACCELERATED PATH.  The reference
 * uses ICU UAX#29 word breaking, ICU NFKC case folding / transliteration and
 * the Snowball stemmer (ref src/core/tokenizer.c:234-302,
 * src/core/filters_builtin.c); none of those libraries exist here, so this
 * front end implements the part that needs no Unicode tables:
 *   - words by the UAX #29 rules ("i.b.m", "doesn't", "snake_case", "3.14"
 *     are one word each; ref src/tests/t_tokenize.c:17-62 are the golden
 *     cases); outside ASCII a code point is a letter unless its block holds
 *     none (punctuation, symbols, emoji), no dictionary segmentation;
 *   - "normalizer" lower-cases and strips diacritics (the reference's NFKC
 *     case folding + "NFKD; [:Nonspacing Mark:] Remove; Latin-ASCII") for
 *     code points of up to two UTF-8 bytes: Latin-1, Latin Extended-A,
 *     Greek, Cyrillic, combining marks (ref src/tests/t_utf8.c:70-74,
 *     124-127); no compatibility decomposition beyond that;
 *   - "stopwords" drops words listed in <basedir>/filters/stopwords/<lang>
 *     (one per line), exactly where the reference looks for them
 *     (filters_builtin.c:93-127);
 *   - "stemmer" is the Snowball english algorithm, restated in stem_en.c, for
 *     lang = "en"; any other language is refused (NXS_ERR_INVALID) unless
 *     NXSB_STEMMER_PASSTHROUGH=1 asks for the identity.
 * Token-set semantics (dedup by string, first-seen order, per-token counts)
 * are the reference's (tokenizer.c:94-117).
 */
#ifndef NXSB_TOKENIZER_H
#define NXSB_TOKENIZER_H

#include "nxs_impl.h"

typedef enum { FILT_NORMALIZER, FILT_STOPWORDS, FILT_STEMMER } filter_kind_t;

typedef struct {
	unsigned	count;
	filter_kind_t	kinds[8];
	strmap_t *	stopwords;	/* NULL = none for this language */
	bool		stem_english;	/* "stemmer" with lang = en; else it is the identity */
} filter_pipeline_t;

/* The Snowball "english" stemmer over a lower-case word in place; returns the new length. */
size_t		stem_english(char *w, size_t len);

typedef struct {
	char *		str;
	uint32_t	len;
	uint32_t	count;		/* occurrences */
	uint32_t	term_id;	/* 0 = unresolved */
} token_t;

typedef struct {
	token_t *	list;
	uint32_t	count, cap;
	uint32_t	seen;		/* all occurrences, incl. duplicates */
	size_t		data_len;
	strmap_t *	map;		/* string -> list index */
} tokenset_t;

filter_pipeline_t *filter_pipeline_create(nxs_t *, nxs_params_t *);
void		filter_pipeline_destroy(filter_pipeline_t *);

tokenset_t *	tokenset_create(void);
void		tokenset_destroy(tokenset_t *);

/* The pipeline over one word in place: 1 = keep, 0 = discarded. */
int		filter_apply(const filter_pipeline_t *, char *buf, size_t *len);

/*
 * Run one value through the pipeline and add it to the set.  *slot gets the
 * token's index in the set, or -1 if a filter discarded it.  -1 on error.
 */
int		tokenize_value(filter_pipeline_t *, tokenset_t *,
		    const char *val, size_t len, int32_t *slot);
tokenset_t *	tokenize(filter_pipeline_t *, const char *text, size_t len);

#endif
