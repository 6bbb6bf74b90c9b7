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
		 * libstemmer is not in this build; the English algorithm is
		 * restated in stem_en.c.  For any other language a pipeline that
		 * names the stemmer would index and query UNSTEMMED terms under
		 * parameters that promise stemmed ones (and miss the terms of an
		 * index the reference built): an error, not a silent
		 * pass-through.  NXSB_STEMMER_PASSTHROUGH=1 asks for the identity
		 * knowingly, whatever the language (what the tests' compiled
		 * reference is built with).
		 */
		if (kind == FILT_STEMMER) {
			const char *ok = getenv("NXSB_STEMMER_PASSTHROUGH");
			const char *lang = nxs_params_get_str(params, "lang");
			static bool warned;

			if (ok && strcmp(ok, "1") == 0) {
				if (!warned) {
					warned = true;
					fprintf(stderr, "nxsearch-b200: warning: the `stemmer' filter "
					    "passes terms through unchanged (NXSB_STEMMER_PASSTHROUGH=1)\n");
				}
			} else if (!lang || strcmp(lang, "en") == 0 || strcmp(lang, "english") == 0) {
				fp->stem_english = true;
			} else {
				nxs_set_error(nxs, NXS_ERR_INVALID,
				    "filter `stemmer' is not available for language `%s' in this "
				    "build (no libstemmer; English only); drop it from `filters' "
				    "or set NXSB_STEMMER_PASSTHROUGH=1 to index unstemmed terms", lang);
				free(names);
				strmap_destroy(fp->stopwords);
				free(fp);
				return NULL;
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
 * Simple lower-case mapping of a code point below U+0800, 0 = unchanged:
 * the Latin-1 Supplement, Latin Extended-A, Greek and Cyrillic blocks
 * (ref src/tests/t_utf8.c:73-74, "ĄČĘĖĮŠŲŪŽ" -> "ąčęėįšųūž").
 */
static uint32_t
lower_2byte(uint32_t cp)
{
	if ((cp >= 0xc0 && cp <= 0xde && cp != 0xd7) ||		/* Latin-1 */
	    (cp >= 0x391 && cp <= 0x3ab && cp != 0x3a2) ||	/* Greek */
	    (cp >= 0x410 && cp <= 0x42f))			/* Cyrillic */
		return cp + 0x20;
	if (cp >= 0x400 && cp <= 0x40f)
		return cp + 0x50;
	if (cp == 0x178)
		return 0xff;
	if (cp == 0x386)
		return 0x3ac;
	if (cp >= 0x388 && cp <= 0x38a)
		return cp + 0x25;
	if (cp == 0x38c)
		return 0x3cc;
	if (cp == 0x38e || cp == 0x38f)
		return cp + 0x3f;
	/* Pairs (upper, upper + 1): even upper ... */
	if (((cp >= 0x100 && cp <= 0x12f) || (cp >= 0x132 && cp <= 0x137) ||
	    (cp >= 0x14a && cp <= 0x177) || (cp >= 0x460 && cp <= 0x481) ||
	    (cp >= 0x48a && cp <= 0x4bf) || (cp >= 0x4d0 && cp <= 0x4ff)) && !(cp & 1))
		return cp + 1;
	/* ... and odd upper. */
	if (((cp >= 0x139 && cp <= 0x148) || (cp >= 0x179 && cp <= 0x17e) ||
	    (cp >= 0x4c1 && cp <= 0x4ce)) && (cp & 1))
		return cp + 1;
	return 0;
}

/*
 * What the reference's transliteration "NFKD; [:Nonspacing Mark:] Remove;
 * Latin-ASCII; NFKC" (ref src/utils/utf8.c:30-31) leaves of a lower-case code
 * point below U+0800: an ASCII string (returned through *ascii), another code
 * point (Greek without tonos, Cyrillic without breve / diaeresis), 0 for a
 * combining mark (removed), or the code point itself.
 */
static uint32_t
fold_2byte(uint32_t cp, const char **ascii)
{
	static const char *const latin1[32] = {	/* U+00E0 .. U+00FF */
		"a", "a", "a", "a", "a", "a", "ae", "c", "e", "e", "e", "e", "i", "i", "i", "i",
		"d", "n", "o", "o", "o", "o", "o", NULL, "o", "u", "u", "u", "u", "y", "th", "y",
	};
	static const char *const ext_a[128] = {	/* U+0100 .. U+017F, either case */
		"a", "a", "a", "a", "a", "a", "c", "c", "c", "c", "c", "c", "c", "c", "d", "d",
		"d", "d", "e", "e", "e", "e", "e", "e", "e", "e", "e", "e", "g", "g", "g", "g",
		"g", "g", "g", "g", "h", "h", "h", "h", "i", "i", "i", "i", "i", "i", "i", "i",
		"i", "i", "ij", "ij", "j", "j", "k", "k", "q", "l", "l", "l", "l", "l", "l", NULL,
		NULL, "l", "l", "n", "n", "n", "n", "n", "n", "'n", "ng", "ng", "o", "o", "o", "o",
		"o", "o", "oe", "oe", "r", "r", "r", "r", "r", "r", "s", "s", "s", "s", "s", "s",
		"s", "s", "t", "t", "t", "t", "t", "t", "u", "u", "u", "u", "u", "u", "u", "u",
		"u", "u", "u", "u", "w", "w", "y", "y", "y", "z", "z", "z", "z", "z", "z", "s",
	};

	*ascii = NULL;
	if (cp >= 0xe0 && cp <= 0xff)
		*ascii = latin1[cp - 0xe0];
	else if (cp >= 0x100 && cp <= 0x17f)
		*ascii = ext_a[cp - 0x100];
	else if (cp == 0xdf)
		*ascii = "ss";
	else if (cp == 0xaa)
		*ascii = "a";
	else if (cp == 0xba)
		*ascii = "o";
	else if (cp == 0xb2 || cp == 0xb3)
		*ascii = cp == 0xb2 ? "2" : "3";
	else if (cp == 0xb9)
		*ascii = "1";
	if (*ascii)
		return cp;
	if (cp >= 0x300 && cp <= 0x36f)
		return 0;
	switch (cp) {
	case 0xb5:  return 0x3bc;			/* micro sign -> mu */
	case 0x3ac: return 0x3b1;
	case 0x3ad: return 0x3b5;
	case 0x3ae: return 0x3b7;
	case 0x3af: case 0x3ca: case 0x390: return 0x3b9;
	case 0x3cc: return 0x3bf;
	case 0x3cd: case 0x3cb: case 0x3b0: return 0x3c5;
	case 0x3ce: return 0x3c9;
	case 0x3c2: return 0x3c3;			/* final sigma (case folding) */
	case 0x439: case 0x45d: return 0x438;
	case 0x451: case 0x450: return 0x435;
	case 0x457: return 0x456;
	case 0x453: return 0x433;
	case 0x45c: return 0x43a;
	case 0x45e: return 0x443;
	}
	return cp;
}

/*
 * "normalizer": the reference applies NFKC case folding and then strips
 * diacritics with the transliteration above (ref src/core/filters_builtin.c:
 * 55-74, src/utils/utf8.c:263-330,212-255), all through ICU.  Here, in place
 * (the result is never longer): ASCII is lower-cased; code points of two
 * UTF-8 bytes -- Latin-1, Latin Extended-A, Greek, Cyrillic, the combining
 * marks -- are lower-cased and folded as the reference's golden cases show
 * (ref src/tests/t_utf8.c:70-74,124-127); of the longer sequences the curly
 * apostrophes, the full-width letters and digits and the Latin ligatures
 * fold to ASCII, everything else is kept as it is (no general compatibility
 * decomposition: "Ⅷ" stays, since "viii" would not fit in place).
 */
static size_t
normalize_inplace(unsigned char *buf, size_t len)
{
	size_t o = 0;

	for (size_t k = 0; k < len; k++) {
		const unsigned char c = buf[k];

		if (c >= 'A' && c <= 'Z') {
			buf[o++] = c + ('a' - 'A');
		} else if ((c & 0xe0) == 0xc0 && k + 1 < len && (buf[k + 1] & 0xc0) == 0x80) {
			uint32_t cp = ((uint32_t)(c & 0x1f) << 6) | (buf[k + 1] & 0x3f);
			const uint32_t lo = lower_2byte(cp);
			const char *ascii;

			cp = fold_2byte(lo ? lo : cp, &ascii);
			k++;
			if (ascii) {
				while (*ascii)
					buf[o++] = (unsigned char)*ascii++;
			} else if (cp) {
				buf[o++] = 0xc0 | (cp >> 6);
				buf[o++] = 0x80 | (cp & 0x3f);
			}
		} else if (c == 0xe2 && k + 2 < len && buf[k + 1] == 0x80 &&
		    (buf[k + 2] == 0x98 || buf[k + 2] == 0x99 || buf[k + 2] == 0xa4)) {
			/* Latin-ASCII: the curly apostrophes and the one-dot leader inside a word */
			buf[o++] = buf[k + 2] == 0xa4 ? '.' : '\'';
			k += 2;
		} else if (c == 0xef && k + 2 < len && (buf[k + 1] == 0xbc || buf[k + 1] == 0xbd) &&
		    (buf[k + 2] & 0xc0) == 0x80) {
			/* NFKC case folding of the full-width forms: U+FF10-19, U+FF21-3A, U+FF41-5A */
			const uint32_t cp = 0xff00u | ((uint32_t)(buf[k + 1] & 0x03) << 6) | (buf[k + 2] & 0x3f);

			if (cp >= 0xff10 && cp <= 0xff19) {
				buf[o++] = (unsigned char)('0' + (cp - 0xff10));
			} else if (cp >= 0xff21 && cp <= 0xff3a) {
				buf[o++] = (unsigned char)('a' + (cp - 0xff21));
			} else if (cp >= 0xff41 && cp <= 0xff5a) {
				buf[o++] = (unsigned char)('a' + (cp - 0xff41));
			} else {
				buf[o++] = c;
				buf[o++] = buf[k + 1];
				buf[o++] = buf[k + 2];
			}
			k += 2;
		} else if (c == 0xef && k + 2 < len && buf[k + 1] == 0xac && buf[k + 2] >= 0x80 &&
		    buf[k + 2] <= 0x86) {
			/* ... and of the Latin ligatures U+FB00-06 */
			static const char *const lig[7] = { "ff", "fi", "fl", "ffi", "ffl", "st", "st" };

			for (const char *p = lig[buf[k + 2] - 0x80]; *p; p++)
				buf[o++] = (unsigned char)*p;
			k += 2;
		} else {
			buf[o++] = c;
		}
	}
	return o;
}

/*
 * The filter pipeline over one word, in place.  Returns 1 to keep the token,
 * 0 if a filter discarded it (or left it empty: ref filters.c:206-208).
 */
int
filter_apply(const filter_pipeline_t *fp, char *buf, size_t *lenp)
{
	size_t len = *lenp;

	for (unsigned i = 0; i < fp->count; i++) {
		switch (fp->kinds[i]) {
		case FILT_NORMALIZER:
			len = normalize_inplace((unsigned char *)buf, len);
			buf[len] = '\0';
			break;
		case FILT_STOPWORDS:
			if (fp->stopwords && strmap_get(fp->stopwords, buf, len, NULL))
				return 0;
			break;
		case FILT_STEMMER:
			if (fp->stem_english) {	/* else the identity (NXSB_STEMMER_PASSTHROUGH=1) */
				len = stem_english(buf, len);
				buf[len] = '\0';
			}
			break;
		}
		if (len == 0)
			return 0;
	}
	*lenp = len;
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

/*
 * Word segmentation: the reference walks ICU's UBRK_WORD boundaries and keeps
 * the segments whose rule status is not UBRK_WORD_NONE (ref
 * src/core/tokenizer.c:234-302).  These are the UAX #29 word-break rules with
 * the character classes that need no Unicode tables (ICU's tailoring: the
 * colon is not MidLetter):
 *   WB5, WB8-10, WB13, WB13a/b   letters, digits and '_' join;
 *   WB6/7    letter ( . | ' | U+2019 | U+00B7 ... ) letter joins ("i.b.m", "doesn't");
 *   WB11/12  digit ( . | ' | , | ; ) digit joins ("3.14", "1,000").
 * Outside ASCII a code point is a letter unless it lies in one of the blocks
 * that hold no letters -- Latin-1 punctuation, General Punctuation (curly
 * quotes, dashes, the ellipsis, typographic spaces), currency signs, arrows
 * and mathematical / technical / box / dingbat symbols, CJK and full-width
 * punctuation, emoji (Extended_Pictographic has no word status in ICU) --
 * where it breaks words like ASCII punctuation does.  Scripts that ICU
 * segments with a dictionary (CJK, Thai) stay one word per run.  A segment
 * without a letter or a digit is not a word.  Pinned by the reference's
 * golden cases (ref src/tests/t_tokenize.c:17-62, and :64-79 for the emoji).
 */
enum { WC_OTHER, WC_LETTER, WC_DIGIT, WC_EXTNUMLET, WC_MIDNUMLET, WC_MIDNUM, WC_MIDLETTER };

static int
class_of(uint32_t cp)
{
	if (cp < 0x80) {
		if ((cp >= 'a' && cp <= 'z') || (cp >= 'A' && cp <= 'Z'))
			return WC_LETTER;
		if (cp >= '0' && cp <= '9')
			return WC_DIGIT;
		if (cp == '_')
			return WC_EXTNUMLET;
		if (cp == '.' || cp == '\'')
			return WC_MIDNUMLET;
		if (cp == ',' || cp == ';')
			return WC_MIDNUM;
		return WC_OTHER;
	}
	if (cp < 0xc0)		/* C1 controls, Latin-1 punctuation and signs */
		return cp == 0xaa || cp == 0xb5 || cp == 0xba ? WC_LETTER :
		    cp == 0xb7 ? WC_MIDLETTER : WC_OTHER;
	if (cp == 0xd7 || cp == 0xf7)
		return WC_OTHER;
	if (cp < 0x2000)
		return WC_LETTER;
	if (cp <= 0x206f) {	/* General Punctuation */
		if (cp == 0x2018 || cp == 0x2019 || cp == 0x2024)
			return WC_MIDNUMLET;
		if (cp == 0x2027)
			return WC_MIDLETTER;
		if (cp == 0x2044)
			return WC_MIDNUM;
		if (cp == 0x203f || cp == 0x2040 || cp == 0x2054)
			return WC_EXTNUMLET;
		if (cp == 0x200c || cp == 0x200d)	/* joiners are transparent (WB4) */
			return WC_LETTER;
		return WC_OTHER;
	}
	if ((cp >= 0x20a0 && cp <= 0x20cf) ||	/* currency */
	    (cp >= 0x2190 && cp <= 0x23ff) ||	/* arrows, mathematical, technical */
	    (cp >= 0x2460 && cp <= 0x24b5) ||	/* enclosed digits */
	    (cp >= 0x24ea && cp <= 0x27bf) ||	/* ..., box drawing, shapes, dingbats */
	    (cp >= 0x2900 && cp <= 0x2bff) ||	/* more arrows and operators */
	    (cp >= 0x2e00 && cp <= 0x2e7f) ||	/* supplemental punctuation */
	    (cp >= 0x3000 && cp <= 0x3004) || (cp >= 0x3008 && cp <= 0x3020) ||
	    cp == 0x3030 || (cp >= 0x303d && cp <= 0x303f) ||	/* CJK punctuation */
	    (cp >= 0xfe50 && cp <= 0xfe6b) ||	/* small form variants */
	    (cp >= 0xff01 && cp <= 0xff0f) || (cp >= 0xff1a && cp <= 0xff20) ||
	    (cp >= 0xff3b && cp <= 0xff40 && cp != 0xff3f) ||
	    (cp >= 0xff5b && cp <= 0xff65) ||	/* full-width punctuation */
	    cp == 0xfffd || (cp >= 0x1f000 && cp <= 0x1faff))	/* emoji */
		return WC_OTHER;
	return WC_LETTER;
}

/*
 * The class of the code point at text[i] and its length in bytes; the end of
 * the text (or a NUL) is WC_OTHER of length 0.  A byte that does not start a
 * well-formed sequence counts as a letter of one byte.
 */
static int
class_at(const unsigned char *t, size_t i, size_t len, size_t *adv)
{
	uint32_t cp;
	size_t n;

	*adv = 0;
	if (i >= len || t[i] == 0)
		return WC_OTHER;
	*adv = 1;
	if (t[i] < 0x80)
		return class_of(t[i]);
	n = (t[i] & 0xe0) == 0xc0 ? 2 : (t[i] & 0xf0) == 0xe0 ? 3 : (t[i] & 0xf8) == 0xf0 ? 4 : 1;
	if (n == 1 || i + n > len)
		return WC_LETTER;
	cp = t[i] & (0xffu >> (n + 1));
	for (size_t k = 1; k < n; k++) {
		if ((t[i + k] & 0xc0) != 0x80)
			return WC_LETTER;
		cp = (cp << 6) | (t[i + k] & 0x3f);
	}
	*adv = n;
	return class_of(cp);
}

tokenset_t *
tokenize(filter_pipeline_t *fp, const char *text, size_t len)
{
	const unsigned char *t = (const unsigned char *)text;
	tokenset_t *ts = tokenset_create();
	size_t i = 0;

	if (!ts)
		return NULL;
	while (i < len && t[i]) {
		bool word = false;
		int last = WC_OTHER;
		size_t s, adv, nadv;
		int32_t slot;

		for (;;) {
			const int c = class_at(t, i, len, &adv);

			if (adv == 0 || c == WC_LETTER || c == WC_DIGIT || c == WC_EXTNUMLET)
				break;
			i += adv;
		}
		s = i;
		for (;;) {
			const int c = class_at(t, i, len, &adv);

			if (adv == 0)
				break;
			if (c == WC_LETTER || c == WC_DIGIT || c == WC_EXTNUMLET) {
				word |= c != WC_EXTNUMLET;
				last = c;
			} else {
				const int next = class_at(t, i + adv, len, &nadv);

				if (!((c == WC_MIDNUMLET && last == next &&
				    (last == WC_LETTER || last == WC_DIGIT)) ||
				    (c == WC_MIDLETTER && last == WC_LETTER && next == WC_LETTER) ||
				    (c == WC_MIDNUM && last == WC_DIGIT && next == WC_DIGIT)))
					break;
			}
			i += adv;
		}
		if (word && tokenize_value(fp, ts, text + s, i - s, &slot) == -1) {
			tokenset_destroy(ts);
			return NULL;
		}
	}
	return ts;
}
