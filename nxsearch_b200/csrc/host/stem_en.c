/*
 * The Snowball "english" (Porter2) stemmer.
 *
 * The reference's `stemmer' filter calls libstemmer (sb_stemmer_new(lang) /
 * sb_stemmer_stem, ref src/core/filters_builtin.c:207-239), a third-party
 * library that is neither vendored in the reference nor present in this
 * image.  This file restates the published algorithm for English
 * (snowballstem.org, algorithms/english/stemmer.html, the definition that
 * libstemmer 2.0 compiles from english.sbl); other languages stay refused.
 * Anchors: the reference's scoring goldens that need it ("fox" finds
 * "foxes", ref src/tests/t_scoring.c:19-68) and the sample vocabulary of
 * the published description (tests/test_host.py).
 *
 * OUT OF THE ACCELERATED PATH: host C, one word at a time.
 */
#include <stdbool.h>
#include <stddef.h>
#include <string.h>

#include "tokenizer.h"

typedef struct {
	char *	w;
	size_t	n;	/* current length */
	size_t	p1, p2;	/* start of R1, R2 */
} stem_t;

static inline bool
is_v(char c)
{
	return c == 'a' || c == 'e' || c == 'i' || c == 'o' || c == 'u' || c == 'y';
}

static inline bool
ends(const stem_t *z, const char *s, size_t l)
{
	return z->n >= l && memcmp(z->w + z->n - l, s, l) == 0;
}

#define ENDS(z, lit)	ends((z), (lit), sizeof(lit) - 1)

/* Replace the last `cut' bytes by s (never longer than what it replaces). */
static inline void
put(stem_t *z, size_t cut, const char *s)
{
	const size_t l = strlen(s);

	memcpy(z->w + z->n - cut, s, l);
	z->n = z->n - cut + l;
}

/* The first position after a vowel followed by a non-vowel, searching from `from'. */
static size_t
region_after(const stem_t *z, size_t from)
{
	size_t i = from;

	while (i < z->n && !is_v(z->w[i]))
		i++;
	while (i < z->n && is_v(z->w[i]))
		i++;
	return i < z->n ? i + 1 : z->n;
}

/* Does the word, cut at `end', end in a short syllable? */
static bool
short_syllable(const stem_t *z, size_t end)
{
	const char *w = z->w;

	if (end >= 3 && !is_v(w[end - 1]) && w[end - 1] != 'w' && w[end - 1] != 'x' &&
	    w[end - 1] != 'Y' && is_v(w[end - 2]) && !is_v(w[end - 3]))
		return true;
	return end == 2 && is_v(w[0]) && !is_v(w[1]);
}

static bool
has_vowel(const stem_t *z, size_t end)
{
	for (size_t i = 0; i < end; i++)
		if (is_v(z->w[i]))
			return true;
	return false;
}

/* The longest of the suffixes in list[] the word ends with, or -1. */
static int
longest(const stem_t *z, const char *const *list, int count, size_t *len)
{
	int best = -1;

	*len = 0;
	for (int i = 0; i < count; i++) {
		const size_t l = strlen(list[i]);

		if (l > *len && ends(z, list[i], l)) {
			best = i;
			*len = l;
		}
	}
	return best;
}

static bool
exception1(stem_t *z)
{
	static const char *const from[] = {
		"skis", "skies", "dying", "lying", "tying", "idly", "gently", "ugly",
		"early", "only", "singly", "sky", "news", "howe", "atlas", "cosmos",
		"bias", "andes",
	};
	static const char *const to[] = {
		"ski", "sky", "die", "lie", "tie", "idl", "gentl", "ugli",
		"earli", "onli", "singl", "sky", "news", "howe", "atlas", "cosmos",
		"bias", "andes",
	};

	for (unsigned i = 0; i < sizeof(from) / sizeof(from[0]); i++) {
		if (z->n == strlen(from[i]) && memcmp(z->w, from[i], z->n) == 0) {
			put(z, z->n, to[i]);
			return true;
		}
	}
	return false;
}

static bool
exception2(const stem_t *z)
{
	static const char *const same[] = {
		"inning", "outing", "canning", "herring", "earring", "proceed",
		"exceed", "succeed",
	};

	for (unsigned i = 0; i < sizeof(same) / sizeof(same[0]); i++)
		if (z->n == strlen(same[i]) && memcmp(z->w, same[i], z->n) == 0)
			return true;
	return false;
}

static void
step_1a(stem_t *z)
{
	if (ENDS(z, "'s'"))
		z->n -= 3;
	else if (ENDS(z, "'s"))
		z->n -= 2;
	else if (ENDS(z, "'"))
		z->n -= 1;

	if (ENDS(z, "sses")) {
		put(z, 4, "ss");
	} else if (ENDS(z, "ied") || ENDS(z, "ies")) {
		put(z, 3, z->n - 3 > 1 ? "i" : "ie");
	} else if (ENDS(z, "us") || ENDS(z, "ss")) {
		/* nothing */
	} else if (ENDS(z, "s")) {
		/* a vowel somewhere before the letter in front of the s */
		if (z->n >= 2 && has_vowel(z, z->n - 2))
			z->n -= 1;
	}
}

static void
step_1b(stem_t *z)
{
	static const char *const doubles[] = { "bb", "dd", "ff", "gg", "mm", "nn", "pp", "rr", "tt" };
	size_t l;

	if (ENDS(z, "eedly") || ENDS(z, "eed")) {
		l = ENDS(z, "eedly") ? 5 : 3;
		if (z->n - l >= z->p1)
			put(z, l, "ee");
		return;
	}
	if (ENDS(z, "ingly"))
		l = 5;
	else if (ENDS(z, "edly"))
		l = 4;
	else if (ENDS(z, "ing"))
		l = 3;
	else if (ENDS(z, "ed"))
		l = 2;
	else
		return;
	if (!has_vowel(z, z->n - l))
		return;
	z->n -= l;
	if (ENDS(z, "at") || ENDS(z, "bl") || ENDS(z, "iz")) {
		z->w[z->n++] = 'e';
		return;
	}
	for (unsigned i = 0; i < sizeof(doubles) / sizeof(doubles[0]); i++) {
		if (ends(z, doubles[i], 2)) {
			z->n -= 1;
			return;
		}
	}
	/* a short word: R1 is empty and the word ends in a short syllable */
	if (z->p1 == z->n && short_syllable(z, z->n))
		z->w[z->n++] = 'e';
}

static void
step_1c(stem_t *z)
{
	if (z->n >= 3 && (z->w[z->n - 1] == 'y' || z->w[z->n - 1] == 'Y') && !is_v(z->w[z->n - 2]))
		z->w[z->n - 1] = 'i';
}

static void
step_2(stem_t *z)
{
	static const char *const suf[] = {
		"tional", "enci", "anci", "abli", "entli", "izer", "ization", "ational",
		"ation", "ator", "alism", "aliti", "alli", "fulness", "ousli", "ousness",
		"iveness", "iviti", "biliti", "bli", "ogi", "fulli", "lessli", "li",
	};
	static const char *const rep[] = {
		"tion", "ence", "ance", "able", "ent", "ize", "ize", "ate",
		"ate", "ate", "al", "al", "al", "ful", "ous", "ous",
		"ive", "ive", "ble", "ble", "og", "ful", "less", "",
	};
	size_t l;
	const int k = longest(z, suf, sizeof(suf) / sizeof(suf[0]), &l);

	if (k < 0 || z->n - l < z->p1)
		return;
	if (strcmp(suf[k], "ogi") == 0) {
		if (z->n - l >= 1 && z->w[z->n - l - 1] == 'l')
			put(z, l, "og");
	} else if (strcmp(suf[k], "li") == 0) {
		if (z->n - l >= 1 && strchr("cdeghkmnrt", z->w[z->n - l - 1]) != NULL)
			z->n -= 2;
	} else {
		put(z, l, rep[k]);
	}
}

static void
step_3(stem_t *z)
{
	static const char *const suf[] = {
		"tional", "ational", "alize", "icate", "iciti", "ical", "ful", "ness", "ative",
	};
	static const char *const rep[] = {
		"tion", "ate", "al", "ic", "ic", "ic", "", "", "",
	};
	size_t l;
	const int k = longest(z, suf, sizeof(suf) / sizeof(suf[0]), &l);

	if (k < 0 || z->n - l < z->p1)
		return;
	if (strcmp(suf[k], "ative") == 0) {
		if (z->n - l >= z->p2)
			z->n -= l;
	} else {
		put(z, l, rep[k]);
	}
}

static void
step_4(stem_t *z)
{
	static const char *const suf[] = {
		"al", "ance", "ence", "er", "ic", "able", "ible", "ant", "ement", "ment",
		"ent", "ism", "ate", "iti", "ous", "ive", "ize", "ion",
	};
	size_t l;
	const int k = longest(z, suf, sizeof(suf) / sizeof(suf[0]), &l);

	if (k < 0 || z->n - l < z->p2)
		return;
	if (strcmp(suf[k], "ion") == 0) {
		if (z->n - l >= 1 && (z->w[z->n - l - 1] == 's' || z->w[z->n - l - 1] == 't'))
			z->n -= l;
	} else {
		z->n -= l;
	}
}

static void
step_5(stem_t *z)
{
	if (z->n == 0)
		return;
	if (z->w[z->n - 1] == 'e') {
		if (z->n - 1 >= z->p2 || (z->n - 1 >= z->p1 && !short_syllable(z, z->n - 1)))
			z->n -= 1;
	} else if (z->w[z->n - 1] == 'l') {
		if (z->n - 1 >= z->p2 && z->n >= 2 && z->w[z->n - 2] == 'l')
			z->n -= 1;
	}
}

/*
 * Stem the lower-case word w[0 .. len) in place and return its new length
 * (never more than len).  Bytes outside [a-z'] count as non-vowels.
 */
size_t
stem_english(char *w, size_t len)
{
	stem_t z = { .w = w, .n = len };

	if (exception1(&z) || len < 3)
		return z.n;

	/* prelude: no leading apostrophe; y at the start or after a vowel is a consonant */
	if (w[0] == '\'') {
		memmove(w, w + 1, len - 1);
		z.n--;
	}
	if (z.n && w[0] == 'y')
		w[0] = 'Y';
	for (size_t i = 1; i < z.n; i++)
		if (w[i] == 'y' && is_v(w[i - 1]))
			w[i] = 'Y';

	/* R1 (with the three prefixes that overstem otherwise) and R2 */
	if (z.n >= 5 && (memcmp(w, "gener", 5) == 0 || memcmp(w, "arsen", 5) == 0))
		z.p1 = 5;
	else if (z.n >= 6 && memcmp(w, "commun", 6) == 0)
		z.p1 = 6;
	else
		z.p1 = region_after(&z, 0);
	z.p2 = region_after(&z, z.p1);

	step_1a(&z);
	if (!exception2(&z)) {
		step_1b(&z);
		step_1c(&z);
		step_2(&z);
		step_3(&z);
		step_4(&z);
		step_5(&z);
	}
	for (size_t i = 0; i < z.n; i++)
		if (w[i] == 'Y')
			w[i] = 'y';
	return z.n;
}
