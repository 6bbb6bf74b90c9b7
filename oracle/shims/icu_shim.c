/*
 * ORACLE SHIM (test infrastructure, not product code).
 *
 * Replaces the two pieces of ICU the reference's indexing/query front end
 * needs -- UBRK_WORD segmentation (core/tokenizer.c:234-302) and the
 * UTF-8 <-> UTF-16 helpers of utils/utf8.c:112-190 -- with an ASCII-only
 * equivalent: the UAX #29 word-break rules over the ASCII classes (letters,
 * digits and '_' join; letter (.|') letter and digit (.|'|,|;) digit join;
 * the colon does not, as in ICU's tailoring), non-ASCII units taken as
 * letters.  A segment without a letter or a digit -- and everything else --
 * is a UBRK_WORD_NONE segment.
 */
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

#include <unicode/utypes.h>
#include <unicode/ubrk.h>

#include "strbuf.h"
#include "utf8.h"

struct UBreakIterator {
	const UChar *	text;
	int32_t		len;
	int32_t		pos;
	int32_t		status;
};

enum { WC_OTHER, WC_LETTER, WC_DIGIT, WC_EXTNUMLET, WC_MIDNUMLET, WC_MIDNUM };

static int
word_class(UChar c)
{
	if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c >= 0x80)
		return WC_LETTER;
	if (c >= '0' && c <= '9')
		return WC_DIGIT;
	if (c == '_')
		return WC_EXTNUMLET;
	if (c == '.' || c == '\'')
		return WC_MIDNUMLET;
	if (c == ',' || c == ';')
		return WC_MIDNUM;
	return WC_OTHER;
}

const char *
u_errorName(UErrorCode ec)
{
	return ec ? "U_SHIM_ERROR" : "U_ZERO_ERROR";
}

UBreakIterator *
ubrk_open(int type, const char *locale, const UChar *text, int32_t len,
    UErrorCode *ec)
{
	UBreakIterator *it = calloc(1, sizeof(*it));

	(void)type; (void)locale;
	if (!it) {
		*ec = 7;
		return NULL;
	}
	if (len < 0) {
		for (len = 0; text[len]; len++)
			;
	}
	it->text = text;
	it->len = len;
	return it;
}

void
ubrk_close(UBreakIterator *it)
{
	free(it);
}

int32_t
ubrk_first(UBreakIterator *it)
{
	it->pos = 0;
	return 0;
}

int32_t
ubrk_next(UBreakIterator *it)
{
	int32_t p = it->pos;

	if (p >= it->len) {
		return UBRK_DONE;
	}
	const int first = word_class(it->text[p]);

	if (first == WC_LETTER || first == WC_DIGIT || first == WC_EXTNUMLET) {
		int last = WC_OTHER, word = 0;

		while (p < it->len) {
			const int c = word_class(it->text[p]);
			const int next = p + 1 < it->len ? word_class(it->text[p + 1]) : WC_OTHER;

			if (c == WC_LETTER || c == WC_DIGIT || c == WC_EXTNUMLET) {
				word |= c != WC_EXTNUMLET;
				last = c;
			} else if (!((c == WC_MIDNUMLET && last == next &&
			    (last == WC_LETTER || last == WC_DIGIT)) ||
			    (c == WC_MIDNUM && last == WC_DIGIT && next == WC_DIGIT))) {
				break;
			}
			p++;
		}
		it->status = word ? UBRK_WORD_LETTER : UBRK_WORD_NONE;
	} else {
		p++;
		it->status = UBRK_WORD_NONE;
	}
	it->pos = p;
	return p;
}

int32_t
ubrk_getRuleStatus(UBreakIterator *it)
{
	return it->status;
}

/*
 * UTF-8 helpers with the signatures of utils/utf8.h.  Proper decoding of
 * multi-byte sequences into UTF-16 (and back) so that non-ASCII text at
 * least round-trips; no normalisation is attempted.
 */

ssize_t
utf8_to_utf16(utf8_ctx_t *ctx, const char *u8, uint16_t *buf, size_t count)
{
	const unsigned char *s = (const unsigned char *)u8;
	size_t n = 0;

	(void)ctx;
	while (*s) {
		uint32_t cp;
		unsigned extra;

		if (*s < 0x80) { cp = *s; extra = 0; }
		else if ((*s & 0xe0) == 0xc0) { cp = *s & 0x1f; extra = 1; }
		else if ((*s & 0xf0) == 0xe0) { cp = *s & 0x0f; extra = 2; }
		else if ((*s & 0xf8) == 0xf0) { cp = *s & 0x07; extra = 3; }
		else { cp = 0xfffd; extra = 0; }
		s++;
		while (extra-- && (*s & 0xc0) == 0x80) {
			cp = (cp << 6) | (*s++ & 0x3f);
		}
		if (cp >= 0x10000) {
			if (n + 2 >= count) return -1;
			cp -= 0x10000;
			buf[n++] = 0xd800 | (cp >> 10);
			buf[n++] = 0xdc00 | (cp & 0x3ff);
		} else {
			if (n + 1 >= count) return -1;
			buf[n++] = cp;
		}
	}
	if (n >= count) return -1;
	buf[n] = 0;
	return n;
}

ssize_t
utf8_from_utf16_new(utf8_ctx_t *ctx, const uint16_t *u16, size_t count,
    strbuf_t *buf)
{
	char tmp[4], *out;
	size_t cap = count * 3 + 1, n = 0;

	(void)ctx;
	if ((out = malloc(cap)) == NULL) {
		return -1;
	}
	for (size_t i = 0; i < count; i++) {
		uint32_t cp = u16[i];
		unsigned l;

		if (cp >= 0xd800 && cp < 0xdc00 && i + 1 < count) {
			cp = 0x10000 + ((cp - 0xd800) << 10) + (u16[++i] - 0xdc00);
		}
		if (cp < 0x80) { tmp[0] = cp; l = 1; }
		else if (cp < 0x800) {
			tmp[0] = 0xc0 | (cp >> 6); tmp[1] = 0x80 | (cp & 0x3f); l = 2;
		} else if (cp < 0x10000) {
			tmp[0] = 0xe0 | (cp >> 12); tmp[1] = 0x80 | ((cp >> 6) & 0x3f);
			tmp[2] = 0x80 | (cp & 0x3f); l = 3;
		} else {
			tmp[0] = 0xf0 | (cp >> 18); tmp[1] = 0x80 | ((cp >> 12) & 0x3f);
			tmp[2] = 0x80 | ((cp >> 6) & 0x3f); tmp[3] = 0x80 | (cp & 0x3f);
			l = 4;
		}
		if (n + l >= cap) {
			cap = cap * 2 + 8;
			out = realloc(out, cap);
		}
		memcpy(out + n, tmp, l);
		n += l;
	}
	out[n] = '\0';
	if (strbuf_acquire(buf, out, n) == -1) {
		free(out);
		return -1;
	}
	free(out);
	return n;
}
