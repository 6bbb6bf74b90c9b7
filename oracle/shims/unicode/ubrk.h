/*
 * ORACLE SHIM: stand-in for <unicode/ubrk.h>.  The break iterator is an
 * ASCII word segmenter (see ../icu_shim.c); synthetic corpora use lowercase
 * ASCII alphanumerics only, where it cannot diverge from real UAX#29.
 */
#ifndef NXSB_ORACLE_SHIM_UBRK_H
#define NXSB_ORACLE_SHIM_UBRK_H

#include "utypes.h"

typedef struct UBreakIterator UBreakIterator;

#define UBRK_WORD	1
#define UBRK_DONE	((int32_t)-1)
#define UBRK_WORD_NONE	0
#define UBRK_WORD_LETTER 200

UBreakIterator *ubrk_open(int, const char *, const UChar *, int32_t, UErrorCode *);
void	ubrk_close(UBreakIterator *);
int32_t	ubrk_first(UBreakIterator *);
int32_t	ubrk_next(UBreakIterator *);
int32_t	ubrk_getRuleStatus(UBreakIterator *);

#endif
