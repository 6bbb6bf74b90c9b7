/*
 * ORACLE SHIM (test infrastructure, not product code).
 *
 * Minimal stand-in for ICU's <unicode/utypes.h> so that the reference's
 * core/tokenizer.c compiles unmodified in an image without ICU headers.
 * Only the handful of names tokenizer.c touches are declared.
 */
#ifndef NXSB_ORACLE_SHIM_UTYPES_H
#define NXSB_ORACLE_SHIM_UTYPES_H

#include <stdint.h>

typedef uint16_t UChar;
typedef int UErrorCode;

#define U_ZERO_ERROR	0
#define U_FAILURE(ec)	((ec) > 0)

const char *u_errorName(UErrorCode);

#endif
