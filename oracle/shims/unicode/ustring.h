/* ORACLE SHIM: stand-in for <unicode/ustring.h>; nothing is needed from it. */
#ifndef NXSB_ORACLE_SHIM_USTRING_H
#define NXSB_ORACLE_SHIM_USTRING_H
#include "utypes.h"
#endif
