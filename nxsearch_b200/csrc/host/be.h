/* Big-endian accessors for the on-disk index (ref src/utils/mmrw.c). */
#ifndef NXSB_BE_H
#define NXSB_BE_H

#include <endian.h>
#include <stdint.h>
#include <string.h>

static inline void be_put16(uint8_t *p, uint16_t v) { v = htobe16(v); memcpy(p, &v, 2); }
static inline void be_put32(uint8_t *p, uint32_t v) { v = htobe32(v); memcpy(p, &v, 4); }
static inline void be_put64(uint8_t *p, uint64_t v) { v = htobe64(v); memcpy(p, &v, 8); }
static inline uint16_t be_get16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return be16toh(v); }
static inline uint32_t be_get32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return be32toh(v); }
static inline uint64_t be_get64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return be64toh(v); }

#endif
