/*
 * Open-addressing hash maps used by the host side: byte-string -> u32 (the
 * term map, reference: idx->term_map in src/index/idxterm.c:45-53, and the
 * token de-duplication of src/core/tokenizer.c:94-117) and u64 -> u32 (the
 * document map, reference: idx->dt_map in src/index/idxdoc.c:29-75).
 */
#ifndef NXSB_HASHMAP_H
#define NXSB_HASHMAP_H

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

typedef struct strmap strmap_t;

strmap_t *	strmap_create(size_t hint);
void		strmap_destroy(strmap_t *);
size_t		strmap_count(const strmap_t *);
int		strmap_reserve(strmap_t *, size_t n);
/*
 * Insert key -> val unless present.  Returns 1 if inserted, 0 if the key was
 * already there (*cur gets its value), -1 on OOM.  The key bytes are copied.
 */
int		strmap_put(strmap_t *, const void *key, size_t len,
		    uint32_t val, uint32_t *cur);
bool		strmap_get(const strmap_t *, const void *key, size_t len,
		    uint32_t *val);

typedef struct u64map u64map_t;

u64map_t *	u64map_create(size_t hint);
void		u64map_destroy(u64map_t *);
size_t		u64map_count(const u64map_t *);
int		u64map_reserve(u64map_t *, size_t n);
void		u64map_prefetch(const u64map_t *, uint64_t key);
int		u64map_put(u64map_t *, uint64_t key, uint32_t val, uint32_t *cur);
bool		u64map_get(const u64map_t *, uint64_t key, uint32_t *val);
bool		u64map_del(u64map_t *, uint64_t key);

uint64_t	nxsb_hash_bytes(const void *, size_t);

#endif
