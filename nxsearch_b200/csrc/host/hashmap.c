/*
 * Open-addressing (linear probing) hash maps; see hashmap.h.
 */
#include <stdlib.h>
#include <string.h>

#include "hashmap.h"

static inline uint64_t
mix64(uint64_t x)
{
	x ^= x >> 32;
	x *= UINT64_C(0xd6e8feb86659fd93);
	x ^= x >> 32;
	x *= UINT64_C(0xd6e8feb86659fd93);
	x ^= x >> 32;
	return x;
}

uint64_t
nxsb_hash_bytes(const void *key, size_t len)
{
	const uint8_t *p = key;
	uint64_t h = UINT64_C(0x9e3779b97f4a7c15) ^ (len * UINT64_C(0xff51afd7ed558ccd));

	while (len >= 8) {
		uint64_t w;
		memcpy(&w, p, 8);
		h = mix64(h ^ w) + UINT64_C(0x2545f4914f6cdd1d);
		p += 8, len -= 8;
	}
	if (len) {
		uint64_t w = 0;
		memcpy(&w, p, len);
		h = mix64(h ^ w ^ ((uint64_t)len << 56));
	}
	return mix64(h);
}

/*
 * String map.  Slots hold (hash, arena offset, length, value); key bytes
 * live in one growing arena so a million short terms cost one allocation.
 */

typedef struct {
	uint64_t	hash;
	uint64_t	off;
	uint32_t	len;
	uint32_t	val;
	uint8_t		used;
} strslot_t;

struct strmap {
	strslot_t *	slots;
	size_t		nslots;		// power of two
	size_t		count;
	char *		arena;
	size_t		arena_len, arena_cap;
};

static size_t
pow2_for(size_t hint)
{
	size_t n = 16;
	while (n < hint * 2)
		n <<= 1;
	return n;
}

strmap_t *
strmap_create(size_t hint)
{
	strmap_t *m = calloc(1, sizeof(*m));

	if (!m)
		return NULL;
	m->nslots = pow2_for(hint);
	m->slots = calloc(m->nslots, sizeof(strslot_t));
	if (!m->slots) {
		free(m);
		return NULL;
	}
	return m;
}

void
strmap_destroy(strmap_t *m)
{
	if (m) {
		free(m->slots);
		free(m->arena);
		free(m);
	}
}

size_t
strmap_count(const strmap_t *m)
{
	return m->count;
}

static int
strmap_grow(strmap_t *m)
{
	const size_t n = m->nslots * 2;
	strslot_t *ns = calloc(n, sizeof(strslot_t));

	if (!ns)
		return -1;
	for (size_t i = 0; i < m->nslots; i++) {
		const strslot_t *s = &m->slots[i];
		size_t j;

		if (!s->used)
			continue;
		for (j = s->hash & (n - 1); ns[j].used; j = (j + 1) & (n - 1))
			;
		ns[j] = *s;
	}
	free(m->slots);
	m->slots = ns;
	m->nslots = n;
	return 0;
}

/* Room for n keys without rehashing on the way (bulk loads). */
int
strmap_reserve(strmap_t *m, size_t n)
{
	while (n * 10 > m->nslots * 7)
		if (strmap_grow(m) == -1)
			return -1;
	return 0;
}

int
strmap_put(strmap_t *m, const void *key, size_t len, uint32_t val,
    uint32_t *cur)
{
	const uint64_t h = nxsb_hash_bytes(key, len);
	size_t i;

	if ((m->count + 1) * 10 > m->nslots * 7 && strmap_grow(m) == -1)
		return -1;

	for (i = h & (m->nslots - 1); m->slots[i].used;
	    i = (i + 1) & (m->nslots - 1)) {
		const strslot_t *s = &m->slots[i];

		if (s->hash == h && s->len == len &&
		    memcmp(m->arena + s->off, key, len) == 0) {
			if (cur)
				*cur = s->val;
			return 0;
		}
	}
	if (m->arena_len + len + 1 > m->arena_cap) {
		size_t ncap = m->arena_cap ? m->arena_cap * 2 : 4096;
		char *na;

		while (ncap < m->arena_len + len + 1)
			ncap *= 2;
		if ((na = realloc(m->arena, ncap)) == NULL)
			return -1;
		m->arena = na;
		m->arena_cap = ncap;
	}
	memcpy(m->arena + m->arena_len, key, len);
	m->arena[m->arena_len + len] = '\0';
	m->slots[i] = (strslot_t){
		.hash = h, .off = m->arena_len, .len = len, .val = val, .used = 1
	};
	m->arena_len += len + 1;
	m->count++;
	return 1;
}

bool
strmap_get(const strmap_t *m, const void *key, size_t len, uint32_t *val)
{
	const uint64_t h = nxsb_hash_bytes(key, len);

	for (size_t i = h & (m->nslots - 1); m->slots[i].used;
	    i = (i + 1) & (m->nslots - 1)) {
		const strslot_t *s = &m->slots[i];

		if (s->hash == h && s->len == len &&
		    memcmp(m->arena + s->off, key, len) == 0) {
			if (val)
				*val = s->val;
			return true;
		}
	}
	return false;
}

/*
 * u64 map with tombstone-free deletion (backward shift).
 */

typedef struct {
	uint64_t	key;
	uint32_t	val;
	uint32_t	used;
} u64slot_t;

struct u64map {
	u64slot_t *	slots;
	size_t		nslots;
	size_t		count;
};

u64map_t *
u64map_create(size_t hint)
{
	u64map_t *m = calloc(1, sizeof(*m));

	if (!m)
		return NULL;
	m->nslots = pow2_for(hint);
	m->slots = calloc(m->nslots, sizeof(u64slot_t));
	if (!m->slots) {
		free(m);
		return NULL;
	}
	return m;
}

void
u64map_destroy(u64map_t *m)
{
	if (m) {
		free(m->slots);
		free(m);
	}
}

size_t
u64map_count(const u64map_t *m)
{
	return m->count;
}

static int
u64map_grow(u64map_t *m)
{
	const size_t n = m->nslots * 2;
	u64slot_t *ns = calloc(n, sizeof(u64slot_t));

	if (!ns)
		return -1;
	for (size_t i = 0; i < m->nslots; i++) {
		const u64slot_t *s = &m->slots[i];
		size_t j;

		if (!s->used)
			continue;
		for (j = mix64(s->key) & (n - 1); ns[j].used; j = (j + 1) & (n - 1))
			;
		ns[j] = *s;
	}
	free(m->slots);
	m->slots = ns;
	m->nslots = n;
	return 0;
}

/* Room for n keys without growing on the way (bulk loads). */
int
u64map_reserve(u64map_t *m, size_t n)
{
	while (n * 10 > m->nslots * 7)
		if (u64map_grow(m) == -1)
			return -1;
	return 0;
}

/* Start fetching the home slot of a key that is about to be looked up. */
void
u64map_prefetch(const u64map_t *m, uint64_t key)
{
	__builtin_prefetch(&m->slots[mix64(key) & (m->nslots - 1)], 1, 0);
}

int
u64map_put(u64map_t *m, uint64_t key, uint32_t val, uint32_t *cur)
{
	size_t i;

	if ((m->count + 1) * 10 > m->nslots * 7 && u64map_grow(m) == -1)
		return -1;
	for (i = mix64(key) & (m->nslots - 1); m->slots[i].used;
	    i = (i + 1) & (m->nslots - 1)) {
		if (m->slots[i].key == key) {
			if (cur)
				*cur = m->slots[i].val;
			return 0;
		}
	}
	m->slots[i] = (u64slot_t){ .key = key, .val = val, .used = 1 };
	m->count++;
	return 1;
}

bool
u64map_get(const u64map_t *m, uint64_t key, uint32_t *val)
{
	for (size_t i = mix64(key) & (m->nslots - 1); m->slots[i].used;
	    i = (i + 1) & (m->nslots - 1)) {
		if (m->slots[i].key == key) {
			if (val)
				*val = m->slots[i].val;
			return true;
		}
	}
	return false;
}

bool
u64map_del(u64map_t *m, uint64_t key)
{
	const size_t mask = m->nslots - 1;
	size_t i, j;

	for (i = mix64(key) & mask; m->slots[i].used; i = (i + 1) & mask) {
		if (m->slots[i].key == key)
			break;
	}
	if (!m->slots[i].used)
		return false;

	/* Backward-shift the cluster so probes never cross a hole. */
	for (j = (i + 1) & mask; m->slots[j].used; j = (j + 1) & mask) {
		const size_t home = mix64(m->slots[j].key) & mask;

		if (((j - home) & mask) >= ((j - i) & mask)) {
			m->slots[i] = m->slots[j];
			i = j;
		}
	}
	m->slots[i].used = 0;
	m->count--;
	return true;
}
