/*
 * Fuzzy term matching on the device (sm_100a, integer pipes).
 *
 * The reference resolves a query token that is not in the vocabulary by a
 * BK-tree search with Wagner-Fischer distances (ref src/index/idxterm.c:
 * 210-249, src/algo/bktree.c:219-275, src/algo/levdist.c:67-150): ~13-19 %
 * of a 1 M-term vocabulary visited per lookup, ~9 lookups/s on one core.
 *
 * Here ONE WARP PER QUERY TERM scans the whole vocabulary with Myers'
 * bit-parallel edit distance (Hyyro's global-distance form): the query is
 * the <= 64-bit pattern, each lane walks the bytes of a different vocabulary
 * term.  Terms are bucketed by length in fixed 16-byte (or 64-byte) slots,
 * so a lane's term arrives as one coalesced 16-byte load and only the
 * buckets with |len(term) - len(query)| <= 2 are scanned at all.
 *
 * Before Myers runs, a 64-bit SIGNATURE of each term (bit h(byte) set for
 * every byte it contains) is tested against the query's: an edit changes at
 * most one byte occurrence on either side, so d(q,t) <= 2 needs
 * popc(sig_q & ~sig_t) <= 2 and popc(sig_t & ~sig_q) <= 2 -- a necessary
 * condition, whatever h is, so no candidate is lost.  POPC issues on the
 * quarter-rate XU pipe, so the bulk test is ONE of them per pair: the
 * signatures folded to 32 bits (another valid h) must differ in at most 4
 * bits; the two-sided 64-bit test runs only for what passes.  The folded
 * signatures of a bucket are staged in shared memory once per CTA (a ring of
 * stages filled by a producer warp with bulk TMA copies) and tested by all
 * its warps (query terms, sorted by length on the host so that a CTA's share
 * buckets); the few survivors queue up per warp and go through Myers 32
 * at a time, one per lane.  Candidate sets, distances and the chosen term are
 * unchanged.
 *
 * The reference's answer is NOT "nearest" or "most popular" (SURVEY 8a F3):
 * its child range is half-open and its selection loop never updates the
 * running maximum, so it returns the first candidate in BFS order that the
 * pruned search reaches.  That is reproduced exactly from the scan: a term t
 * with d(q,t) <= 2 is a reference candidate iff every tree edge on the path
 * root -> t satisfies edge in [max(d(q,parent)-2,0), min(d(q,parent)+2,63)),
 * and the winner is the candidate of least BFS rank with a non-zero total.
 */
#ifndef NXSB_GPU_FUZZY_CUH
#define NXSB_GPU_FUZZY_CUH

#include <algorithm>
#include <vector>
#include <cstring>

#include "common.cuh"

#define FZ_TOLERANCE	2	/* ref index/index.h:26 */
#define FZ_EDGE_MAX	63	/* ref algo/bktree.h:11 */
#define FZ_WARPS	8	/* query terms per CTA, 64-bit patterns */
#define FZ_WARPS32	16	/* ... 32-bit patterns */
#define FZ_CHUNK	1024	/* signatures per ring stage */
#define FZ_STAGES	4	/* ring depth */
#define FZ_MAX_QLEN	64	/* pattern bits */
#define FZ_ROOT		0xffffffffu

/* Allocation counters of engine.cu (nxsb_alloc_events). */
static cudaError_t counted_malloc(void **p, size_t bytes);
static void counted_free(void *p);

struct FuzzyImage {
	bool		loaded = false;
	uint32_t	n_terms = 0;
	/* original order (term index = id - 1) */
	unsigned char *	d_blob = nullptr;
	uint32_t *	d_off = nullptr;	// [V + 1]
	unsigned char *	d_live = nullptr;	// total > 0
	uint32_t *	d_parent = nullptr;
	unsigned char *	d_edge = nullptr;
	uint32_t *	d_rank = nullptr;
	/* scan order: 16-byte slots bucketed by length 1..16 */
	uint4 *		d_slot16 = nullptr;
	uint32_t *	d_slot16_term = nullptr;
	unsigned long long *d_sig16 = nullptr;	// [n16] byte-presence signatures
	uint32_t *	d_sig32 = nullptr;	// [n16 + pad] the same folded to 32 bits
	uint32_t	n16 = 0;
	uint32_t	len16_start[18] = { 0 };	// bucket L = [start[L], start[L+1])
	/* terms longer than 16 bytes, any length (generic path) */
	uint32_t *	d_long_term = nullptr;
	uint32_t	n_long = 0;
	uint32_t	max_len = 0;
};

/* Signature bit of a byte: collision-free on [a-z0-9], any map is valid. */
__host__ __device__ __forceinline__ unsigned long long
fz_sig_bit(unsigned char b)
{
	return 1ull << ((((uint32_t)b * 225u) >> 5) & 63u);
}

static void
fuzzy_free(FuzzyImage &f)
{
	cudaFree(f.d_blob); cudaFree(f.d_off); cudaFree(f.d_live);
	cudaFree(f.d_parent); cudaFree(f.d_edge); cudaFree(f.d_rank);
	cudaFree(f.d_slot16); cudaFree(f.d_slot16_term); cudaFree(f.d_long_term);
	cudaFree(f.d_sig16); cudaFree(f.d_sig32);
	f = FuzzyImage();
}

static int
fuzzy_load(FuzzyImage &f, uint32_t V, const char *blob, const uint32_t *off,
    const uint64_t *total, const uint32_t *parent, const uint8_t *edge,
    const uint32_t *rank, cudaStream_t st)
{
	fuzzy_free(f);

	std::vector<unsigned char> live(V ? V : 1);
	std::vector<uint32_t> cnt(18, 0), longs;
	for (uint32_t t = 0; t < V; t++) {
		const uint32_t len = off[t + 1] - off[t];

		live[t] = total ? (total[t] > 0) : 1;
		f.max_len = len > f.max_len ? len : f.max_len;
		if (len >= 1 && len <= 16)
			cnt[len]++;
		else
			longs.push_back(t);
	}
	f.len16_start[0] = f.len16_start[1] = 0;
	for (int L = 1; L <= 16; L++)
		f.len16_start[L + 1] = f.len16_start[L] + cnt[L];
	f.n16 = f.len16_start[17];
	f.n_long = longs.size();

	std::vector<uint4> slots(f.n16 ? f.n16 : 1);
	std::vector<unsigned long long> sigs(f.n16 ? f.n16 : 1);
	std::vector<uint32_t> sigs32((size_t)f.n16 + 8, 0u);
	std::vector<uint32_t> slot_term(f.n16 ? f.n16 : 1), cur(f.len16_start, f.len16_start + 18);
	for (uint32_t t = 0; t < V; t++) {
		const uint32_t len = off[t + 1] - off[t];

		if (len < 1 || len > 16)
			continue;
		unsigned char b[16] = { 0 };
		memcpy(b, blob + off[t], len);
		memcpy(&slots[cur[len]], b, 16);
		unsigned long long sg = 0;
		for (uint32_t i = 0; i < len; i++)
			sg |= fz_sig_bit(b[i]);
		sigs[cur[len]] = sg;
		sigs32[cur[len]] = (uint32_t)(sg | (sg >> 32));
		slot_term[cur[len]] = t;
		cur[len]++;
	}

	const size_t blob_len = off[V];
	if (cudaMalloc(&f.d_blob, blob_len + 16) || cudaMalloc(&f.d_off, ((size_t)V + 1) * 4) ||
	    cudaMalloc(&f.d_live, V ? V : 1) || cudaMalloc(&f.d_parent, (size_t)(V ? V : 1) * 4) ||
	    cudaMalloc(&f.d_edge, V ? V : 1) || cudaMalloc(&f.d_rank, (size_t)(V ? V : 1) * 4) ||
	    cudaMalloc(&f.d_slot16, slots.size() * 16) ||
	    cudaMalloc(&f.d_slot16_term, slot_term.size() * 4) ||
	    cudaMalloc(&f.d_sig16, sigs.size() * 8) ||
	    cudaMalloc(&f.d_sig32, sigs32.size() * 4) ||
	    cudaMalloc(&f.d_long_term, (longs.size() ? longs.size() : 1) * 4))
		return -1;
	cudaMemcpyAsync(f.d_blob, blob, blob_len, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(f.d_off, off, ((size_t)V + 1) * 4, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(f.d_live, live.data(), V, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(f.d_parent, parent, (size_t)V * 4, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(f.d_edge, edge, V, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(f.d_rank, rank, (size_t)V * 4, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(f.d_slot16, slots.data(), (size_t)f.n16 * 16, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(f.d_slot16_term, slot_term.data(), (size_t)f.n16 * 4, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(f.d_sig16, sigs.data(), (size_t)f.n16 * 8, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(f.d_sig32, sigs32.data(), sigs32.size() * 4, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(f.d_long_term, longs.data(), longs.size() * 4, cudaMemcpyHostToDevice, st);
	if (cudaStreamSynchronize(st) != cudaSuccess)
		return -1;
	f.n_terms = V;
	f.loaded = true;
	return 0;
}

/*
 * The per-term "total > 0" flags again (ref idxterm.c:239 reads the totals
 * from the mapped file at search time: they move with every add / remove,
 * also of documents that bring no new term).
 */
static int
fuzzy_update_live(FuzzyImage &f, uint32_t V, const uint64_t *total, cudaStream_t st)
{
	if (!f.loaded || V != f.n_terms)
		return -1;
	std::vector<unsigned char> live(V ? V : 1);

	for (uint32_t t = 0; t < V; t++)
		live[t] = total[t] > 0;
	if (cudaMemcpyAsync(f.d_live, live.data(), V, cudaMemcpyHostToDevice, st) != cudaSuccess ||
	    cudaStreamSynchronize(st) != cudaSuccess)
		return -1;
	return 0;
}

/*
 * Myers / Hyyro bit-parallel Levenshtein distance: pattern (the query) in
 * the low m bits of a W-bit word, one column update per text byte.
 * peq[c] = bitmask of pattern positions holding byte c.
 */
template <typename W>
struct MyersState {
	W pv, mv;
	int score;
	W hb;

	__device__ __forceinline__ void init(int m)
	{
		pv = m >= (int)(8 * sizeof(W)) ? ~(W)0 : (((W)1 << m) - 1);
		mv = 0;
		score = m;
		hb = (W)1 << (m - 1);
	}
	__device__ __forceinline__ void step(W eq)
	{
		const W xv = eq | mv;
		const W xh = (((eq & pv) + pv) ^ pv) | eq;
		W ph = mv | ~(xh | pv);
		W mh = pv & xh;

		score += (ph & hb) ? 1 : ((mh & hb) ? -1 : 0);
		ph = (ph << 1) | 1;
		mh <<= 1;
		pv = mh | ~(xv | ph);
		mv = ph & xv;
	}
};

/* Distance from the query to an arbitrary term read byte-wise from the blob. */
template <typename W>
__device__ int
myers_blob(const W *peq, int m, const unsigned char *s, uint32_t len)
{
	MyersState<W> st;

	st.init(m);
	for (uint32_t i = 0; i < len; i++)
		st.step(peq[s[i]]);
	return st.score;
}

/*
 * Would the reference's pruned BFS reach term t?  Walk t's ancestors and
 * test each edge against the exclusive-upper-bound child range of
 * ref bktree.c:151-157,258-264.
 */
template <typename W>
__device__ bool
bk_reachable(const FuzzyImage &f, const W *peq, int m, uint32_t t)
{
	uint32_t c = t;

	for (;;) {
		const uint32_t p = f.d_parent[c];
		if (p == FZ_ROOT)
			return true;
		const uint32_t s = f.d_off[p];
		const int d = myers_blob<W>(peq, m, f.d_blob + s, f.d_off[p + 1] - s);
		const int lo = d - FZ_TOLERANCE < 0 ? 0 : d - FZ_TOLERANCE;
		const int hi = d + FZ_TOLERANCE > FZ_EDGE_MAX ? FZ_EDGE_MAX : d + FZ_TOLERANCE;
		const int e = f.d_edge[c];

		if (e < lo || e >= hi)
			return false;
		c = p;
	}
}

struct FuzzyBest {
	uint32_t rank, term, dist;
};

/*
 * Optional candidate lists (nxsb_engine_fuzzy_candidates): every vocabulary
 * term within the tolerance is recorded, with whether the reference's pruned
 * BK-tree walk would visit it (ref bktree.c:252-254 pushes exactly those) and
 * whether its total is non-zero.  list[q * cap + i] = { BFS rank, term index,
 * distance | reached << 8 | live << 9, 0 }; cnt[q] counts past cap.
 */
struct FuzzyCandOut {
	uint32_t *	cnt = nullptr;
	uint4 *		list = nullptr;
	uint32_t	cap = 0;
};

template <typename W>
__device__ __forceinline__ void
fuzzy_consider(const FuzzyImage &f, const W *peq, int m, uint32_t t, int d,
    FuzzyBest &best, uint32_t &n_true, const FuzzyCandOut &co, uint32_t qi)
{
	if (d > FZ_TOLERANCE)
		return;
	n_true++;
	if (co.cnt) {
		/* The diagnostic path: every match, reached or not. */
		const bool reached = bk_reachable<W>(f, peq, m, t);
		const uint32_t at = atomicAdd(co.cnt + qi, 1u);

		if (at < co.cap)
			co.list[(size_t)qi * co.cap + at] = make_uint4(f.d_rank[t], t,
			    (uint32_t)d | (reached ? 0x100u : 0u) | (f.d_live[t] ? 0x200u : 0u), 0u);
		if (reached && f.d_live[t] && f.d_rank[t] < best.rank) {
			best.rank = f.d_rank[t];
			best.term = t;
			best.dist = d;
		}
		return;
	}
	if (!f.d_live[t])		/* idxterm.c:239: total must be > 0 */
		return;
	const uint32_t r = f.d_rank[t];
	if (r < best.rank && bk_reachable<W>(f, peq, m, t)) {
		best.rank = r;
		best.term = t;
		best.dist = d;
	}
}

/*
 * One warp per query term.  W = uint32_t for queries of <= 32 bytes,
 * unsigned long long for <= 64; NW query terms per CTA.
 */
template <typename W, int NW>
__global__ void __launch_bounds__((NW + 1) * 32)
fuzzy_scan_kernel(const FuzzyImage f, const unsigned char *__restrict__ qblob,
    const uint32_t *__restrict__ qoff, const uint32_t *__restrict__ qsel,
    uint32_t n_sel, uint32_t *__restrict__ out_term,
    uint32_t *__restrict__ out_dist, uint32_t *__restrict__ out_true,
    const FuzzyCandOut co, unsigned long long *__restrict__ out_key)
{
	/*
	 * A call with few query terms would leave most SMs idle (one warp per
	 * term, and a term's scan is ~4000 rounds long): gridDim.y CTAs then
	 * share a group of terms, CTA y taking every gridDim.y-th chunk of the
	 * signatures, and the answers meet in out_key by atomicMin over
	 * (BFS rank, distance, term) -- see fuzzy_unpack_kernel.
	 */
	const uint32_t part = blockIdx.y, n_parts = gridDim.y;
	__shared__ W s_peq[NW][256];
	__shared__ __align__(128) uint32_t s_sig[FZ_STAGES][FZ_CHUNK];
	__shared__ __align__(8) unsigned long long s_bar[2 * FZ_STAGES];	/* full[], empty[] */
	__shared__ uint32_t s_queue[NW][64];	/* slots that passed the 64-bit signature test */
	__shared__ uint32_t s_first[NW][160];	/* ... the folded one (up to 128 join per round) */
	__shared__ int s_lo, s_hi;

	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t wi = blockIdx.x * NW + warp;
	const bool producer = warp == NW;	/* the extra warp feeds the ring */
	const bool active = !producer && wi < n_sel;
	const uint32_t full0 = smem_addr(s_bar), empty0 = smem_addr(s_bar + FZ_STAGES);
	W *peq = s_peq[producer ? 0 : warp];
	uint32_t *queue = s_queue[producer ? 0 : warp], *first = s_first[producer ? 0 : warp];
	uint32_t qi = 0;
	int m = 0, l_lo = 1, l_hi = 0;
	unsigned long long qsig = 0;

	if (threadIdx.x == 0) {
		s_lo = 17;
		s_hi = 0;
		for (uint32_t st = 0; st < FZ_STAGES; st++) {
			mbar_init(full0 + 8 * st, 1);
			mbar_init(empty0 + 8 * st, NW);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();
	if (active) {
		qi = qsel[wi];
		const unsigned char *q = qblob + qoff[qi];

		m = (int)(qoff[qi + 1] - qoff[qi]);
		for (int c = lane; c < 256; c += 32)
			peq[c] = 0;
		__syncwarp();
		for (int i = lane; i < m; i += 32) {
			/* Distinct lanes may share a byte value: OR atomically. */
			if (sizeof(W) == 4)
				atomicOr(reinterpret_cast<unsigned int *>(&peq[q[i]]), 1u << i);
			else
				atomicOr(reinterpret_cast<unsigned long long *>(&peq[q[i]]), 1ull << i);
			qsig |= fz_sig_bit(q[i]);
		}
		for (int o = 16; o; o >>= 1)
			qsig |= __shfl_xor_sync(0xffffffffu, qsig, o);
		/* 16-byte slots, only the length buckets within the tolerance. */
		l_lo = m - FZ_TOLERANCE < 1 ? 1 : m - FZ_TOLERANCE;
		l_hi = m + FZ_TOLERANCE > 16 ? 16 : m + FZ_TOLERANCE;
		if (lane == 0 && l_lo <= l_hi) {
			atomicMin(&s_lo, l_lo);
			atomicMax(&s_hi, l_hi);
		}
	}
	__syncthreads();
	const int cta_lo = s_lo, cta_hi = s_hi;

	FuzzyBest best = { 0xffffffffu, 0xffffffffu, 0 };
	uint32_t n_true = 0, nq = 0;

	/* Myers over the first cnt queued survivors, one per lane. */
	auto drain = [&](uint32_t cnt) {
		if (lane < cnt) {
			const uint32_t s = queue[lane];
			int L = l_lo;
			const uint4 w = __ldg(f.d_slot16 + s);

			while (s >= f.len16_start[L + 1])	/* the slot's length bucket */
				L++;
			const uint32_t ww[4] = { w.x, w.y, w.z, w.w };
			MyersState<W> st;

			st.init(m);
#pragma unroll
			for (int j = 0; j < 16; j++) {
				if (j < L)
					st.step(peq[(ww[j >> 2] >> ((j & 3) * 8)) & 0xffu]);
			}
			if (st.score <= FZ_TOLERANCE)
				fuzzy_consider<W>(f, peq, m, __ldg(f.d_slot16_term + s),
				    st.score, best, n_true, co, qi);
		}
		__syncwarp();
	};

	/*
	 * The folded signatures of the CTA's buckets go through a ring of
	 * FZ_STAGES shared-memory stages, FZ_CHUNK at a time: the extra warp
	 * issues one 1-D bulk TMA copy per chunk (contiguous in d_sig32) as soon
	 * as every consumer warp has released the stage, the consumers wait on
	 * the stage's "full" mbarrier -- so a warp that queues and verifies many
	 * survivors may fall a few chunks behind the others without stopping
	 * them.  A warp tests 128 signatures per round (one 16-byte load of four
	 * per lane).  Buckets ascend by length, so the slots of THIS query's
	 * buckets are one range [my0, my1).
	 */
	/* Chunks start at multiples of 4 slots: 16-byte aligned copies of d_sig32 (padded). */
	const uint32_t g0 = cta_lo <= cta_hi ? f.len16_start[cta_lo] & ~3u : 0u;
	const uint32_t g1 = cta_lo <= cta_hi ? f.len16_start[cta_hi + 1] : 0u;
	const uint32_t my0 = l_lo <= l_hi ? f.len16_start[l_lo] : 0u;
	const uint32_t my1 = l_lo <= l_hi ? f.len16_start[l_hi + 1] : 0u;
	const uint32_t q32 = (uint32_t)(qsig | (qsig >> 32));

	if (producer) {
		if (lane == 0) {
			uint32_t ps = 0, pph = 1;

			for (uint32_t c0 = g0 + part * FZ_CHUNK; c0 < g1; c0 += n_parts * FZ_CHUNK) {
				const uint32_t n = g1 - c0 < FZ_CHUNK ? g1 - c0 : FZ_CHUNK;
				const uint32_t bytes = ((n + 3u) & ~3u) * 4u;

				mbar_wait(empty0 + 8 * ps, pph);
				mbar_arrive_expect_tx(full0 + 8 * ps, bytes);
				tma_load_1d(smem_addr(s_sig[ps]), f.d_sig32 + c0, bytes, full0 + 8 * ps);
				if (++ps == FZ_STAGES) {
					ps = 0;
					pph ^= 1u;
				}
			}
		}
		return;
	}
	/*
	 * The two-sided test on the full signatures of 32 queued slots (the last
	 * 32), one per lane: 32 independent loads in flight instead of a
	 * divergent one per hit.  What passes joins the Myers queue.
	 */
	uint32_t nf = 0;
	auto sift = [&]() {
		const uint32_t cnt = nf < 32 ? nf : 32;
		const uint32_t at = nf - cnt;
		bool pass = false;
		uint32_t slot = 0;

		if (lane < cnt) {
			slot = first[at + lane];
			const unsigned long long tl = __ldg(f.d_sig16 + slot);

			pass = __popcll(qsig & ~tl) <= FZ_TOLERANCE &&
			    __popcll(tl & ~qsig) <= FZ_TOLERANCE;
		}
		const uint32_t pm = __ballot_sync(0xffffffffu, pass);

		nf = at;
		if (pm) {
			if (pass)
				queue[nq + __popc(pm & ((1u << lane) - 1u))] = slot;
			nq += __popc(pm);
			__syncwarp();
			if (nq >= 32) {
				drain(32);
				/* The leftovers (< 32) move to the front. */
				const uint32_t left = nq - 32;
				const uint32_t e = lane < left ? queue[32 + lane] : 0u;

				__syncwarp();
				if (lane < left)
					queue[lane] = e;
				nq = left;
			}
		}
		__syncwarp();
	};
	uint32_t cs = 0, cph = 0;
	for (uint32_t c0 = g0 + part * FZ_CHUNK; c0 < g1; c0 += n_parts * FZ_CHUNK) {
		const uint32_t n = g1 - c0 < FZ_CHUNK ? g1 - c0 : FZ_CHUNK;

		mbar_wait(full0 + 8 * cs, cph);
		if (active && c0 < my1 && c0 + n > my0) {
			const uint4 *sg4 = reinterpret_cast<const uint4 *>(s_sig[cs]);

			for (uint32_t base = 0; base < n; base += 128) {
				const uint4 t4 = sg4[(base >> 2) + lane];
				const uint32_t ts[4] = { t4.x, t4.y, t4.z, t4.w };
				const uint32_t slot0 = c0 + base + 4u * lane;
				uint32_t hit = 0;

#pragma unroll
				for (int u = 0; u < 4; u++)
					hit |= (__popc(q32 ^ ts[u]) <= 2 * FZ_TOLERANCE &&
					    slot0 + u >= my0 && slot0 + u < my1) ? 1u << u : 0u;
				if (!__any_sync(0xffffffffu, hit != 0))
					continue;
				/* Queue what passed (slot order is irrelevant to the answer). */
#pragma unroll
				for (int u = 0; u < 4; u++) {
					const bool h = (hit >> u) & 1u;
					const uint32_t hm = __ballot_sync(0xffffffffu, h);

					if (h)
						first[nf + __popc(hm & ((1u << lane) - 1u))] = slot0 + u;
					nf += __popc(hm);
				}
				__syncwarp();
				while (nf >= 32)
					sift();
			}
		}
		/* Done with the stage (idle warps included: the count is NW). */
		__syncwarp();
		if (lane == 0)
			mbar_arrive(empty0 + 8 * cs);
		if (++cs == FZ_STAGES) {
			cs = 0;
			cph ^= 1u;
		}
	}
	if (!active)
		return;
	while (nf)
		sift();
	drain(nq);

	/* Terms longer than 16 bytes: byte-wise from the blob. */
	for (uint32_t i = lane + 32 * part; i < f.n_long; i += 32 * n_parts) {
		const uint32_t t = f.d_long_term[i];
		const uint32_t s = f.d_off[t], len = f.d_off[t + 1] - s;
		const int diff = (int)len - m;

		if (diff > FZ_TOLERANCE || diff < -FZ_TOLERANCE)
			continue;
		const int d = myers_blob<W>(peq, m, f.d_blob + s, len);
		fuzzy_consider<W>(f, peq, m, t, d, best, n_true, co, qi);
	}

	/* Warp argmin over BFS rank. */
	for (int o = 16; o; o >>= 1) {
		const uint32_t r = __shfl_xor_sync(0xffffffffu, best.rank, o);
		const uint32_t t = __shfl_xor_sync(0xffffffffu, best.term, o);
		const uint32_t d = __shfl_xor_sync(0xffffffffu, best.dist, o);

		if (r < best.rank) {
			best.rank = r;
			best.term = t;
			best.dist = d;
		}
		n_true += __shfl_xor_sync(0xffffffffu, n_true, o);
	}
	if (lane == 0 && out_key) {
		if (best.term != 0xffffffffu)
			atomicMin(out_key + qi, ((unsigned long long)best.rank << 32) |
			    ((unsigned long long)best.dist << 30) | best.term);
		if (out_true && n_true)
			atomicAdd(out_true + qi, n_true);
	} else if (lane == 0) {
		out_term[qi] = best.term == 0xffffffffu ? 0 : best.term + 1;
		out_dist[qi] = best.dist;
		if (out_true)
			out_true[qi] = n_true;
	}
}

/* out_key[i] = min over the parts of (rank << 32 | dist << 30 | term), ~0 = no match. */
__global__ void __launch_bounds__(256)
fuzzy_unpack_kernel(const unsigned long long *__restrict__ key, uint32_t n,
    uint32_t *__restrict__ out_term, uint32_t *__restrict__ out_dist)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

	if (i >= n)
		return;
	const unsigned long long k = key[i];

	out_term[i] = k == ~0ull ? 0u : (uint32_t)(k & 0x3fffffffu) + 1u;
	out_dist[i] = k == ~0ull ? 0u : (uint32_t)(k >> 30) & 3u;
}

/*
 * Device buffers of the lookups, grown on demand and kept: a stream of
 * batches allocates nothing (cudaFree synchronises the whole device).
 */
struct FuzzyScratch {
	unsigned char *	d_qblob = nullptr;
	uint32_t *	d_words = nullptr;	/* [qoff | sel32 | sel64 | term | dist | true | cand cnt] */
	uint4 *		d_cand = nullptr;
	size_t		blob_cap = 0, words_cap = 0, cand_cap = 0;
	std::vector<uint32_t> sel32, sel64;
};

static void
fuzzy_scratch_free(FuzzyScratch &z)
{
	counted_free(z.d_qblob);
	counted_free(z.d_words);
	counted_free(z.d_cand);
	z = FuzzyScratch();
}

template <typename T>
static bool
fuzzy_grow(T *&p, size_t &cap, size_t want)
{
	if (want <= cap)
		return true;
	counted_free(p);
	p = nullptr;
	cap = 0;
	const size_t n = want + want / 2 + 1024;

	if (counted_malloc(reinterpret_cast<void **>(&p), n * sizeof(T)) != cudaSuccess) {
		p = nullptr;
		return false;
	}
	cap = n;
	return true;
}

/*
 * out_*: host arrays [n].  cand_cap > 0: also the candidate lists, host arrays
 * cand_cnt[n] and cand[n * cand_cap] (see FuzzyCandOut), unsorted.
 */
/* Where fuzzy_enqueue left its results (device memory of the scratch). */
struct FuzzyOut {
	uint32_t *	term = nullptr;		/* [n] 1-based term id, 0 = none */
	uint32_t *	dist = nullptr;
	uint32_t *	n_true = nullptr;
	uint32_t *	cand_cnt = nullptr;
	uint32_t *	extra = nullptr;	/* [extra_words] for the caller */
};

/*
 * Everything of a lookup call that can be queued: copies in, kernels, results
 * left on the device (n >= 1).  No host synchronisation.
 */
static int
fuzzy_enqueue(FuzzyImage &f, FuzzyScratch &z, uint32_t n, const char *qblob, const uint32_t *qoff,
    uint32_t cand_cap, size_t extra_words, cudaStream_t st, int n_sms, int *launches, FuzzyOut *out)
{
	std::vector<uint32_t> &sel32 = z.sel32, &sel64 = z.sel64;

	sel32.clear();
	sel64.clear();
	for (uint32_t i = 0; i < n; i++) {
		const uint32_t m = qoff[i + 1] - qoff[i];

		/*
		 * Empty queries never reach the fuzzy search; patterns beyond
		 * 64 bytes are reported as "no match" (documented limit).
		 */
		if (m >= 1 && m <= 32)
			sel32.push_back(i);
		else if (m >= 33 && m <= FZ_MAX_QLEN)
			sel64.push_back(i);
	}
	/* Query terms of like length share a CTA, hence its staged buckets. */
	auto by_len = [&](uint32_t a, uint32_t b) {
		return qoff[a + 1] - qoff[a] < qoff[b + 1] - qoff[b];
	};
	std::stable_sort(sel32.begin(), sel32.end(), by_len);
	std::stable_sort(sel64.begin(), sel64.end(), by_len);

	const size_t nn = n;
	const size_t o_qoff = 0, o_s32 = o_qoff + nn + 1, o_s64 = o_s32 + nn, o_term = o_s64 + nn,
	    o_dist = o_term + nn, o_true = o_dist + nn, o_cnt = o_true + nn, o_key = (o_cnt + nn + 1) & ~(size_t)1,
	    o_extra = o_key + 2 * nn, words = o_extra + extra_words;

	if (!fuzzy_grow(z.d_qblob, z.blob_cap, (size_t)qoff[n] + 16) ||
	    !fuzzy_grow(z.d_words, z.words_cap, words) ||
	    (cand_cap && !fuzzy_grow(z.d_cand, z.cand_cap, nn * cand_cap)))
		return -1;
	uint32_t *w = z.d_words;
	FuzzyCandOut co;

	if (cand_cap) {
		co.cnt = w + o_cnt;
		co.list = z.d_cand;
		co.cap = cand_cap;
	}
	cudaMemcpyAsync(z.d_qblob, qblob, qoff[n], cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(w + o_qoff, qoff, (nn + 1) * 4, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(w + o_s32, sel32.data(), sel32.size() * 4, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(w + o_s64, sel64.data(), sel64.size() * 4, cudaMemcpyHostToDevice, st);
	cudaMemsetAsync(w + o_term, 0, 4 * nn * 4, st);		/* term, dist, true, cand cnt */
	/* Few terms: several CTAs per group of terms, so that the whole GPU scans. */
	const uint32_t ctas32 = (uint32_t)((sel32.size() + FZ_WARPS32 - 1) / FZ_WARPS32);
	const uint32_t ctas64 = (uint32_t)((sel64.size() + FZ_WARPS - 1) / FZ_WARPS);
	const uint32_t target = (uint32_t)n_sms * 4u;
	uint32_t parts = ctas32 + ctas64 ? std::min(64u, std::max(1u, target / (ctas32 + ctas64))) : 1u;
	if (f.n_terms >= (1u << 30))
		parts = 1;		/* the packed key holds 30 bits of term */
	unsigned long long *keys = parts > 1 ? reinterpret_cast<unsigned long long *>(w + o_key) : nullptr;

	if (keys)
		cudaMemsetAsync(keys, 0xff, nn * 8, st);
	if (!sel32.empty()) {
		fuzzy_scan_kernel<uint32_t, FZ_WARPS32><<<dim3(ctas32, parts),
		    (FZ_WARPS32 + 1) * 32, 0, st>>>(f, z.d_qblob, w + o_qoff, w + o_s32,
		    sel32.size(), w + o_term, w + o_dist, w + o_true, co, keys);
		(*launches)++;
	}
	if (!sel64.empty()) {
		fuzzy_scan_kernel<unsigned long long, FZ_WARPS><<<dim3(ctas64, parts),
		    (FZ_WARPS + 1) * 32, 0, st>>>(f, z.d_qblob, w + o_qoff, w + o_s64,
		    sel64.size(), w + o_term, w + o_dist, w + o_true, co, keys);
		(*launches)++;
	}
	if (keys) {
		fuzzy_unpack_kernel<<<(n + 255) / 256, 256, 0, st>>>(keys, n, w + o_term, w + o_dist);
		(*launches)++;
	}
	out->term = w + o_term;
	out->dist = w + o_dist;
	out->n_true = w + o_true;
	out->cand_cnt = w + o_cnt;
	out->extra = w + o_extra;
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

static int
fuzzy_run(FuzzyImage &f, FuzzyScratch &z, uint32_t n, const char *qblob, const uint32_t *qoff,
    uint32_t *out_term, uint32_t *out_dist, uint32_t *out_true,
    uint32_t cand_cap, uint32_t *cand_cnt, uint4 *cand,
    cudaStream_t st, int n_sms, int *launches)
{
	const size_t nn = n;
	FuzzyOut o;

	if (n == 0)
		return 0;
	if (fuzzy_enqueue(f, z, n, qblob, qoff, cand_cap, 0, st, n_sms, launches, &o) != 0)
		return -1;
	cudaMemcpyAsync(out_term, o.term, nn * 4, cudaMemcpyDeviceToHost, st);
	cudaMemcpyAsync(out_dist, o.dist, nn * 4, cudaMemcpyDeviceToHost, st);
	if (out_true)
		cudaMemcpyAsync(out_true, o.n_true, nn * 4, cudaMemcpyDeviceToHost, st);
	if (cand_cap) {
		cudaMemcpyAsync(cand_cnt, o.cand_cnt, nn * 4, cudaMemcpyDeviceToHost, st);
		cudaMemcpyAsync(cand, z.d_cand, nn * cand_cap * sizeof(uint4), cudaMemcpyDeviceToHost, st);
	}
	return cudaStreamSynchronize(st) == cudaSuccess ? 0 : -1;
}

/* tokens[pos[i]] = term[i]: the lookups' answers take their places in a batch's token list. */
__global__ void __launch_bounds__(256)
fuzzy_patch_tokens_kernel(uint32_t *__restrict__ tokens, const uint32_t *__restrict__ pos,
    const uint32_t *__restrict__ term, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

	if (i < n)
		tokens[pos[i]] = term[i];
}

#endif
