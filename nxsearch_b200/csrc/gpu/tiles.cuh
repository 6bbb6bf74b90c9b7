/*
 * score_tiles_kernel: the hot kernel (see the header comment of score.cuh).
 *
 * Round-1 profile (profiles/r1_*): the first version of this kernel was
 * instruction-issue bound -- ~155 thread-instructions per posting, most of
 * them in two full 64-bit-key passes over the accumulator and an IEEE
 * division -- with the L2 hit rate at 91 % and DRAM almost idle.  This
 * version is built around the instruction budget instead:
 *
 *   - postings arrive two at a time in one 16-byte load;
 *   - log(tf + 1) and the BM25 length term K0 + K1 * dl come from small
 *     shared-memory tables (one test on the packed word selects the rare
 *     slow path), the division is a reciprocal + multiply;
 *   - the top-k scan reads the accumulator as float4 and compares raw floats
 *     with the query's threshold; the few survivors are compacted into a
 *     512-entry shared buffer and sorted there.  The exact 64-bit radix
 *     select over the whole tile only runs when that buffer overflows (the
 *     first tiles of a query, before a threshold exists).
 */
#ifndef NXSB_GPU_TILES_CUH
#define NXSB_GPU_TILES_CUH

#define CAND_SMEM	512u	/* compacted candidates per tile */
#define DENTAB_N	512	/* K0 + K1 * dl for dl < 512 */

/* (float)log(tf + 1), ref ranking.c:90,168; table for tf < 256. */
__device__ __forceinline__ float
log_tf(uint32_t tf, const float *s_logtab)
{
	return tf < LOGTAB_N ? s_logtab[tf] : (float)log((double)tf + 1.0);
}

/*
 * One posting's score.  TF-IDF is bit-exact with ref ranking.c:90-96 (float
 * tf times float idf, correctly rounded).  BM25 (ranking.c:168-175) runs in
 * fp32 from double-precision host constants with a reciprocal instead of a
 * division: measured error < 1e-6 relative to the reference's fp64
 * evaluation (budget 1e-5).
 */
template <bool WIDE, int ALGO>
__device__ __forceinline__ float
score_posting(const ScoreParams &p, const float *s_logtab, const float *s_dentab,
    uint32_t doc, uint32_t w, float idf)
{
	if (ALGO == NXSB_ALGO_TFIDF) {
		const uint32_t tf = WIDE ? w : (w & 0xffffu);
		return __fmul_rn(log_tf(tf, s_logtab), idf);
	}
	float x, d;
	if (WIDE) {
		x = log_tf(w, s_logtab);
		d = __fmaf_rn(p.K1, (float)(int)__ldg(p.doc_len + doc), p.K0);
	} else if ((w & 0xfe00ff00u) == 0) {
		/* tf < 256 and dl < 512: both from shared-memory tables. */
		x = s_logtab[w & 0xffu];
		d = s_dentab[w >> 16];
	} else {
		x = log_tf(w & 0xffffu, s_logtab);
		d = __fmaf_rn(p.K1, (float)(int)(w >> 16), p.K0);
	}
	return __fmul_rn(__fdividef(x, __fadd_rn(x, d)), idf);
}

__device__ __forceinline__ uint32_t
block_sum_u32(uint32_t v, uint32_t *s_scratch)
{
	/* s_scratch: one word, zeroed by the caller before a barrier. */
	for (int o = 16; o; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0 && v)
		atomicAdd(s_scratch, v);
	__syncthreads();
	return *s_scratch;
}

/*
 * Selection key inside a tile (radix fallback): 46 bits, score bits above
 * the 14-bit local document index.
 */
__device__ __forceinline__ unsigned long long
sel_key(float v, uint32_t local)
{
	return ((unsigned long long)__float_as_uint(v) << TILE_SHIFT) | local;
}

template <bool LOGIC, bool WIDE, int ALGO>
__global__ void __launch_bounds__(TILE_THREADS, LOGIC ? 2 : 3)
score_tiles_kernel(const ScoreParams p)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	float *acc = reinterpret_cast<float *>(smem_raw);		// [TILE_DOCS]
	uint32_t *bits = reinterpret_cast<uint32_t *>(acc + TILE_DOCS);	// LOGIC
	uint32_t *mask = bits + (LOGIC ? p.max_tokens * TILE_WORDS : 0);	// LOGIC

	__shared__ float s_logtab[LOGTAB_N];
	__shared__ float s_dentab[DENTAB_N];
	__shared__ __align__(16) unsigned long long s_cand[CAND_SMEM];
	__shared__ uint32_t s_lo[NXSB_MAX_QUERY_TOKENS], s_hi[NXSB_MAX_QUERY_TOKENS];
	__shared__ uint32_t s_hist[256];
	__shared__ uint32_t s_item, s_any, s_cnt, s_emit, s_base, s_want, s_ncand;
	__shared__ unsigned long long s_prefix, s_theta;

	const uint32_t tid = threadIdx.x;

	for (uint32_t i = tid; i < LOGTAB_N; i += TILE_THREADS)
		s_logtab[i] = p.logtab[i];
	for (uint32_t i = tid; i < DENTAB_N; i += TILE_THREADS)
		s_dentab[i] = __fmaf_rn(p.K1, (float)(int)i, p.K0);
	{
		/* Invariant: the accumulator is all-zero between work items. */
		float4 *a4 = reinterpret_cast<float4 *>(acc);
		for (uint32_t i = tid; i < TILE_DOCS / 4; i += TILE_THREADS)
			a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
	}

	const unsigned long long n_items = (unsigned long long)p.n_q * p.ntiles;

	for (;;) {
		__syncthreads();
		if (tid == 0) {
			s_item = atomicAdd(p.work_counter, 1u);
			s_any = 0;
			s_cnt = 0;
			s_emit = 0;
			s_ncand = 0;
		}
		__syncthreads();
		const unsigned long long item = s_item;
		if (item >= n_items)
			break;

		/* Tile-major, highest tile first (ties prefer higher ids). */
		const uint32_t tile = p.ntiles - 1 - (uint32_t)(item / p.n_q);
		const uint32_t slot = (uint32_t)(item % p.n_q);
		const QDesc qd = p.queries[p.qlist[slot]];
		const uint32_t ntok = qd.n_tokens;
		const uint32_t tile_lo = tile << TILE_SHIFT;

		if (tid < ntok) {
			const DTok &t = p.toks[qd.tok_off + tid];
			const uint32_t lo = __ldg(t.skip + tile);
			const uint32_t hi = __ldg(t.skip + tile + 1);

			s_lo[tid] = lo;
			s_hi[tid] = hi;
			if (hi > lo)
				s_any = 1;
		}
		/* One read of the threshold per item, shared by all threads. */
		if (tid == 32)
			s_theta = *(volatile unsigned long long *)(p.thr + slot);
		if (LOGIC) {
			for (uint32_t i = tid; i < ntok * TILE_WORDS; i += TILE_THREADS)
				bits[i] = 0;
		}
		__syncthreads();
		if (!s_any)
			continue;

		/*
		 * Stream each token's slice; token-list order = summation order
		 * (ref results.c:135-137).  Within one list a document occurs at
		 * most once, so the read-modify-write needs no atomics.
		 */
		for (uint32_t j = 0; j < ntok; j++) {
			const DTok &t = p.toks[qd.tok_off + j];
			const float idf = t.idf;
			unsigned long long g0 = t.post_off + s_lo[j];
			unsigned long long g1 = t.post_off + s_hi[j];

			auto accumulate = [&](uint32_t doc, uint32_t w) {
				const uint32_t local = doc - tile_lo;

				acc[local] += score_posting<WIDE, ALGO>(p, s_logtab,
				    s_dentab, doc, w, idf);
				if (LOGIC)
					atomicOr(&bits[j * TILE_WORDS + (local >> 5)],
					    1u << (local & 31));
			};

			/* Peel to 16-byte alignment; the rest goes two per load. */
			if ((g0 & 1) && g0 < g1) {
				if (tid == 0) {
					const uint2 v = __ldg(p.post + g0);
					accumulate(v.x, v.y);
				}
				g0++;
			}
			if ((g1 & 1) && g0 < g1) {
				if (tid == 32) {
					const uint2 v = __ldg(p.post + g1 - 1);
					accumulate(v.x, v.y);
				}
				g1--;
			}
			const uint4 *p4 = reinterpret_cast<const uint4 *>(p.post + g0);
			const uint32_t n4 = (uint32_t)((g1 - g0) >> 1);
			uint32_t i = tid;

			for (; i + TILE_THREADS < n4; i += 2 * TILE_THREADS) {
				const uint4 a = __ldg(p4 + i);
				const uint4 b = __ldg(p4 + i + TILE_THREADS);

				accumulate(a.x, a.y);
				accumulate(a.z, a.w);
				accumulate(b.x, b.y);
				accumulate(b.z, b.w);
			}
			if (i < n4) {
				const uint4 a = __ldg(p4 + i);

				accumulate(a.x, a.y);
				accumulate(a.z, a.w);
			}
			__syncthreads();
		}

		/*
		 * Boolean logic (ref get_expr_bitmap, search.c:118-174): each
		 * thread evaluates the postfix program on one 32-document word
		 * of every token bitmap.
		 */
		if (LOGIC) {
			for (uint32_t w = tid; w < TILE_WORDS; w += TILE_THREADS) {
				uint32_t st[NXSB_MAX_QUERY_TOKENS + 1];
				int sp = 0;

				for (uint32_t c = 0; c < qd.n_prog; c++) {
					const int32_t op = p.prog[qd.prog_off + c];

					if (op >= 0) {
						st[sp++] = bits[op * TILE_WORDS + w];
					} else if (op == NXSB_OP_EMPTY) {
						st[sp++] = 0;
					} else {
						const uint32_t b = st[--sp];
						const uint32_t a = st[sp - 1];

						st[sp - 1] = op == NXSB_OP_AND ? (a & b) :
						    op == NXSB_OP_OR ? (a | b) : (a & ~b);
					}
				}
				mask[w] = sp ? st[sp - 1] : 0;
			}
			__syncthreads();
		}

		/*
		 * Top-k of the tile.  A document competes only if its key
		 * (score, id) beats the query's published threshold -- the
		 * k-th best key of some already finished tile, hence a lower
		 * bound of the final k-th best.
		 */
		const unsigned long long theta = s_theta;
		const uint32_t theta_bits = (uint32_t)(theta >> 32);
		const uint32_t theta_doc = (uint32_t)theta;
		/* Fast filter: v >= ths (the smallest denormal when unset). */
		const float ths = theta ? __uint_as_float(theta_bits) : __uint_as_float(1u);

		/* Exact test: valid document whose key is above the threshold. */
		auto candidate = [&](uint32_t i, float v) -> bool {
			if (LOGIC ? !((mask[i >> 5] >> (i & 31)) & 1u) : !(v > 0.f))
				return false;
			const uint32_t vb = __float_as_uint(v);
			return vb > theta_bits || (vb == theta_bits &&
			    (theta == 0 || tile_lo + i > theta_doc));
		};
		auto push = [&](uint32_t i, float v) {
			if (v >= ths && candidate(i, v)) {
				const uint32_t s = atomicAdd(&s_ncand, 1u);
				if (s < CAND_SMEM)
					s_cand[s] = make_key(v, tile_lo + i);
			}
		};

		{
			const float4 *a4 = reinterpret_cast<const float4 *>(acc);
#pragma unroll 4
			for (uint32_t i4 = tid; i4 < TILE_DOCS / 4; i4 += TILE_THREADS) {
				const float4 v = a4[i4];

				if (v.x >= ths || v.y >= ths || v.z >= ths || v.w >= ths) {
					push(4 * i4 + 0, v.x);
					push(4 * i4 + 1, v.y);
					push(4 * i4 + 2, v.z);
					push(4 * i4 + 3, v.w);
				}
			}
		}
		__syncthreads();
		const uint32_t total = s_ncand;

		if (total != 0 && total <= CAND_SMEM) {
			/* Common case: the survivors fit the shared buffer. */
			uint32_t n_emit = total;

			if (total > p.k) {
				uint32_t npow2 = 2;

				while (npow2 < total)
					npow2 <<= 1;
				for (uint32_t i = total + tid; i < npow2; i += TILE_THREADS)
					s_cand[i] = 0;
				bitonic_sort_desc(s_cand, npow2);
				n_emit = p.k;
			}
			if (tid == 0)
				s_base = atomicAdd(p.cand_count + slot, n_emit);
			__syncthreads();
			unsigned long long *out = p.cand +
			    (unsigned long long)slot * p.cand_cap + s_base;
			for (uint32_t i = tid; i < n_emit; i += TILE_THREADS)
				out[i] = s_cand[i];
			if (total > p.k && tid == 0)
				atomicMax(p.thr + slot, s_cand[p.k - 1]);
		} else if (total != 0) {
			/*
			 * Overflow (typically the first tiles of a query, before
			 * any threshold exists): exact MSB-first radix select of
			 * the k-th largest 46-bit key over the whole tile --
			 * digits 8,8,8,8 over the score bits, 7,7 over the id.
			 */
			unsigned long long kth = 0;	// emit keys >= kth
			uint32_t n_emit = total;

			if (total > p.k) {
				const int shifts[6] = { 38, 30, 22, 14, 7, 0 };
				const int widths[6] = { 8, 8, 8, 8, 7, 7 };

				if (tid == 0) {
					s_prefix = 0;
					s_want = p.k;
				}
				for (int ps = 0; ps < 6; ps++) {
					const int sh = shifts[ps], wd = widths[ps];

					if (tid < 256)
						s_hist[tid] = 0;
					__syncthreads();
					const unsigned long long prefix = s_prefix;
					for (uint32_t i = tid; i < TILE_DOCS; i += TILE_THREADS) {
						const float v = acc[i];
						const unsigned long long sk = sel_key(v, i);

						if (candidate(i, v) && (sk >> (sh + wd)) == prefix)
							atomicAdd(&s_hist[(uint32_t)(sk >> sh) &
							    ((1u << wd) - 1)], 1u);
					}
					__syncthreads();
					if (tid == 0) {
						uint32_t want = s_want, cum = 0;
						int b = (1 << wd) - 1;

						for (; b > 0; b--) {
							if (cum + s_hist[b] >= want)
								break;
							cum += s_hist[b];
						}
						s_want = want - cum;
						s_prefix = (prefix << wd) | (unsigned)b;
					}
					__syncthreads();
				}
				kth = s_prefix;
				n_emit = p.k;
			}
			if (tid == 0)
				s_base = atomicAdd(p.cand_count + slot, n_emit);
			__syncthreads();
			unsigned long long *out = p.cand +
			    (unsigned long long)slot * p.cand_cap + s_base;
			for (uint32_t i = tid; i < TILE_DOCS; i += TILE_THREADS) {
				const float v = acc[i];

				if (candidate(i, v) && sel_key(v, i) >= kth)
					out[atomicAdd(&s_emit, 1u)] = make_key(v, tile_lo + i);
			}
			if (total > p.k && tid == 0) {
				/* Publish the tile's k-th best as the new lower bound. */
				const uint32_t local = (uint32_t)kth & (TILE_DOCS - 1);

				atomicMax(p.thr + slot,
				    ((kth >> TILE_SHIFT) << 32) | (tile_lo + local));
			}
		}

		/* Restore the all-zero invariant for the next item. */
		__syncthreads();
		{
			float4 *a4 = reinterpret_cast<float4 *>(acc);
			for (uint32_t i = tid; i < TILE_DOCS / 4; i += TILE_THREADS)
				a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
		}
	}
}

#endif
