"""Experiment (CPU, numpy): how much of a C2 batch survives exact top-k pruning
with per-(term, block) score maxima, for several block sizes."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from nxsearch_b200 import tools
sys.path.insert(0, "/root/repo")
import importlib.util
spec = importlib.util.spec_from_file_location("bench", "/root/repo/bench.py")

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
V = 1_000_000
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 256
t0 = time.time()
corpus = tools.Corpus.generate(N, V)
print("gen", time.time() - t0, corpus.n_pairs, flush=True)
df = np.asarray(corpus.term_df)
qt = corpus.query_terms(4 * 1024)
OP_OR = -3
queries = []
pos = 0
for i in range(1024):
    nt = 1 + (i % 4)
    leaves = [int(t) for t in qt[pos:pos + nt]]; pos += nt
    toks = []
    for t in reversed(leaves):
        if t not in toks: toks.append(t)
    queries.append(toks)
queries = queries[:NQ]
need = np.unique(np.concatenate([np.array(q, dtype=np.uint32) for q in queries]))
print("distinct terms", len(need), "sum df", df[need - 1].sum(), flush=True)
pairs = corpus.pairs.reshape(-1, 2)
terms = pairs[:, 0]; cnt = pairs[:, 1]
doc_of = np.repeat(np.arange(N, dtype=np.uint32), np.diff(corpus.doc_off).astype(np.int64))
t0 = time.time()
lut = np.zeros(V + 1, dtype=bool); lut[need] = True
m = lut[terms]
ft, fd, fc = terms[m], doc_of[m], cnt[m]
del m, doc_of
order = np.argsort(ft, kind="stable")
ft, fd, fc = ft[order], fd[order], fc[order]
starts = np.searchsorted(ft, need); ends = np.searchsorted(ft, need, side="right")
plist = {int(t): (fd[s:e], fc[s:e]) for t, s, e in zip(need, starts, ends)}
print("index", time.time() - t0, flush=True)
dl = np.asarray(corpus.doc_len).astype(np.float64)
adl = float(corpus.token_count // corpus.doc_count)
k1, b = 1.2000000476837158, 0.75
def weights(t):
    d, c = plist[t]
    tf = np.log(c.astype(np.float64) + 1)
    idf = np.log((N - df[t-1] + 0.5) / (df[t-1] + 0.5) + 1)
    return d, (tf / (tf + k1 * (1 - b + b * dl[d] / adl)) * idf).astype(np.float32)

K = 10
for BS in (16384, 2048, 256, 64):
    nb = (N + BS - 1) // BS
    tot_post = 0; surv_post_final = 0; surv_post_dyn = 0; tot_blocks = 0; surv_blocks = 0
    per_nt = {1: [0, 0], 2: [0, 0], 3: [0, 0], 4: [0, 0]}
    for q in queries:
        acc = np.zeros(N, dtype=np.float32)
        ub = np.zeros(nb, dtype=np.float32)
        cntb = np.zeros(nb, dtype=np.int64)
        for t in q:
            d, w = weights(t)
            acc[d] += w
            bm = np.zeros(nb, dtype=np.float32)
            np.maximum.at(bm, d // BS, w)
            ub += bm
            cntb += np.bincount(d // BS, minlength=nb)
        nz = np.count_nonzero(acc)
        kk = min(K, nz)
        if kk == 0: continue
        thr = np.partition(acc, N - kk)[N - kk]
        keep = ub > thr  # ties lose except same block... optimistic
        # dynamic: process blocks high->low in waves of W blocks, threshold from what was seen
        total = cntb.sum(); tot_post += total
        sf = cntb[keep].sum(); surv_post_final += sf
        per_nt[len(q)][0] += total; per_nt[len(q)][1] += sf
        tot_blocks += np.count_nonzero(cntb); surv_blocks += np.count_nonzero(keep & (cntb > 0))
    print(f"BS={BS}: postings {tot_post/1e6:.1f}M, survive(final thr) {surv_post_final/1e6:.1f}M ({100*surv_post_final/tot_post:.1f}%), blocks {tot_blocks} -> {surv_blocks}", flush=True)
    for nt, (a, s) in per_nt.items():
        if a: print(f"   {nt}-term: {a/1e6:.1f}M -> {s/1e6:.1f}M ({100*s/a:.1f}%)", flush=True)
