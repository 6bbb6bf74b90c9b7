import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from nxsearch_b200 import tools, engine, dist as nxdist
from test_gpu_engine import c1_queries
corpus = tools.Corpus.generate(10_000, 50_000)
qs = c1_queries(corpus, 400)
def check(e, name):
    for algo in (1, 0):
        for k in (10, 100):
            batch = engine.Batch.from_lists(algo, k, qs)
            e.set_pruning(False); full = e.search(batch)
            e.set_pruning(True); got = e.search(batch)
            bad = [q for q in range(len(qs)) if not (np.array_equal(got[1][q,:got[0][q]], full[1][q,:full[0][q]]) and got[0][q] == full[0][q])]
            if bad:
                print(name, "algo", algo, "k", k, "bad", bad[:8], len(bad))
                q = bad[0]; toks = qs[q][0]
                g, f = got[1][q,:got[0][q]], full[1][q,:full[0][q]]
                print("   toks", toks, "counts", got[0][q], full[0][q])
                print("   missing", [int(x) for x in f if x not in set(g.tolist())][:10], "extra", [int(x) for x in g if x not in set(f.tolist())][:10])
whole = engine.Engine(0); whole.load_corpus(corpus); check(whole, "whole")
print("df of rank terms", [int(corpus.term_df[i]) for i in range(8)])
for g in range(8):
    lo, hi = nxdist.shard_range(corpus.n_docs, g, 8)
    e = engine.Engine(0)
    e.load_corpus(corpus, lo=lo, hi=hi, df=corpus.term_df, token_count=corpus.token_count, doc_count=corpus.doc_count)
    check(e, f"shard{g}")
    e.close()
print("done")
