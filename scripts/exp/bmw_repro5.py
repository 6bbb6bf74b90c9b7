import sys, os, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from nxsearch_b200 import tools, engine, dist as nxdist
from test_gpu_engine import c1_queries
import _oracle
corpus = tools.Corpus.generate(10_000, 50_000)
qs = c1_queries(corpus, 400)
lo, hi = nxdist.shard_range(corpus.n_docs, 1, 8)
e = engine.Engine(0)
e.load_corpus(corpus, lo=lo, hi=hi, df=corpus.term_df, token_count=corpus.token_count, doc_count=corpus.doc_count)
def same(a, b, q):
    return a[0][q] == b[0][q] and np.array_equal(a[1][q,:a[0][q]], b[1][q,:b[0][q]])
cnt = {"pruned_search": 0, "pruned_run": 0, "full_search": 0, "full_run": 0}
for rep in range(15):
    for algo, k in ((0, 100), (1, 10), (0, 10), (1, 100)):
        batch = engine.Batch.from_lists(algo, k, qs)
        res = {}
        for mode in ("pruned", "full"):
            e.set_pruning(mode == "pruned")
            res[mode + "_search"] = e.search(batch)
            h = e.upload(batch); e.run(h); res[mode + "_run"] = e.fetch(h, len(qs), k); e.release(h)
        for q in range(len(qs)):
            votes = {n: sum(same(res[n], res[m], q) for m in res) for n in res}
            best = max(votes.values())
            for n, v in votes.items():
                if v < best:
                    cnt[n] += 1
                    print("odd one out:", n, "rep", rep, "algo", algo, "k", k, "q", q, qs[q][0], votes)
print(cnt)
