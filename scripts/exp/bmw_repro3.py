import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from nxsearch_b200 import tools, engine, dist as nxdist
from test_gpu_engine import c1_queries
corpus = tools.Corpus.generate(10_000, 50_000)
qs = c1_queries(corpus, 400)
whole = engine.Engine(0); whole.load_corpus(corpus)
shards = []
for g in range(8):
    lo, hi = nxdist.shard_range(corpus.n_docs, g, 8)
    e = engine.Engine(0)
    e.load_corpus(corpus, lo=lo, hi=hi, df=corpus.term_df, token_count=corpus.token_count, doc_count=corpus.doc_count)
    shards.append((e, lo, hi))
nbad = 0
for rep in range(12):
    for algo in (1, 0):
        for k in (10, 100):
            batch = engine.Batch.from_lists(algo, k, qs)
            whole.set_pruning(False); ref = whole.search(batch); whole.set_pruning(True)
            got = whole.search(batch)
            for q in range(len(qs)):
                if not np.array_equal(got[1][q,:got[0][q]], ref[1][q,:ref[0][q]]):
                    nbad += 1; print("WHOLE mismatch rep", rep, algo, k, "q", q, qs[q][0])
            for (e, lo, hi) in shards:
                e.set_pruning(False); r2 = e.search(batch); e.set_pruning(True)
                h = e.upload(batch); e.run(h); g2 = e.fetch(h, len(qs), k); e.release(h)
                for q in range(len(qs)):
                    if not (np.array_equal(g2[1][q,:g2[0][q]], r2[1][q,:r2[0][q]]) and np.array_equal(g2[2][q,:g2[0][q]].view(np.uint32), r2[2][q,:r2[0][q]].view(np.uint32))):
                        nbad += 1
                        gg, ff = g2[1][q,:g2[0][q]], r2[1][q,:r2[0][q]]
                        print("SHARD", lo, "mismatch rep", rep, algo, k, "q", q, qs[q][0], "missing", [int(x) for x in ff if x not in set(gg.tolist())][:6], "extra", [int(x) for x in gg if x not in set(ff.tolist())][:6], "counts", g2[0][q], r2[0][q])
print("total bad", nbad)
