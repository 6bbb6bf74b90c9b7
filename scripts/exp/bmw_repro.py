import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from nxsearch_b200 import tools, engine
import test_gpu_bmw as T
corpus = tools.Corpus.generate(T.N_DOCS, T.N_TERMS)
e = engine.Engine(0); e.load_corpus(corpus)
qs = T.or_queries(corpus, 512)
for algo, limit in ((0, 100), (1, 128), (0, 10), (0,100)):
    batch = engine.Batch.from_lists(algo, limit, qs)
    e.set_pruning(False); full = e.search(batch)
    e.set_pruning(True)
    for rep in range(3):
        got = e.search(batch)
        bad = [q for q in range(512) if not np.array_equal(got[1][q,:got[0][q]], full[1][q,:full[0][q]])]
        print(algo, limit, rep, "bad queries:", bad[:10], len(bad))
        for q in bad[:2]:
            toks = qs[q][0]
            print("  q", q, "toks", toks, "df", [int(corpus.term_df[t-1]) for t in toks])
            g, f = got[1][q,:got[0][q]], full[1][q,:full[0][q]]
            miss = [int(x) for x in f if x not in set(g.tolist())]
            print("  missing", miss[:10], "scores full", full[2][q,:5], "got", got[2][q,:5])
