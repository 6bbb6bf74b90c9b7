import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from nxsearch_b200 import tools, engine, dist as nxdist
from test_gpu_engine import c1_queries
corpus = tools.Corpus.generate(10_000, 50_000)
qs = c1_queries(corpus, 400)[200:328]
lo, hi = nxdist.shard_range(corpus.n_docs, 1, 8)
e = engine.Engine(0)
e.load_corpus(corpus, lo=lo, hi=hi, df=corpus.term_df, token_count=corpus.token_count, doc_count=corpus.doc_count)
for algo, k in ((0, 100), (1, 10)):
    batch = engine.Batch.from_lists(algo, k, qs)
    h = e.upload(batch); e.run(h); g2 = e.fetch(h, len(qs), k); e.release(h)
print("ran")
