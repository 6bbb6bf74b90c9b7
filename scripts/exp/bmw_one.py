"""Dev probe for ncu: one query shape, a few launches.  bmw_one.py SHAPE"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from nxsearch_b200 import tools, engine as eng
import bench
shape = sys.argv[1] if len(sys.argv) > 1 else "mix"
docs = 10_000_000
c = tools.Corpus.generate(docs, 1_000_000)
df = np.asarray(c.term_df)
qt = c.query_terms(64 * 1024)
col_min = max((docs >> 5) // 2, 64)
e = eng.Engine(0); e.load_corpus(c)
def uniq(ts):
    out = []
    for t in reversed(ts):
        if t not in out: out.append(int(t))
    return out
head = [int(t) for t in qt if df[t - 1] >= col_min]
if shape == "ccc":
    qs = [(uniq([head[3 * i], head[3 * i + 1], head[3 * i + 2]]), None) for i in range(1024)]
else:
    qs = [(t, p) for t, p, _ in bench.make_queries(c.query_terms(4 * 1024), 1024)]
h = e.upload(eng.Batch.from_lists(eng.ALGO_BM25, 10, qs))
for _ in range(4): e.run(h)
e.sync()
print(shape, e.timings(4)["score_tiles"] / 4)
e.close()
