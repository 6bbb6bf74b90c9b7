"""Dev probe: the heavy query shapes only (see bmw_classes.py)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from nxsearch_b200 import tools, engine as eng
import bench

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
def run(name, qs):
    b = eng.Batch.from_lists(eng.ALGO_BM25, 10, qs)
    h = e.upload(b)
    for _ in range(2): e.run(h)
    e.sync(); n = 8
    for _ in range(n): e.run(h)
    e.sync()
    t = e.timings(n)
    print(f"{name:24s} {t['score_tiles'] / n:7.3f} ms", flush=True)
    e.release(h)
head = [int(t) for t in qt if df[t - 1] >= col_min]
mid = [int(t) for t in qt if 2048 <= df[t - 1] < col_min]
rare = [int(t) for t in qt if df[t - 1] < 2048]
for nt in (1, 4):
    run(f"{nt}-term", [(uniq(qt[i * nt:(i + 1) * nt]), None) for i in range(1024)])
run("col+col+col", [(uniq([head[3 * i], head[3 * i + 1], head[3 * i + 2]]), None) for i in range(1024)])
run("mid+mid", [(uniq([mid[2 * i], mid[2 * i + 1]]), None) for i in range(1024)])
run("col+rare", [(uniq([head[2 * i], rare[2 * i + 1]]), None) for i in range(1024)])
qs = bench.make_queries(c.query_terms(4 * 1024), 1024)
run("C2 mix", [(t, p) for t, p, _ in qs])
e.close()
