"""Dev probe: how much of the scorer's work is threshold warm-up?  With the
keepthr build a batch's second run starts from its final thresholds."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from nxsearch_b200 import tools, engine as eng
import bench
c = tools.Corpus.generate(10_000_000, 1_000_000)
e = eng.Engine(0); e.load_corpus(c)
qt = c.query_terms(64 * 1024)
def uniq(ts):
    out = []
    for t in reversed(ts):
        if t not in out: out.append(int(t))
    return out
sets = {f"{nt}-term": [(uniq(qt[i * nt:(i + 1) * nt]), None) for i in range(1024)] for nt in (1, 2, 3, 4)}
sets["mix"] = [(t, p) for t, p, _ in bench.make_queries(c.query_terms(4 * 1024), 1024)]
for name, qs in sets.items():
    h = e.upload(eng.Batch.from_lists(eng.ALGO_BM25, 10, qs))
    for rep in range(3):
        e.pruning_stats(reset=True)
        e.run(h); e.sync()
        st = e.pruning_stats(); st.pop("phase_cycles", None)
        print(f"{name:8s} run {rep}: {e.timings(1)['score_tiles']:.3f} ms  blocks/q {st['blocks_scored'] / 1024:.1f} rounds/q {st['rounds'] / 1024:.1f}", flush=True)
    e.release(h)
e.close()
