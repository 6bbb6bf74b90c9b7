"""Dev probe: C3 (boolean TF-IDF top-100) through the pruned scorer, by template."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from nxsearch_b200 import tools, engine as eng
import bench
docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
c = tools.Corpus.generate(docs, 1_000_000)
df = np.asarray(c.term_df)
e = eng.Engine(0); e.load_corpus(c)
qt = c.query_terms(6 * 4096, seed=tools.SEED + 3)
allq = bench.make_queries(qt, 4096, "bool")
def run(name, qs, limit=100):
    h = e.upload(eng.Batch.from_lists(eng.ALGO_TFIDF, limit, [(t, p) for t, p, _ in qs]))
    for pr in (True, False):
        e.set_pruning(pr)
        h2 = e.upload(eng.Batch.from_lists(eng.ALGO_TFIDF, limit, [(t, p) for t, p, _ in qs]))
        for _ in range(2): e.run(h2)
        e.sync(); e.pruning_stats(reset=True)
        n = 4
        for _ in range(n): e.run(h2)
        e.sync()
        t = e.timings(n); st = e.pruning_stats(); ph = st.pop("phase_cycles", None)
        print(f"{name:14s} pruned={pr!s:5s} {sum(t.values()) / n:8.3f} ms  score {t.get('score_tiles', 0) / n:8.3f} topk {t.get('topk', 0) / n:6.3f}  "
              f"blocks/q {st['blocks_scored'] / n / len(qs):9.1f} rounds/q {st['rounds'] / n / len(qs):6.1f}", flush=True)
        if ph and sum(ph.values()):
            tot = sum(ph.values()); print("     ", {k: f"{100 * v / tot:.0f}%" for k, v in ph.items()})
        e.release(h2)
    e.release(h)
for s in range(4):
    run(["a AND b", "(a|b) AND c", "a AND NOT b", "(a|b)&(c|d)-(e|f)"][s], [allq[i] for i in range(s, 4096, 4)][:1024])
run("mix", allq[:1024])
run("mix top-10", allq[:1024], 10)
e.close()
