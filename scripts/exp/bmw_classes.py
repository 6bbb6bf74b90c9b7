"""Dev probe: where the block-max scorer's time goes by query shape -- batches
of 1024 queries with exactly n terms, and 2-term batches by term popularity.

    python scripts/exp/bmw_classes.py [docs]
"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from nxsearch_b200 import tools, engine as eng

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
c = tools.Corpus.generate(docs, 1_000_000)
df = np.asarray(c.term_df)
qt = c.query_terms(64 * 1024)
col_min = max((docs >> 5) // 2, 64)
e = eng.Engine(0); e.load_corpus(c)


def run(name, qs):
    b = eng.Batch.from_lists(eng.ALGO_BM25, 10, qs)
    h = e.upload(b)
    for _ in range(2): e.run(h)
    e.sync(); e.pruning_stats(reset=True)
    n = 8
    for _ in range(n): e.run(h)
    e.sync()
    t = e.timings(n); st = e.pruning_stats(); ph = st.pop("phase_cycles", None)
    if ph and sum(ph.values()):
        tot = sum(ph.values())
        print("     ", {k: f"{100 * v / tot:.0f}%" for k, v in ph.items()})
    named = sum(int(df[t - 1]) for toks, _ in qs for t in toks)
    print(f"{name:28s} {t['score_tiles'] / n:7.3f} ms  blocks/q {st['blocks_scored'] / n / len(qs):8.1f}  "
          f"postings/q {st['postings_scored'] / n / len(qs):9.0f}  rounds/q {st['rounds'] / n / len(qs):6.1f}  named/q {named / len(qs):.3g}", flush=True)
    e.release(h)


def uniq(ts):
    out = []
    for t in reversed(ts):
        if t not in out: out.append(int(t))
    return out


for nt in (1, 2, 3, 4):
    qs = [(uniq(qt[i * nt:(i + 1) * nt]), None) for i in range(1024)]
    run(f"{nt}-term", qs)
head = [int(t) for t in qt if df[t - 1] >= col_min]
mid = [int(t) for t in qt if 2048 <= df[t - 1] < col_min]
rare = [int(t) for t in qt if df[t - 1] < 2048]
print("share of query terms: column", len(head) / len(qt), "mid", len(mid) / len(qt), "rare", len(rare) / len(qt))
for na, a in (("col", head), ("mid", mid), ("rare", rare)):
    run(f"1-term {na}", [([a[i]], None) for i in range(1024)])
for (na, a), (nb, b) in ((("col", head), ("col", head)), (("col", head), ("mid", mid)), (("col", head), ("rare", rare)),
                         (("mid", mid), ("mid", mid)), (("mid", mid), ("rare", rare)), (("rare", rare), ("rare", rare))):
    qs = [(uniq([a[2 * i], b[2 * i + 1]]), None) for i in range(1024)]
    run(f"2-term {na}+{nb}", qs)
qs = [(uniq([head[3 * i], head[3 * i + 1], head[3 * i + 2]]), None) for i in range(1024)]
run("3-term col+col+col", qs)
qs = [(uniq([head[3 * i], head[3 * i + 1], rare[i]]), None) for i in range(1024)]
run("3-term col+col+rare", qs)
qs = [(uniq([head[3 * i], mid[3 * i + 1], rare[i]]), None) for i in range(1024)]
run("3-term col+mid+rare", qs)
e.close()
