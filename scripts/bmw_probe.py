"""Dev probe: the block-max scorer on C2 batches -- device time per batch,
blocks / postings actually scored, per block size (NXSB_BMW_SHIFT).

    python scripts/bmw_probe.py [docs] [steps] [shift ...]
"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from nxsearch_b200 import tools, engine as eng
import bench

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
shifts = [int(x) for x in sys.argv[3:]] or [6]
c = tools.Corpus.generate(docs, 1_000_000)
nb = 8
qt = c.query_terms(4 * 1024 * nb)
qs = bench.make_queries(qt, 1024 * nb)
hb = [eng.Batch.from_lists(eng.ALGO_BM25, 10, [(t, p) for t, p, _ in qs[i * 1024:(i + 1) * 1024]]) for i in range(nb)]
named = sum(int(c.term_df[t - 1]) for t, _, _ in qs for t in t) / nb
import time
primes = [int(x) for x in os.environ.get("PROBE_PRIME", "1").split(",")]
for shift, prime in [(s, p) for s in shifts for p in primes]:
    os.environ["NXSB_BMW_SHIFT"] = str(shift)
    os.environ["NXSB_PRIME"] = str(prime)
    t0 = time.time()
    e = eng.Engine(0); e.load_corpus(c)
    print(f"prime {prime}: image built in {time.time() - t0:.2f}s")
    hs = [e.upload(b) for b in hb]
    for i in range(3): e.run(hs[i % nb])
    e.sync()
    e.pruning_stats(reset=True)
    for i in range(steps): e.run(hs[i % nb])
    e.sync()
    t = e.timings(steps)
    st = e.pruning_stats()
    ph = st.pop("phase_cycles", None)
    if ph:
        tot = sum(ph.values()) or 1
        print("   phases:", {k: f"{100 * v / tot:.1f}%" for k, v in ph.items()})
    print(f"shift {shift}:", {k: round(v / steps, 3) for k, v in t.items()},
          {k: round(v / steps) for k, v in st.items()},
          f"postings named per batch {named:.3g}, scored {st['postings_scored'] / steps / named:.4%}", flush=True)
    for h in hs: e.release(h)
    e.close()
