"""Dev probe: device time of the scoring kernel for one library build.

    NXSB_LIBRARY=.../libnxsearch.so python scripts/kbench.py [docs] [steps]

Prints ms per 1024-query C2 batch per kernel family (CUDA events inside the
engine).  Used to A/B kernel parameter variants (scripts/build_variant.sh)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from nxsearch_b200 import tools, engine as eng
import bench

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
c = tools.Corpus.generate(docs, 1_000_000)
nb = 8
qt = c.query_terms(4 * 1024 * nb)
qs = bench.make_queries(qt, 1024 * nb)
hb = [eng.Batch.from_lists(eng.ALGO_BM25, 10, [(t, p) for t, p, _ in qs[i * 1024:(i + 1) * 1024]]) for i in range(nb)]
e = eng.Engine(0); e.load_corpus(c)
hs = [e.upload(b) for b in hb]
for i in range(3): e.run(hs[i % nb])
e.sync()
for i in range(steps): e.run(hs[i % nb])
e.sync()
t = e.timings(steps)
print(os.environ.get("NXSB_LIBRARY", "default"), {k: round(v / steps, 3) for k, v in t.items()})

if hasattr(e.lib if hasattr(e, "lib") else None, "nxsb_engine_prof") or True:
    import ctypes as C
    from nxsearch_b200._lib import load_library
    lib = load_library()
    if hasattr(lib, "nxsb_engine_prof"):
        out = (C.c_uint64 * 32)()
        lib.nxsb_engine_prof.argtypes = [C.c_void_p, C.c_void_p]
        lib.nxsb_engine_prof(e.h if hasattr(e, "h") else e._h, out)
        names = ["other", "wait_full", "tok_barrier", "full_stage", "partial", "bar_pre_epi", "sparse_collect",
                 "scan_zero", "bar_post", "rank_emit", "dense_run", "zero_fill", "P_other", "P_wait_empty"]
        cons = sum(out[i] for i in range(12)); prod = out[12] + out[13]
        print("consumer:", {names[i]: round(100 * out[i] / cons, 1) for i in range(12)})
        print("producer:", {names[i]: round(100 * out[i] / prod, 1) for i in (12, 13)})
