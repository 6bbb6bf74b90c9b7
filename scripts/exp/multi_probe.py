"""Dev probe: the replicated engine's host overhead.  Same GPU, N replicas
(NXS_GPU_DEVICES=0,0,...) against one engine, pipelined batch calls."""
import ctypes as C, os, sys, time, tempfile, shutil
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from nxsearch_b200 import capi, tools
import bench
docs = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
corpus = tools.Corpus.generate(docs, 1_000_000)
base = tempfile.mkdtemp(prefix="nxsb_mp_")
boot = capi.Nxs(base); boot.create_index("b").close(); boot.close()
corpus.write(f"{base}/data/b/nxsterms", f"{base}/data/b/nxsdtmap")
qs = bench.make_queries(corpus.query_terms(4 * 1024 * 16), 1024 * 16)
strings = [bench.query_string(corpus, q).encode() for q in qs]
for reps in (1, 2, 4):
    if reps > 1:
        os.environ["NXS_GPU_DEVICES"] = ",".join(["0"] * reps)
    nxs = capi.Nxs(base); idx = nxs.open_index("b")
    per = 1024 * reps
    arrays = [(C.c_char_p * per)(*strings[i * per:(i + 1) * per]) for i in range(len(strings) // per)]
    params = dict(algo="BM25", fuzzymatch=False)
    idx.search_batch_arrays(arrays[0], 10, **params)
    steps = 40
    t0 = time.perf_counter()
    ticket = idx.search_batch_begin(arrays[0], 10, **params)
    for s in range(steps):
        nxt = idx.search_batch_begin(arrays[(s + 1) % len(arrays)], 10, **params) if s + 1 < steps else None
        idx.search_batch_end_arrays(ticket); ticket = nxt
    dt = time.perf_counter() - t0
    print(f"replicas {reps}: {per} queries/call, {per * steps / dt:.0f} q/s, {dt / steps * 1e3:.2f} ms/call", flush=True)
    idx.close(); nxs.close()
shutil.rmtree(base, ignore_errors=True)
