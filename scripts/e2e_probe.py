"""Dev probe: where does the end-to-end time of a 1024-query batch go?"""
import sys, time, tempfile, shutil
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from nxsearch_b200 import tools, engine as eng, capi
import bench

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
c = tools.Corpus.generate(docs, 1_000_000)
qt = c.query_terms(4 * 1024 * 8)
qs = bench.make_queries(qt, 1024 * 8)
batches = [qs[i * 1024:(i + 1) * 1024] for i in range(8)]
hb = [eng.Batch.from_lists(eng.ALGO_BM25, 10, [(t, p) for t, p, _ in b]) for b in batches]
e = eng.Engine(0); e.load_corpus(c)
for b in hb[:3]: e.search(b)
t0 = time.perf_counter()
for i in range(10): e.search(hb[i % 8])
dt = (time.perf_counter() - t0) / 10
print(f"engine C ABI (host descriptors -> host results): {dt*1e3:.2f} ms/batch", e.timings(1))
base = tempfile.mkdtemp()
nx = capi.Nxs(base); nx.create_index("b").close()
c.write(f"{base}/data/b/nxsterms", f"{base}/data/b/nxsdtmap")
ix = nx.open_index("b")
strings = [[" OR ".join(c.term(t) for t in lv).encode() for _, _, lv in b] for b in batches]
for s in strings[:3]: ix.search_batch(s, limit=10, algo="BM25", fuzzymatch=False)
t0 = time.perf_counter()
for i in range(10): ix.search_batch(strings[i % 8], limit=10, algo="BM25", fuzzymatch=False)
dt = (time.perf_counter() - t0) / 10
print(f"C API nxs_index_search_batch (strings -> python lists): {dt*1e3:.2f} ms/batch")
# the same without building python result lists
import ctypes as C
lib = ix._lib
arrs = [(C.c_char_p * 1024)(*s) for s in strings]
out = (C.c_void_p * 1024)()
p = capi.Params(lib, limit=10, algo="BM25", fuzzymatch=False)
t0 = time.perf_counter()
for i in range(10):
    lib.nxs_index_search_batch(ix.h, p.h, arrs[i % 8], 1024, out)
    for h in out: lib.nxs_resp_release(h)
dt = (time.perf_counter() - t0) / 10
print(f"C API nxs_index_search_batch (raw, responses released): {dt*1e3:.2f} ms/batch")
shutil.rmtree(base)
