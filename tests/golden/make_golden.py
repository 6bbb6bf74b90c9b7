#!/usr/bin/env python
"""Generate tests/golden/c1_reference.npz from the COMPILED REFERENCE.

Run in the build container (needs /root/reference, through oracle/_ref):

    python tests/golden/make_golden.py

What is recorded is what the reference's own nxs_index_search /
idxterm_fuzzysearch return (ref src/query/search.c:285-342,
src/index/idxterm.c:210-249) on the deterministic C1 corpus of SURVEY 8d
(10k documents, 50k-term Zipf vocabulary, seed "nxs_B200"), so the parity
tests keep a reference-made anchor where oracle/_ref cannot be rebuilt:

  or_*      400 OR queries (100 each of 1..4 terms), top-10, BM25 and TF-IDF
  bool_*    80 boolean queries (the four C3 templates), top-100, both algorithms
  fuzzy_*   400 misspelt terms -> the term id the reference picks (0 = none)

Queries are stored as strings (NUL-joined); results as (count, ids, f32 score
bits) padded to the limit.  The file is ~150 KB.
"""
from __future__ import annotations

import ctypes as C
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

N_OR, N_BOOL, N_FUZZY = 400, 80, 400
OUT = HERE / "c1_reference.npz"


def or_queries(corpus) -> list[str]:
    qt = corpus.query_terms(N_OR * 4, seed=0x6E78735F42323030 + 101)
    out, pos = [], 0
    for i in range(N_OR):
        nt = 1 + i // (N_OR // 4)
        out.append(" OR ".join(corpus.term(int(t)) for t in qt[pos:pos + nt]))
        pos += nt
    return out


def bool_queries(corpus) -> list[str]:
    qt = [int(t) for t in corpus.query_terms(6 * N_BOOL, seed=0x6E78735F42323030 + 102)]
    w = corpus.term
    out = []
    for i in range(N_BOOL):
        a, b, c, d, e, f = qt[6 * i: 6 * i + 6]
        out.append([f"{w(a)} AND {w(b)}", f"({w(a)} OR {w(b)}) AND {w(c)}", f"{w(a)} AND NOT {w(b)}",
                    f"({w(a)} OR {w(b)}) AND ({w(c)} OR {w(d)}) AND NOT ({w(e)} OR {w(f)})"][i % 4])
    return out


def run(idx, queries, limit, algo):
    cnt = np.zeros(len(queries), dtype=np.uint32)
    ids = np.zeros((len(queries), limit), dtype=np.uint64)
    bits = np.zeros((len(queries), limit), dtype=np.uint32)
    for i, q in enumerate(queries):
        res = idx.search(q, limit=limit, algo=algo, fuzzymatch=False)
        cnt[i] = len(res)
        for j, (d, s) in enumerate(res):
            ids[i, j] = d
            bits[i, j] = np.float32(s).view(np.uint32)
    return cnt, ids, bits


def main() -> None:
    import _oracle
    from nxsearch_b200 import capi, tools

    if _oracle.build_ref() is None:
        raise SystemExit("oracle/_ref is not available: run this where /root/reference exists")
    corpus = tools.Corpus.generate(10_000, 50_000)
    base = tempfile.mkdtemp(prefix="nxsb_golden_")
    nxs = capi.Nxs(base, lib=_oracle.ref())
    nxs.create_index("c1").close()
    corpus.write(f"{base}/data/c1/nxsterms", f"{base}/data/c1/nxsdtmap")
    idx = nxs.open_index("c1")

    out = {}
    oq, bq = or_queries(corpus), bool_queries(corpus)
    out["or_queries"] = np.frombuffer("\0".join(oq).encode(), dtype=np.uint8)
    out["bool_queries"] = np.frombuffer("\0".join(bq).encode(), dtype=np.uint8)
    for algo, key in (("BM25", "bm25"), ("TF-IDF", "tfidf")):
        out[f"or_{key}_count"], out[f"or_{key}_ids"], out[f"or_{key}_bits"] = run(idx, oq, 10, algo)
        out[f"bool_{key}_count"], out[f"bool_{key}_ids"], out[f"bool_{key}_bits"] = run(idx, bq, 100, algo)

    lib = _oracle.ref()
    lib.idxterm_fuzzysearch.restype = C.c_void_p
    lib.idxterm_fuzzysearch.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    fq = corpus.fuzzy_terms(N_FUZZY, seed=0x6E78735F42323030 + 103)
    picks = np.zeros(len(fq), dtype=np.uint32)
    for i, q in enumerate(fq):
        term = lib.idxterm_fuzzysearch(idx.h, q, len(q))
        picks[i] = C.cast(term, C.POINTER(C.c_uint32))[0] if term else 0   # idxterm_t.id, index.h:43
    out["fuzzy_queries"] = np.frombuffer(b"\0".join(fq), dtype=np.uint8)
    out["fuzzy_pick"] = picks
    out["corpus_shape"] = np.array([corpus.n_docs, corpus.n_terms, corpus.n_pairs, corpus.token_count], dtype=np.uint64)

    idx.close()
    nxs.close()
    shutil.rmtree(base, ignore_errors=True)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes); fuzzy hits {int((picks != 0).sum())}/{len(fq)}")


if __name__ == "__main__":
    main()
