"""How far is the kernels' fp32 BM25 from the reference's fp64 evaluation?

    python scripts/bm25_deviation.py [--docs N] > profiles/r2_bm25_deviation.json

Every (term count, document length) pair the synthetic C2 corpus can hold
(tf 1..111, dl 16..111, tf <= dl) and then some (tf up to 255, dl up to 4096),
times a ladder of document frequencies from 1 to N, through the kernels' own
arithmetic (nxsb_engine_score_pairs: st_score = log table, one FFMA,
rcp.approx, two FMUL, fp32 idf) against ref src/algo/ranking.c:135-176 in
float64, the result rounded to float as bm25()'s return value is.  The budget
north_star sets is 1e-5 relative; the tests state it, this script measures it.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from nxsearch_b200 import engine as eng, tools  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=200_000, help="image to take N, K0, K1 from")
    ap.add_argument("--vocab", type=int, default=50_000)
    args = ap.parse_args()
    corpus = tools.Corpus.generate(args.docs, args.vocab)
    e = eng.Engine(0)
    e.load_corpus(corpus)
    N = corpus.doc_count
    adl = float(corpus.token_count // corpus.doc_count)         # integer quotient, ranking.c:163
    k, b = np.float64(np.float32(1.2)), np.float64(np.float32(0.75))
    tfs, dls = [], []
    for dl in list(range(1, 256)) + [300, 512, 1000, 4096, 65535]:
        for tf in range(1, min(dl, 255) + 1):
            tfs.append(tf)
            dls.append(dl)
    tfs, dls = np.array(tfs, dtype=np.uint32), np.array(dls, dtype=np.uint32)
    dfs = sorted(set([1, 2, 3, 5, 10, 100, 1000, N // 100, N // 10, N // 3, N // 2, N - 1, N]))
    worst = {"rel": 0.0}
    hist = np.zeros(8, dtype=np.int64)            # relative error decades 1e-9 .. 1e-2
    total = 0
    for df in dfs:
        idf64 = np.log(((N - df + 0.5) / (df + 0.5)) + 1)
        idf32 = np.float32(idf64)                 # what the engine uploads (upload_stats)
        got = e.score_pairs(eng.ALGO_BM25, tfs, dls, np.full(len(tfs), idf32, dtype=np.float32)).astype(np.float64)
        T = np.log(tfs.astype(np.float64) + 1)
        ref = (T / (T + k * (1 - b + b * dls.astype(np.float64) / adl)) * idf64).astype(np.float32).astype(np.float64)
        rel = np.abs(got - ref) / np.maximum(ref, 1e-300)
        i = int(np.argmax(rel))
        if rel[i] > worst["rel"]:
            worst = {"rel": float(rel[i]), "tf": int(tfs[i]), "dl": int(dls[i]), "df": int(df),
                     "gpu": float(got[i]), "reference": float(ref[i])}
        hist += np.histogram(np.log10(np.maximum(rel, 1e-12)), bins=[-13, -9, -8, -7, -6, -5, -4, -3, 0])[0]
        total += len(rel)
    # TF-IDF must be exact
    idf_t = np.float32(np.log(np.float64(np.float32(N) / np.float32(max(N // 7, 1)))) + 1)
    got = e.score_pairs(eng.ALGO_TFIDF, tfs, dls, np.full(len(tfs), idf_t, dtype=np.float32))
    ref = np.log(tfs.astype(np.float64) + 1).astype(np.float32) * idf_t
    e.close()
    print(json.dumps({
        "what": "BM25: kernels' fp32 arithmetic vs ref ranking.c:135-176 in float64 (returned as float)",
        "N": int(N), "adl": adl, "pairs": int(len(tfs)), "document_frequencies": [int(d) for d in dfs],
        "evaluations": int(total), "max_relative_error": worst["rel"], "worst_case": worst,
        "budget": 1e-5, "within_budget": bool(worst["rel"] <= 1e-5),
        "relative_error_histogram": {"<1e-9": int(hist[0]), "1e-9..1e-8": int(hist[1]), "1e-8..1e-7": int(hist[2]),
                                     "1e-7..1e-6": int(hist[3]), "1e-6..1e-5": int(hist[4]), "1e-5..1e-4": int(hist[5]),
                                     "1e-4..1e-3": int(hist[6]), ">=1e-3": int(hist[7])},
        "tfidf_bit_exact": bool(np.array_equal(got.view(np.uint32), ref.astype(np.float32).view(np.uint32))),
    }, indent=1))


if __name__ == "__main__":
    main()
