"""A small workload for compute-sanitizer over the pruned scorer (OR and
boolean batches, both algorithms, several limits), the weight-table kernels of
the image build and the fuzzy scan split over CTAs.

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitize_bmw.py
"""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from nxsearch_b200 import tools, engine as eng
from test_gpu_stream import bool_queries, or_queries

corpus = tools.Corpus.generate(120_000, 20_000)        # 3750 blocks, long + short + column lists
e = eng.Engine(0)
e.load_corpus(corpus)
ors, bools = or_queries(corpus, 48), bool_queries(corpus, 48)
for algo, k in ((eng.ALGO_BM25, 10), (eng.ALGO_TFIDF, 100)):
    for qs in (ors, bools, ors[:8] + bools[:8]):
        b = eng.Batch.from_lists(algo, k, qs)
        pruned = e.search(b)
        if os.environ.get("SAN_PRUNED_ONLY"):      # racecheck: keep the (TMA) stream kernel's known reports out
            continue
        e.set_pruning(False)
        full = e.search(b)
        e.set_pruning(True)
        assert np.array_equal(pruned[0], full[0])
        for q in range(len(qs)):
            n = int(pruned[0][q])
            assert np.array_equal(pruned[1][q, :n], full[1][q, :n]), q
parent, edge, rank = corpus.bk_mirror()
e.load_vocab(corpus.term_blob, corpus.term_off, corpus.term_total, parent, edge, rank)
e.fuzzy(corpus.fuzzy_terms(40))
e.fuzzy_candidates(corpus.fuzzy_terms(8), cap=512)
e.close()
print("ran")
