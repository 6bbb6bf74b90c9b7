"""TEST INFRASTRUCTURE: the committed reference-made fixtures under tests/golden/.

tests/golden/c1_reference.npz was produced by tests/golden/make_golden.py from
the COMPILED reference (oracle/_ref) on the deterministic C1 corpus; it is the
anchor that does not depend on /root/reference or oracle/_ref being present.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

PATH = Path(__file__).resolve().parent / "golden" / "c1_reference.npz"


class Golden:
    def __init__(self):
        z = np.load(PATH)
        self.or_queries = bytes(z["or_queries"]).decode().split("\0")
        self.bool_queries = bytes(z["bool_queries"]).decode().split("\0")
        self.fuzzy_queries = bytes(z["fuzzy_queries"]).split(b"\0")
        self.fuzzy_pick = z["fuzzy_pick"]
        self.corpus_shape = tuple(int(x) for x in z["corpus_shape"])
        self._z = z

    def results(self, family: str, algo: str, i: int) -> list[tuple[int, float]]:
        """What the reference returned for query i: [(doc id, f32 score)] in its order."""
        n = int(self._z[f"{family}_{algo}_count"][i])
        ids = self._z[f"{family}_{algo}_ids"][i, :n]
        sc = self._z[f"{family}_{algo}_bits"][i, :n].view(np.float32)
        return list(zip(ids.tolist(), sc.tolist()))

    def check_corpus(self, corpus) -> None:
        """The fixtures are only meaningful on the corpus they were made from."""
        shape = (corpus.n_docs, corpus.n_terms, corpus.n_pairs, corpus.token_count)
        assert shape == self.corpus_shape, f"C1 generator drifted: {shape} != {self.corpus_shape}"
