"""TEST INFRASTRUCTURE: ctypes access to the oracle.

* ``port()``  -- oracle/_build/libnxs_oracle.so, the CPU restatement
  (oracle/nxs_oracle.c); built on demand with gcc.
* ``ref()``   -- oracle/_ref/libnxsearch_ref.so, the reference's own C files
  compiled with the shims under oracle/shims (only buildable where
  /root/reference exists; the built .so travels to the GPU box).  Returns
  None when absent.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE = ROOT / "oracle"
PORT_SO = ORACLE / "_build" / "libnxs_oracle.so"
REF_SO = ORACLE / "_ref" / "libnxsearch_ref.so"
REFERENCE_SRC = Path("/root/reference/src")

OP_EMPTY, OP_AND, OP_OR, OP_ANDNOT = -1, -2, -3, -4
TFIDF, BM25 = 0, 1

_port = None
_ref = None


def build_port() -> Path:
    src = ORACLE / "nxs_oracle.c"
    if not PORT_SO.exists() or PORT_SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-s", "-C", str(ORACLE), "port"], check=True)
    return PORT_SO


def build_ref() -> Path | None:
    """Compile the reference where its sources exist; otherwise use what is there."""
    if REFERENCE_SRC.exists():
        subprocess.run(["make", "-s", "-j8", "-C", str(ORACLE), "ref"], check=True)
    return REF_SO if REF_SO.exists() else None


def port():
    global _port
    if _port is None:
        lib = C.CDLL(str(build_port()))
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        lib.ora_index_build.restype = vp
        lib.ora_index_build.argtypes = [u32, vp, vp, vp, vp, u32, vp, vp, vp, u64, u32]
        lib.ora_index_free.argtypes = [vp]
        lib.ora_term_df.restype = u32
        lib.ora_term_df.argtypes = [vp, u32]
        lib.ora_score_pair.restype = C.c_float
        lib.ora_score_pair.argtypes = [vp, C.c_int, u32, u32, u32]
        for name in ("ora_search", "ora_search_all"):
            fn = getattr(lib, name)
            fn.restype = C.c_int64
        lib.ora_search.argtypes = [vp, C.c_int, u64, u32, vp, u32, vp, vp, vp, C.c_size_t]
        lib.ora_search_all.argtypes = [vp, C.c_int, u32, vp, u32, vp, vp, vp, C.c_size_t]
        lib.ora_heap_topn.restype = C.c_size_t
        lib.ora_heap_topn.argtypes = [C.c_size_t, C.c_size_t, vp, vp, vp, vp]
        lib.ora_levdist.restype = C.c_int
        lib.ora_levdist.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        lib.ora_fuzzy.restype = u32
        lib.ora_fuzzy.argtypes = [vp, C.c_char_p, C.c_size_t, vp, vp, C.c_size_t,
                                  C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        lib.ora_term_lookup.restype = u32
        lib.ora_term_lookup.argtypes = [vp, C.c_char_p, C.c_size_t]
        _port = lib
    return _port


def ref():
    """The compiled reference bound through the same prototypes as the product."""
    global _ref
    if _ref is None:
        if not REF_SO.exists():
            return None
        from nxsearch_b200 import capi
        _ref = capi.bind(C.CDLL(str(REF_SO)))
    return _ref


class OracleIndex:
    """The CPU restatement over a tools.Corpus-like object."""

    def __init__(self, corpus):
        self.lib = port()
        self.corpus = corpus
        blob = C.create_string_buffer(corpus.term_blob, len(corpus.term_blob) + 1)
        self._keep = (blob,)
        self.h = self.lib.ora_index_build(
            corpus.n_docs, corpus.doc_ids.ctypes.data, corpus.doc_len.ctypes.data,
            corpus.doc_off.ctypes.data, corpus.pairs.ctypes.data, corpus.n_terms,
            C.cast(blob, C.c_void_p), corpus.term_off.ctypes.data,
            corpus.term_total.ctypes.data, corpus.token_count, corpus.doc_count)

    @staticmethod
    def or_program(n):
        prog = []
        for s in range(n):
            prog.append(s)
            if s:
                prog.append(OP_OR)
        return prog

    def search(self, algo, limit, tokens, prog=None):
        tokens = np.ascontiguousarray(tokens, dtype=np.uint32)
        prog = np.ascontiguousarray(self.or_program(len(tokens)) if prog is None else prog, dtype=np.int32)
        cap = int(min(limit, self.corpus.n_docs)) or 1
        ids = np.zeros(cap, dtype=np.uint64)
        sc = np.zeros(cap, dtype=np.float32)
        n = self.lib.ora_search(self.h, algo, limit, len(tokens), tokens.ctypes.data, len(prog),
                                prog.ctypes.data, ids.ctypes.data, sc.ctypes.data, cap)
        assert n >= 0, "malformed program"
        return ids[:n].copy(), sc[:n].copy()

    def search_all(self, algo, tokens, prog=None):
        tokens = np.ascontiguousarray(tokens, dtype=np.uint32)
        prog = np.ascontiguousarray(self.or_program(len(tokens)) if prog is None else prog, dtype=np.int32)
        cap = max(self.corpus.n_docs, 1)
        ids = np.empty(cap, dtype=np.uint64)        # untouched pages cost nothing
        sc = np.empty(cap, dtype=np.float32)
        n = self.lib.ora_search_all(self.h, algo, len(tokens), tokens.ctypes.data, len(prog),
                                    prog.ctypes.data, ids.ctypes.data, sc.ctypes.data, cap)
        assert n >= 0, "malformed program"
        return ids[:n].copy(), sc[:n].copy()

    def fuzzy(self, q: bytes, cap: int = 4096):
        cands = np.zeros(cap, dtype=np.uint32)
        dists = np.zeros(cap, dtype=np.uint32)
        nc, nv = C.c_size_t(), C.c_size_t()
        t = self.lib.ora_fuzzy(self.h, q, len(q), cands.ctypes.data, dists.ctypes.data, cap,
                               C.byref(nc), C.byref(nv))
        n = min(nc.value, cap)
        return t, cands[:n].copy(), dists[:n].copy(), nv.value

    def fuzzy_true(self, q: bytes, cap: int = 65536):
        """Every term within distance 2 of q by brute force: (term ids, distances), id order."""
        self.lib.ora_fuzzy_true.restype = C.c_size_t
        self.lib.ora_fuzzy_true.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        terms = np.zeros(cap, dtype=np.uint32)
        dists = np.zeros(cap, dtype=np.uint32)
        n = self.lib.ora_fuzzy_true(self.h, q, len(q), terms.ctypes.data, dists.ctypes.data, cap)
        assert n <= cap
        return terms[:n].copy(), dists[:n].copy()

    def close(self):
        if self.h:
            self.lib.ora_index_free(self.h)
            self.h = None


def check_topk(got_ids, got_scores, all_ids, all_scores, limit, *, rtol=1e-5, exact_scores=False):
    """Tie-aware comparison of a top-k list with the full ground truth.

    SURVEY 8a F5 / north_star: ids exact, scores within 1e-5 relative, order
    identical except among ties inside the tolerance; the group cut by the
    limit may hold any of the documents sharing that score.
    """
    order = np.lexsort((-all_ids.astype(np.int64), -all_scores.astype(np.float64)))
    exp_n = min(limit, len(all_ids))
    assert len(got_ids) == exp_n, f"count {len(got_ids)} != {exp_n}"
    if exp_n == 0:
        return
    truth = dict(zip(all_ids.tolist(), all_scores.tolist()))
    assert len(set(got_ids.tolist())) == exp_n, "duplicate doc ids"
    # every returned doc exists, with the right score
    for d, s in zip(got_ids.tolist(), got_scores.tolist()):
        assert d in truth, f"doc {d} does not match the query"
        if exact_scores:
            assert s == truth[d], f"doc {d}: score {s!r} != {truth[d]!r}"
        else:
            assert abs(s - truth[d]) <= rtol * abs(truth[d]), f"doc {d}: score {s} vs {truth[d]}"
    # descending order
    gs = np.asarray(got_scores, dtype=np.float64)
    assert np.all(gs[:-1] >= gs[1:] - rtol * np.abs(gs[:-1])), "not sorted by score"
    # the k-th best true score bounds what may be returned
    kth = float(all_scores[order[exp_n - 1]])
    assert gs.min() >= kth - rtol * abs(kth), "a returned doc scores below the true k-th best"
    # and everything strictly above the cut (beyond tolerance) must be present
    must = all_ids[all_scores > kth + rtol * abs(kth)]
    missing = set(must.tolist()) - set(got_ids.tolist())
    assert not missing, f"missing better docs: {sorted(missing)[:5]}"
