"""ctypes mirror of include/nxsb200_gpu.h: the CUDA engine's C ABI."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import load_library

ALGO_TFIDF, ALGO_BM25 = 0, 1
OP_EMPTY, OP_AND, OP_OR, OP_ANDNOT = -1, -2, -3, -4
TILE_DOCS = 16384


class ShardDesc(C.Structure):
    _fields_ = [
        ("n_docs", C.c_uint32), ("n_terms", C.c_uint32),
        ("doc_ids", C.c_void_p), ("doc_len", C.c_void_p), ("doc_off", C.c_void_p),
        ("pairs", C.c_void_p), ("token_count", C.c_uint64), ("doc_count", C.c_uint32),
        ("df", C.c_void_p),
        ("raw", C.c_void_p), ("raw_off", C.c_void_p), ("raw_n", C.c_void_p),
    ]


class QueryDesc(C.Structure):
    _fields_ = [("tok_off", C.c_uint32), ("n_tokens", C.c_uint32),
                ("prog_off", C.c_uint32), ("n_prog", C.c_uint32)]


class BatchDesc(C.Structure):
    _fields_ = [
        ("algo", C.c_int), ("limit", C.c_uint32), ("n_queries", C.c_uint32),
        ("queries", C.c_void_p), ("tokens", C.c_void_p), ("n_tokens", C.c_uint32),
        ("prog", C.c_void_p), ("n_prog", C.c_uint32),
    ]


def _bind():
    lib = load_library()
    vp, u32, u64, i = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    sig = {
        "nxsb_gpu_device_count": (i, []),
        "nxsb_last_error": (C.c_char_p, []),
        "nxsb_engine_create": (vp, [i]),
        "nxsb_engine_create_replicated": (vp, [C.POINTER(C.c_int), i]),
        "nxsb_engine_create_sharded": (vp, [C.POINTER(C.c_int), i]),
        "nxsb_engine_replica_count": (i, [vp]),
        "nxsb_engine_is_sharded": (i, [vp]),
        "nxsb_engine_destroy": (None, [vp]),
        "nxsb_engine_errmsg": (C.c_char_p, [vp]),
        "nxsb_engine_set_stream": (i, [vp, vp]),
        "nxsb_engine_lane_stream": (vp, [vp, i]),
        "nxsb_engine_lanes_join": (i, [vp]),
        "nxsb_engine_load_shard": (i, [vp, C.POINTER(ShardDesc)]),
        "nxsb_engine_segment_add": (i, [vp, C.POINTER(ShardDesc)]),
        "nxsb_engine_segment_count": (i, [vp]),
        "nxsb_engine_segments_drop": (i, [vp]),
        "nxsb_engine_set_dead": (i, [vp, u32, vp, u32]),
        "nxsb_engine_get_df": (i, [vp, vp, u32]),
        "nxsb_engine_set_global_stats": (i, [vp, vp, u32, u64, u32]),
        "nxsb_engine_search": (i, [vp, C.POINTER(BatchDesc), vp, vp, vp]),
        "nxsb_engine_batch_upload": (i, [vp, C.POINTER(BatchDesc)]),
        "nxsb_engine_batch_run": (i, [vp, i, vp]),
        "nxsb_engine_batch_fetch": (i, [vp, i, vp, vp, vp]),
        "nxsb_engine_batch_release": (i, [vp, i]),
        "nxsb_engine_batch_bytes": (u64, [vp, i]),
        "nxsb_engine_sync": (i, [vp]),
        "nxsb_engine_merge_topk": (i, [vp, vp, u32, u32, u32, vp]),
        "nxsb_engine_load_vocab": (i, [vp, u32, vp, vp, vp, vp, vp, vp]),
        "nxsb_engine_search_begin": (i, [vp, vp]),
        "nxsb_engine_search_begin_dev": (i, [vp, vp, vp]),
        "nxsb_engine_search_end": (i, [vp, i, vp, vp, vp]),
        "nxsb_engine_fuzzy": (i, [vp, u32, vp, vp, vp, vp, vp]),
        "nxsb_engine_fuzzy_candidates": (i, [vp, u32, vp, vp, u32, vp, vp, vp, vp, vp, vp]),
        "nxsb_engine_last_timings": (i, [vp, vp, vp, i]),
        "nxsb_engine_timings": (i, [vp, u32, vp, vp, i]),
        "nxsb_engine_launch_count": (u64, [vp]),
        "nxsb_alloc_events": (u64, []),
        "nxsb_engine_set_pruning": (i, [vp, i]),
        "nxsb_engine_pruning_stats": (i, [vp, vp, i]),
        "nxsb_engine_term_kth": (i, [vp, i, vp, C.c_uint32, vp]),
        "nxsb_engine_score_pairs": (i, [vp, i, u32, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def device_count() -> int:
    return _bind().nxsb_gpu_device_count()


def alloc_events() -> int:
    """Device / pinned allocations and frees made by the library so far."""
    return _bind().nxsb_alloc_events()


@dataclass
class Batch:
    """Host-side batch: queries as (token term ids, postfix program) pairs."""
    algo: int
    limit: int
    queries: np.ndarray  # structured (n, 4) uint32: tok_off, n_tokens, prog_off, n_prog
    tokens: np.ndarray   # uint32
    prog: np.ndarray     # int32

    @classmethod
    def from_lists(cls, algo: int, limit: int, items) -> "Batch":
        """items: iterable of (token_term_ids, program) with program=None => OR of all."""
        q, toks, prog = [], [], []
        for tk, pr in items:
            tk = list(tk)
            if pr is None:
                pr = []
                for s in range(len(tk)):
                    pr.append(s)
                    if s:
                        pr.append(OP_OR)
            q.append((len(toks), len(tk), len(prog), len(pr)))
            toks.extend(tk)
            prog.extend(pr)
        return cls(algo, limit,
                   np.array(q, dtype=np.uint32).reshape(-1, 4),
                   np.array(toks, dtype=np.uint32), np.array(prog, dtype=np.int32))

    def desc(self) -> BatchDesc:
        return BatchDesc(self.algo, self.limit, len(self.queries),
                         self.queries.ctypes.data, self.tokens.ctypes.data, len(self.tokens),
                         self.prog.ctypes.data, len(self.prog))


class Engine:
    """One CUDA device / one document shard.  Raises if no GPU is usable."""

    def __init__(self, device: int = 0, *, devices=None, layout: str = "replicas"):
        """devices=[...]: one engine over several devices -- layout "replicas" (the whole
        image on each, a batch's queries split) or "shards" (a range of the documents on
        each, every query scored everywhere, lists merged on the first device)."""
        self._lib = _bind()
        if devices is None:
            self._h = self._lib.nxsb_engine_create(device)
        else:
            if layout not in ("replicas", "shards"):
                raise ValueError(f"unknown layout {layout!r}")
            arr = (C.c_int * len(devices))(*devices)
            make = self._lib.nxsb_engine_create_sharded if layout == "shards" \
                else self._lib.nxsb_engine_create_replicated
            self._h = make(arr, len(devices))
        if not self._h:
            raise RuntimeError("nxsb_engine_create failed: " + self._lib.nxsb_last_error().decode())

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise RuntimeError(self._lib.nxsb_engine_errmsg(self._h).decode())
        return rc

    def close(self) -> None:
        if self._h:
            self._lib.nxsb_engine_destroy(self._h)
            self._h = None

    def set_stream(self, cuda_stream: int | None) -> None:
        self._check(self._lib.nxsb_engine_set_stream(self._h, cuda_stream))

    def lane_stream(self, lane: int = 0) -> int:
        """cudaStream_t (as an integer) of one of the engine's own streams."""
        return int(self._lib.nxsb_engine_lane_stream(self._h, lane) or 0)

    def lanes_join(self) -> None:
        """Lane 0 waits for what the other lanes have been given so far."""
        self._check(self._lib.nxsb_engine_lanes_join(self._h))

    def load_corpus(self, corpus, *, lo: int = 0, hi: int | None = None, df=None,
                    token_count: int | None = None, doc_count: int | None = None,
                    segment: bool = False) -> None:
        """Load documents [lo, hi) of a tools.Corpus (ids must ascend) as this shard,
        or -- segment=True -- as a delta segment next to the image already loaded."""
        hi = corpus.n_docs if hi is None else hi
        doc_off = corpus.doc_off[lo:hi + 1] - corpus.doc_off[lo]
        base = int(corpus.doc_off[lo])
        pairs = corpus.pairs[2 * base: 2 * int(corpus.doc_off[hi])]
        self.load_docs(corpus.doc_ids[lo:hi], corpus.doc_len[lo:hi], doc_off, pairs, corpus.n_terms,
                       corpus.token_count if token_count is None else token_count,
                       corpus.doc_count if doc_count is None else doc_count, df, segment=segment)

    def load_docs(self, ids, lens, doc_off, pairs, n_terms: int, token_count: int, doc_count: int,
                  df=None, *, segment: bool = False) -> None:
        """Load a document-major shard given as arrays (nxsb_shard_desc_t); ids ascending."""
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        doc_off = np.ascontiguousarray(doc_off, dtype=np.uint64)
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32)
        dfa = None if df is None else np.ascontiguousarray(df, dtype=np.uint32)
        sd = ShardDesc(len(ids), n_terms, ids.ctypes.data, lens.ctypes.data,
                       doc_off.ctypes.data, pairs.ctypes.data, token_count, doc_count,
                       None if dfa is None else dfa.ctypes.data, None, None, None)
        if segment:
            self._check(self._lib.nxsb_engine_segment_add(self._h, C.byref(sd)))
        else:
            self._check(self._lib.nxsb_engine_load_shard(self._h, C.byref(sd)))

    def load_dtmap(self, raw: np.ndarray, ids, lens, raw_off, raw_n, n_terms: int, token_count: int,
                   doc_count: int, df=None, *, segment: bool = False) -> None:
        """Load a shard straight from `nxsdtmap` bytes (nxsb_shard_desc_t.raw): document i's
        raw_n[i] big-endian (term, count) pairs start at raw[raw_off[i]]; decoded on the device."""
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        raw_off = np.ascontiguousarray(raw_off, dtype=np.uint64)
        raw_n = np.ascontiguousarray(raw_n, dtype=np.uint32)
        dfa = None if df is None else np.ascontiguousarray(df, dtype=np.uint32)
        sd = ShardDesc(len(ids), n_terms, ids.ctypes.data, lens.ctypes.data, None, None,
                       token_count, doc_count, None if dfa is None else dfa.ctypes.data,
                       raw.ctypes.data, raw_off.ctypes.data, raw_n.ctypes.data)
        if segment:
            self._check(self._lib.nxsb_engine_segment_add(self._h, C.byref(sd)))
        else:
            self._check(self._lib.nxsb_engine_load_shard(self._h, C.byref(sd)))

    def segment_count(self) -> int:
        return self._lib.nxsb_engine_segment_count(self._h)

    def segments_drop(self) -> None:
        self._check(self._lib.nxsb_engine_segments_drop(self._h))

    def set_dead(self, segment: int, ids) -> None:
        """Ids removed from a segment (0 = base) after it was built."""
        ids = np.ascontiguousarray(sorted(int(x) for x in ids), dtype=np.uint64)
        self._check(self._lib.nxsb_engine_set_dead(self._h, segment, ids.ctypes.data, len(ids)))

    def get_df(self, n_terms: int) -> np.ndarray:
        df = np.zeros(n_terms, dtype=np.uint32)
        self._check(self._lib.nxsb_engine_get_df(self._h, df.ctypes.data, n_terms))
        return df

    def set_global_stats(self, df: np.ndarray, token_count: int, doc_count: int) -> None:
        df = np.ascontiguousarray(df, dtype=np.uint32)
        self._check(self._lib.nxsb_engine_set_global_stats(self._h, df.ctypes.data, len(df), token_count, doc_count))

    def search(self, batch: Batch):
        n, k = len(batch.queries), batch.limit
        counts = np.zeros(n, dtype=np.uint32)
        ids = np.zeros(max(n * k, 1), dtype=np.uint64)
        scores = np.zeros(max(n * k, 1), dtype=np.float32)
        d = batch.desc()
        self._check(self._lib.nxsb_engine_search(self._h, C.byref(d), counts.ctypes.data,
                                                 ids.ctypes.data, scores.ctypes.data))
        return counts, ids[: n * k].reshape(n, k), scores[: n * k].reshape(n, k)

    def search_begin(self, batch: Batch, d_recs: int | None = None) -> int:
        """Submit a batch without waiting (descriptor copy, kernels and -- unless
        d_recs names a device buffer for the records -- the result copy)."""
        d = batch.desc()
        if d_recs is None:
            return self._check(self._lib.nxsb_engine_search_begin(self._h, C.byref(d)))
        return self._check(self._lib.nxsb_engine_search_begin_dev(self._h, C.byref(d), d_recs))

    def search_end(self, handle: int, n: int = 0, k: int = 0, *, discard: bool = False):
        """Wait for a submitted batch; returns (counts, ids, scores) unless discarded."""
        if discard:
            self._check(self._lib.nxsb_engine_search_end(self._h, handle, None, None, None))
            return None
        counts = np.zeros(n, dtype=np.uint32)
        ids = np.zeros(max(n * k, 1), dtype=np.uint64)
        scores = np.zeros(max(n * k, 1), dtype=np.float32)
        self._check(self._lib.nxsb_engine_search_end(self._h, handle, counts.ctypes.data,
                                                     ids.ctypes.data, scores.ctypes.data))
        return counts, ids[: n * k].reshape(n, k), scores[: n * k].reshape(n, k)

    def upload(self, batch: Batch) -> int:
        d = batch.desc()
        return self._check(self._lib.nxsb_engine_batch_upload(self._h, C.byref(d)))

    def run(self, handle: int, d_recs: int | None = None) -> None:
        self._check(self._lib.nxsb_engine_batch_run(self._h, handle, d_recs))

    def fetch(self, handle: int, n: int, k: int):
        counts = np.zeros(n, dtype=np.uint32)
        ids = np.zeros(max(n * k, 1), dtype=np.uint64)
        scores = np.zeros(max(n * k, 1), dtype=np.float32)
        self._check(self._lib.nxsb_engine_batch_fetch(self._h, handle, counts.ctypes.data,
                                                      ids.ctypes.data, scores.ctypes.data))
        return counts, ids[: n * k].reshape(n, k), scores[: n * k].reshape(n, k)

    def release(self, handle: int) -> None:
        self._check(self._lib.nxsb_engine_batch_release(self._h, handle))

    def batch_bytes(self, handle: int) -> int:
        return self._lib.nxsb_engine_batch_bytes(self._h, handle)

    def sync(self) -> None:
        self._check(self._lib.nxsb_engine_sync(self._h))

    def merge_topk(self, d_in: int, n_shards: int, n_queries: int, limit: int, d_out: int) -> None:
        self._check(self._lib.nxsb_engine_merge_topk(self._h, d_in, n_shards, n_queries, limit, d_out))

    def load_vocab(self, blob: bytes, term_off, term_total, parent, edge, rank) -> None:
        term_off = np.ascontiguousarray(term_off, dtype=np.uint32)
        n = len(term_off) - 1
        tt = np.ascontiguousarray(term_total, dtype=np.uint64)
        pa = np.ascontiguousarray(parent, dtype=np.uint32)
        ed = np.ascontiguousarray(edge, dtype=np.uint8)
        rk = np.ascontiguousarray(rank, dtype=np.uint32)
        buf = C.create_string_buffer(blob, len(blob) + 1)
        self._check(self._lib.nxsb_engine_load_vocab(self._h, n, C.cast(buf, C.c_void_p), term_off.ctypes.data,
                                                     tt.ctypes.data, pa.ctypes.data, ed.ctypes.data, rk.ctypes.data))

    def fuzzy(self, queries: list[bytes], want_true: bool = False):
        n = len(queries)
        off = np.zeros(n + 1, dtype=np.uint32)
        off[1:] = np.cumsum([len(q) for q in queries])
        blob = b"".join(queries)
        buf = C.create_string_buffer(blob, len(blob) + 1)
        term = np.zeros(n, dtype=np.uint32)
        dist = np.zeros(n, dtype=np.uint32)
        true = np.zeros(n, dtype=np.uint32) if want_true else None
        self._check(self._lib.nxsb_engine_fuzzy(self._h, n, C.cast(buf, C.c_void_p), off.ctypes.data,
                                                term.ctypes.data, dist.ctypes.data,
                                                None if true is None else true.ctypes.data))
        return term, dist, true

    def fuzzy_candidates(self, queries: list[bytes], cap: int = 1024):
        """Per query: (chosen term, its distance, number of terms within
        distance 2, [(term id, distance, reached, live)] in BFS order) -- the
        reached entries are the reference's bktree_search list."""
        n = len(queries)
        off = np.zeros(n + 1, dtype=np.uint32)
        off[1:] = np.cumsum([len(q) for q in queries])
        blob = b"".join(queries)
        buf = C.create_string_buffer(blob, len(blob) + 1)
        term = np.zeros(n, dtype=np.uint32)
        dist = np.zeros(n, dtype=np.uint32)
        cnt = np.zeros(n, dtype=np.uint32)
        ct = np.zeros((n, cap), dtype=np.uint32)
        cd = np.zeros((n, cap), dtype=np.uint8)
        cf = np.zeros((n, cap), dtype=np.uint8)
        self._check(self._lib.nxsb_engine_fuzzy_candidates(
            self._h, n, C.cast(buf, C.c_void_p), off.ctypes.data, cap, term.ctypes.data, dist.ctypes.data,
            cnt.ctypes.data, ct.ctypes.data, cd.ctypes.data, cf.ctypes.data))
        return term, dist, cnt, ct, cd, cf

    def timings(self, last_runs: int = 1) -> dict[str, float]:
        """Device milliseconds per kernel family, summed over the last runs
        (CUDA events on the engine's stream; synchronises the stream)."""
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        n = self._lib.nxsb_engine_timings(self._h, last_runs, names, ms, 16)
        return {names[i].decode(): float(ms[i]) for i in range(n)}

    def last_timings(self) -> dict[str, float]:
        return self.timings(1)

    @property
    def launches(self) -> int:
        return self._lib.nxsb_engine_launch_count(self._h)

    def set_pruning(self, on: bool) -> bool:
        """Exact block-max pruning of OR queries on/off (batches staged afterwards)."""
        return bool(self._lib.nxsb_engine_set_pruning(self._h, int(on)))

    KTH_STEPS = (1, 2, 4, 10, 20, 50, 100, 128)

    def term_kth(self, algo: int, term_ids) -> np.ndarray:
        """[n, 8] k-th largest weight of each term for k in KTH_STEPS (0: fewer postings)."""
        ids = np.ascontiguousarray(term_ids, dtype=np.uint32)
        out = np.zeros((len(ids), len(self.KTH_STEPS)), dtype=np.float32)
        self._check(self._lib.nxsb_engine_term_kth(self._h, algo, ids.ctypes.data, len(ids), out.ctypes.data))
        return out

    def score_pairs(self, algo: int, tf, dl, idf) -> np.ndarray:
        """Scores of (tf, dl, idf) triples through the kernels' arithmetic (diagnostic)."""
        tf = np.ascontiguousarray(tf, dtype=np.uint32)
        dl = np.ascontiguousarray(dl, dtype=np.uint32)
        idf = np.ascontiguousarray(idf, dtype=np.float32)
        out = np.zeros(len(tf), dtype=np.float32)
        self._check(self._lib.nxsb_engine_score_pairs(self._h, algo, len(tf), tf.ctypes.data, dl.ctypes.data,
                                                      idf.ctypes.data, out.ctypes.data))
        return out

    def pruning_stats(self, reset: bool = False) -> dict[str, int]:
        out = (C.c_uint64 * 16)()
        self._check(self._lib.nxsb_engine_pruning_stats(self._h, out, int(reset)))
        st = {"items": out[0], "blocks_scored": out[1], "postings_scored": out[2], "rounds": out[3]}
        if any(out[4:12]):      # -DBMW_PROF build: cycles of thread 0 per phase
            names = ["bound_select", "bounds_short_lists", "pick", "score", "cut", "emit_merge", "item_setup"]
            st["phase_cycles"] = {n: out[4 + i] for i, n in enumerate(names)}
        return st

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass
