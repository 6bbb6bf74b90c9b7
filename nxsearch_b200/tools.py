"""ctypes mirror of include/nxsb200_tools.h: synthetic corpus and index files."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import load_tools_library as load_library

SEED = 0x6E78735F42323030  # "nxs_B200", SURVEY section 8(d)


class CorpusStruct(C.Structure):
    _fields_ = [
        ("n_docs", C.c_uint32), ("n_terms", C.c_uint32), ("n_pairs", C.c_uint64),
        ("token_count", C.c_uint64), ("doc_count", C.c_uint32),
        ("doc_ids", C.POINTER(C.c_uint64)), ("doc_len", C.POINTER(C.c_uint32)),
        ("doc_off", C.POINTER(C.c_uint64)), ("pairs", C.POINTER(C.c_uint32)),
        ("term_blob", C.c_void_p), ("term_off", C.POINTER(C.c_uint32)),
        ("term_total", C.POINTER(C.c_uint64)), ("term_df", C.POINTER(C.c_uint32)),
    ]


def _bind():
    lib = load_library()
    P = C.POINTER(CorpusStruct)
    lib.nxsb_corpus_generate.restype = P
    lib.nxsb_corpus_generate.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int, C.c_int]
    lib.nxsb_corpus_free.argtypes = [P]
    lib.nxsb_corpus_free.restype = None
    lib.nxsb_corpus_query_terms.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.nxsb_corpus_query_terms.restype = None
    lib.nxsb_corpus_fuzzy_terms.argtypes = [C.c_uint64, P, C.c_void_p, C.c_size_t, C.c_size_t]
    lib.nxsb_corpus_fuzzy_terms.restype = None
    lib.nxsb_write_terms_file.argtypes = [C.c_char_p, P]
    lib.nxsb_write_dtmap_file.argtypes = [C.c_char_p, P]
    lib.nxsb_read_index_files.restype = P
    lib.nxsb_read_index_files.argtypes = [C.c_char_p, C.c_char_p]
    lib.nxsb_bkmirror_build.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


class Corpus:
    """A document-major corpus owned by the C library (numpy views, no copies)."""

    def __init__(self, ptr):
        if not ptr:
            raise MemoryError("corpus allocation / parsing failed")
        self._lib = _bind()
        self.ptr = ptr
        c = ptr.contents
        self.n_docs, self.n_terms, self.n_pairs = c.n_docs, c.n_terms, c.n_pairs
        self.token_count, self.doc_count = c.token_count, c.doc_count
        as_np = np.ctypeslib.as_array
        self.doc_ids = as_np(c.doc_ids, (max(self.n_docs, 1),))[: self.n_docs]
        self.doc_len = as_np(c.doc_len, (max(self.n_docs, 1),))[: self.n_docs]
        self.doc_off = as_np(c.doc_off, (self.n_docs + 1,))
        self.pairs = as_np(c.pairs, (max(2 * self.n_pairs, 1),))[: 2 * self.n_pairs]
        self.term_off = as_np(c.term_off, (self.n_terms + 1,))
        self.term_total = as_np(c.term_total, (max(self.n_terms, 1),))[: self.n_terms]
        self.term_df = as_np(c.term_df, (max(self.n_terms, 1),))[: self.n_terms]
        self.term_blob = C.string_at(c.term_blob, int(self.term_off[-1]))

    @classmethod
    def generate(cls, n_docs: int, n_terms: int, *, seed: int = SEED, first_doc: int = 0,
                 sparse_ids: bool = False, nthreads: int = 0) -> "Corpus":
        return cls(_bind().nxsb_corpus_generate(seed, n_terms, first_doc, n_docs, int(sparse_ids), nthreads))

    @classmethod
    def read(cls, terms_path: str, dtmap_path: str) -> "Corpus":
        return cls(_bind().nxsb_read_index_files(str(terms_path).encode(), str(dtmap_path).encode()))

    def term(self, term_id: int) -> str:
        """The string of 1-based term id."""
        return self.term_blob[self.term_off[term_id - 1]: self.term_off[term_id]].decode()

    def write(self, terms_path: str, dtmap_path: str) -> None:
        if self._lib.nxsb_write_terms_file(str(terms_path).encode(), self.ptr) != 0:
            raise OSError(f"cannot write {terms_path}")
        if self._lib.nxsb_write_dtmap_file(str(dtmap_path).encode(), self.ptr) != 0:
            raise OSError(f"cannot write {dtmap_path}")

    def query_terms(self, n: int, *, seed: int = SEED + 1) -> np.ndarray:
        """n Zipf(1.0) term ids restricted to df >= 1 (SURVEY 8d "Queries")."""
        out = np.zeros(n, dtype=np.uint32)
        df = np.ascontiguousarray(self.term_df, dtype=np.uint32)
        self._lib.nxsb_corpus_query_terms(seed, self.n_terms, df.ctypes.data, out.ctypes.data, n)
        return out

    def fuzzy_terms(self, n: int, *, seed: int = SEED + 2, stride: int = 32) -> list[bytes]:
        buf = C.create_string_buffer(n * stride)
        self._lib.nxsb_corpus_fuzzy_terms(seed, self.ptr, buf, stride, n)
        raw = buf.raw
        return [raw[i * stride:(i + 1) * stride].split(b"\0", 1)[0] for i in range(n)]

    def bk_mirror(self):
        """(parent, edge, rank) arrays of the reference's BK-tree over this vocabulary."""
        n = self.n_terms
        parent = np.zeros(n, dtype=np.uint32)
        edge = np.zeros(n, dtype=np.uint8)
        rank = np.zeros(n, dtype=np.uint32)
        c = self.ptr.contents
        if self._lib.nxsb_bkmirror_build(c.term_blob, C.cast(c.term_off, C.c_void_p), n,
                                         parent.ctypes.data, edge.ctypes.data, rank.ctypes.data) != 0:
            raise MemoryError("BK mirror")
        return parent, edge, rank

    def close(self) -> None:
        if self.ptr:
            self._lib.nxsb_corpus_free(self.ptr)
            self.ptr = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass


_KINDS = {1: "OR", 2: "AND", 3: "NOT", 4: "(", 5: ")", 6: "FF", 7: "QS"}
_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def query_lex(query: str) -> list[str]:
    lib = _bind()
    lib.nxsb_query_lex.restype = C.c_size_t
    lib.nxsb_query_lex.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
    kinds = (C.c_int * 256)()
    n = lib.nxsb_query_lex(query.encode(), kinds, 256)
    return [_KINDS[kinds[i]] for i in range(min(n, 256))]


def query_dump(query: str) -> tuple[str | None, str | None]:
    """(s-expression, None) or (None, syntax error message)."""
    lib = _bind()
    lib.nxsb_query_dump.restype = C.c_void_p
    lib.nxsb_query_dump.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    err = C.c_void_p()
    p = lib.nxsb_query_dump(query.encode(), C.byref(err))
    out = msg = None
    if p:
        out = C.string_at(p).decode()
        _libc.free(p)
    if err.value:
        msg = C.string_at(err.value).decode()
        _libc.free(err.value)
    return out, msg


def query_compile(query: str) -> tuple[list[str], list[int]] | None:
    """(leaf strings in token-list order, postfix program) or None on a syntax error."""
    lib = _bind()
    lib.nxsb_query_compile.restype = C.c_int
    lib.nxsb_query_compile.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32),
                                       C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    buf = C.create_string_buffer(1 << 16)
    prog = (C.c_int32 * 1024)()
    nt, npg = C.c_uint32(), C.c_uint32()
    if lib.nxsb_query_compile(query.encode(), buf, len(buf), C.byref(nt), prog, 1024, C.byref(npg)) != 0:
        return None
    toks = buf.raw.split(b"\0")[: nt.value]
    return [t.decode() for t in toks], [prog[i] for i in range(npg.value)]


def tokenize(text: str | bytes, *, normalize: bool = True, stem: bool = False) -> list[tuple[str, int]]:
    """The text front end on its own: [(word, occurrences)] in first-seen order
    (stem=True: the English stemmer after the normalizer)."""
    lib = _bind()
    lib.nxsb_tokenize.restype = C.c_int
    lib.nxsb_tokenize.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                  C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32]
    raw = text.encode() if isinstance(text, str) else text
    buf = C.create_string_buffer(2 * len(raw) + 16)
    counts = (C.c_uint32 * (len(raw) + 1))()
    n = C.c_uint32()
    if lib.nxsb_tokenize(raw, len(raw), int(normalize) | (2 if stem else 0), buf, len(buf), C.byref(n), counts, len(raw) + 1) != 0:
        raise RuntimeError("nxsb_tokenize failed")
    toks = buf.raw.split(b"\0")[: n.value]
    return [(t.decode(errors="replace"), counts[i]) for i, t in enumerate(toks)]


def query_terms(n_terms: int, df, n: int, *, seed: int = SEED + 1) -> np.ndarray:
    """n Zipf(1.0) term ids over a vocabulary of n_terms, restricted to df >= 1.

    Deterministic in (seed, n_terms, df): every rank of a sharded run draws the
    same query stream from the all-reduced df."""
    out = np.zeros(n, dtype=np.uint32)
    dfa = np.ascontiguousarray(df, dtype=np.uint32)
    _bind().nxsb_corpus_query_terms(seed, n_terms, dfa.ctypes.data, out.ctypes.data, n)
    return out
