"""ctypes mirror of the public C API (include/nxs.h) -- same names, same
argument meaning and error behaviour as the reference's libnxsearch
(reference: src/core/nxs.h:26-101, docs/c-api.md)."""
from __future__ import annotations

import ctypes as C
import json

from ._lib import load_library

ERR_SUCCESS, ERR_FATAL, ERR_SYSTEM, ERR_INVALID, ERR_EXISTS, ERR_MISSING, ERR_LIMIT = range(7)


class NxsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code, self.msg = code, msg


def bind(lib=None):
    """Attach prototypes to a libnxsearch-compatible shared object."""
    lib = lib or load_library()
    vp, cp, u64, sz = C.c_void_p, C.c_char_p, C.c_uint64, C.c_size_t
    sig = {
        "nxs_open": (vp, [cp]), "nxs_close": (None, [vp]),
        "nxs_get_error": (C.c_int, [vp, C.POINTER(cp)]),
        "nxs_params_create": (vp, []), "nxs_params_fromjson": (vp, [vp, cp, sz]),
        "nxs_params_set_strlist": (C.c_int, [vp, cp, C.POINTER(cp), sz]),
        "nxs_params_set_str": (C.c_int, [vp, cp, cp]),
        "nxs_params_set_uint": (C.c_int, [vp, cp, u64]),
        "nxs_params_set_bool": (C.c_int, [vp, cp, C.c_bool]),
        "nxs_params_tojson": (vp, [vp, C.POINTER(sz)]),
        "nxs_params_release": (None, [vp]),
        "nxs_index_create": (vp, [vp, cp, vp]), "nxs_index_destroy": (C.c_int, [vp, cp]),
        "nxs_index_get_params": (vp, [vp]),
        "nxs_index_open": (vp, [vp, cp]), "nxs_index_close": (None, [vp]),
        "nxs_index_add": (C.c_int, [vp, vp, u64, cp, sz]),
        "nxs_index_remove": (C.c_int, [vp, u64]),
        "nxs_index_search": (vp, [vp, vp, cp, sz]),
        "nxs_resp_iter_reset": (None, [vp]),
        "nxs_resp_iter_result": (C.c_bool, [vp, C.POINTER(u64), C.POINTER(C.c_float)]),
        "nxs_resp_resultcount": (C.c_uint, [vp]),
        "nxs_resp_tojson": (vp, [vp, C.POINTER(sz)]),
        "nxs_resp_release": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if hasattr(lib, "nxs_index_search_batch"):
        lib.nxs_index_search_batch.restype = C.c_int
        lib.nxs_index_search_batch.argtypes = [vp, vp, C.POINTER(cp), sz, C.POINTER(vp)]
    if hasattr(lib, "nxs_index_search_batch_begin"):
        lib.nxs_index_search_batch_begin.restype = vp
        lib.nxs_index_search_batch_begin.argtypes = [vp, vp, C.POINTER(cp), sz]
        lib.nxs_index_search_batch_end.restype = C.c_int
        lib.nxs_index_search_batch_end.argtypes = [vp, C.POINTER(vp)]
    if hasattr(lib, "nxsb_resp_collect"):
        lib.nxsb_resp_collect.restype = u64
        lib.nxsb_resp_collect.argtypes = [C.POINTER(vp), sz, C.c_uint32, vp, vp, vp]
    if hasattr(lib, "nxs_luafilter_load"):
        lib.nxs_luafilter_load.restype = C.c_int
        lib.nxs_luafilter_load.argtypes = [vp, cp, cp]
    return lib


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def _take_string(ptr) -> str:
    s = C.string_at(ptr).decode()
    _libc.free(ptr)
    return s


class Params:
    def __init__(self, lib, **kw):
        self._lib = lib
        self.h = lib.nxs_params_create()
        for k, v in kw.items():
            self.set(k, v)

    def set(self, key: str, val) -> None:
        k = key.encode()
        if isinstance(val, bool):
            rc = self._lib.nxs_params_set_bool(self.h, k, val)
        elif isinstance(val, int):
            rc = self._lib.nxs_params_set_uint(self.h, k, val)
        elif isinstance(val, str):
            rc = self._lib.nxs_params_set_str(self.h, k, val.encode())
        else:
            arr = (C.c_char_p * len(val))(*[s.encode() for s in val])
            rc = self._lib.nxs_params_set_strlist(self.h, k, arr, len(val))
        if rc != 0:
            raise MemoryError(key)

    def tojson(self) -> str:
        return _take_string(self._lib.nxs_params_tojson(self.h, None))

    def release(self) -> None:
        if self.h:
            self._lib.nxs_params_release(self.h)
            self.h = None


class Response:
    def __init__(self, lib, h):
        self._lib, self.h = lib, h

    def results(self) -> list[tuple[int, float]]:
        out, d, s = [], C.c_uint64(), C.c_float()
        self._lib.nxs_resp_iter_reset(self.h)
        while self._lib.nxs_resp_iter_result(self.h, C.byref(d), C.byref(s)):
            out.append((d.value, s.value))
        return out

    @property
    def count(self) -> int:
        return self._lib.nxs_resp_resultcount(self.h)

    def tojson(self) -> str:
        return _take_string(self._lib.nxs_resp_tojson(self.h, None))

    def release(self) -> None:
        if self.h:
            self._lib.nxs_resp_release(self.h)
            self.h = None


class Index:
    def __init__(self, nxs: "Nxs", h):
        self.nxs, self._lib, self.h = nxs, nxs._lib, h

    def add(self, doc_id: int, text: str | bytes) -> None:
        b = text.encode() if isinstance(text, str) else text
        if self._lib.nxs_index_add(self.h, None, doc_id, b, len(b)) != 0:
            self.nxs.raise_error()

    def remove(self, doc_id: int) -> None:
        if self._lib.nxs_index_remove(self.h, doc_id) != 0:
            self.nxs.raise_error()

    def image_stats(self) -> dict:
        """nxsb_index_image_stats: how the HBM image got to its current state."""
        out = (C.c_uint64 * 7)()
        self._lib.nxsb_index_image_stats.argtypes = [C.c_void_p, C.c_void_p]
        self._lib.nxsb_index_image_stats.restype = None
        self._lib.nxsb_index_image_stats(self.h, out)
        keys = ("full_builds", "delta_builds", "consolidations", "segments", "dead_noted", "live", "pending")
        return dict(zip(keys, (int(x) for x in out)))

    def term_df(self, term_id: int) -> int:
        self._lib.nxsb_index_term_df.argtypes = [C.c_void_p, C.c_uint32]
        self._lib.nxsb_index_term_df.restype = C.c_uint32
        return int(self._lib.nxsb_index_term_df(self.h, term_id))

    def params_json(self) -> dict:
        p = self._lib.nxs_index_get_params(self.h)
        return json.loads(_take_string(self._lib.nxs_params_tojson(p, None)))

    def search(self, query: str | bytes, **params) -> list[tuple[int, float]]:
        r = self.search_resp(query, **params)
        try:
            return r.results()
        finally:
            r.release()

    def search_resp(self, query: str | bytes, **params) -> Response:
        q = query.encode() if isinstance(query, str) else query
        p = Params(self._lib, **params) if params else None
        h = self._lib.nxs_index_search(self.h, p.h if p else None, q, len(q))
        if p:
            p.release()
        if not h:
            self.nxs.raise_error()
        return Response(self._lib, h)

    def search_batch(self, queries, **params) -> list[list[tuple[int, float]] | None]:
        qs = [q.encode() if isinstance(q, str) else q for q in queries]
        arr = (C.c_char_p * len(qs))(*qs)
        out = (C.c_void_p * len(qs))()
        p = Params(self._lib, **params) if params else None
        rc = self._lib.nxs_index_search_batch(self.h, p.h if p else None, arr, len(qs), out)
        if p:
            p.release()
        if rc != 0:
            self.nxs.raise_error()
        res = []
        for h in out:
            if not h:
                res.append(None)
                continue
            r = Response(self._lib, h)
            res.append(r.results())
            r.release()
        return res

    def search_batch_arrays(self, queries, limit: int, **params):
        """nxs_index_search_batch, results drained in C through the public
        iterator (nxsb_resp_collect) into numpy arrays: (counts[n],
        ids[n, limit], scores[n, limit]).  `queries` may be a prepared
        (c_char_p * n) array."""
        import numpy as np

        if isinstance(queries, C.Array):
            arr, n = queries, len(queries)
        else:
            qs = [q.encode() if isinstance(q, str) else q for q in queries]
            arr, n = (C.c_char_p * len(qs))(*qs), len(qs)
        out = (C.c_void_p * n)()
        p = Params(self._lib, limit=limit, **params)
        rc = self._lib.nxs_index_search_batch(self.h, p.h, arr, n, out)
        p.release()
        if rc != 0:
            self.nxs.raise_error()
        counts = np.zeros(n, dtype=np.uint32)
        ids = np.zeros((n, limit), dtype=np.uint64)
        scores = np.zeros((n, limit), dtype=np.float32)
        self._lib.nxsb_resp_collect(out, n, limit, counts.ctypes.data, ids.ctypes.data, scores.ctypes.data)
        for h in out:
            if h:
                self._lib.nxs_resp_release(h)
        return counts, ids, scores

    def search_batch_begin(self, queries, limit: int, **params):
        """nxs_index_search_batch_begin: parse + resolve + submit, no wait.
        Returns a ticket for search_batch_end_arrays()."""
        if isinstance(queries, C.Array):
            arr, n = queries, len(queries)
        else:
            qs = [q.encode() if isinstance(q, str) else q for q in queries]
            arr, n = (C.c_char_p * len(qs))(*qs), len(qs)
        p = Params(self._lib, limit=limit, **params)
        bt = self._lib.nxs_index_search_batch_begin(self.h, p.h, arr, n)
        p.release()
        if not bt:
            self.nxs.raise_error()
        return (bt, n, limit, arr)

    def search_batch_end_arrays(self, ticket):
        """nxs_index_search_batch_end + nxsb_resp_collect -> (counts, ids, scores)."""
        import numpy as np

        bt, n, limit, _ = ticket
        out = (C.c_void_p * n)()
        if self._lib.nxs_index_search_batch_end(bt, out) != 0:
            self.nxs.raise_error()
        counts = np.zeros(n, dtype=np.uint32)
        ids = np.zeros((n, limit), dtype=np.uint64)
        scores = np.zeros((n, limit), dtype=np.float32)
        self._lib.nxsb_resp_collect(out, n, limit, counts.ctypes.data, ids.ctypes.data, scores.ctypes.data)
        for h in out:
            if h:
                self._lib.nxs_resp_release(h)
        return counts, ids, scores

    def close(self) -> None:
        if self.h:
            self._lib.nxs_index_close(self.h)
            self.h = None


class Nxs:
    """One library instance (reference: nxs_t)."""

    def __init__(self, basedir: str, lib=None):
        self._lib = bind(lib)
        self.h = self._lib.nxs_open(str(basedir).encode())
        if not self.h:
            raise OSError(f"nxs_open({basedir}) failed")

    def error(self) -> tuple[int, str | None]:
        msg = C.c_char_p()
        code = self._lib.nxs_get_error(self.h, C.byref(msg))
        return code, (msg.value.decode() if msg.value else None)

    def raise_error(self):
        code, msg = self.error()
        raise NxsError(code, msg or "unknown error")

    def create_index(self, name: str, **params) -> Index:
        p = Params(self._lib, **params) if params else None
        h = self._lib.nxs_index_create(self.h, name.encode(), p.h if p else None)
        if p:
            p.release()
        if not h:
            self.raise_error()
        return Index(self, h)

    def open_index(self, name: str) -> Index:
        h = self._lib.nxs_index_open(self.h, name.encode())
        if not h:
            self.raise_error()
        return Index(self, h)

    def destroy_index(self, name: str) -> None:
        if self._lib.nxs_index_destroy(self.h, name.encode()) != 0:
            self.raise_error()

    def close(self) -> None:
        if self.h:
            self._lib.nxs_close(self.h)
            self.h = None
