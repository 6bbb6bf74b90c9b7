"""Locate and load libnxsearch.so (never silently substitutes anything)."""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

_LIB = None


def library_path() -> Path:
    # NXSB_LIBRARY: development switch for A/B-ing kernel build variants.
    override = os.environ.get("NXSB_LIBRARY")
    if override:
        return Path(override)
    return Path(__file__).resolve().parent / "lib" / "libnxsearch.so"


def load_library() -> ctypes.CDLL:
    """Load the in-tree shared object; raise if it has not been built."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not path.exists():
            raise RuntimeError(
                f"{path} is missing: build it with `python -m nxsearch_b200._build` "
                "(there is no pure-Python or CPU fallback)")
        _LIB = ctypes.CDLL(str(path))
    return _LIB


_TOOLS = None


def load_tools_library() -> ctypes.CDLL:
    """libnxsb_tools.so: corpus generator, index-file writer, query introspection
    (test and benchmark tooling, kept out of the product library)."""
    global _TOOLS
    if _TOOLS is None:
        path = Path(__file__).resolve().parent / "lib" / "libnxsb_tools.so"
        if not path.exists():
            raise RuntimeError(f"{path} is missing: build it with `python -m nxsearch_b200._build`")
        _TOOLS = ctypes.CDLL(str(path))
    return _TOOLS
