"""Build libnxsearch.so (C11 host + sm_100a CUDA engine) in-tree.

The shared object is the drop-in for the reference's libnxsearch: gcc
compiles the C11 host sources, nvcc cross-compiles the CUDA engine for
sm_100a (no GPU needed at build time) and links both with the static CUDA
runtime, so the library loads on a box without a driver and fails loudly --
NXS_ERR_SYSTEM / NULL engine -- only when a search actually needs the GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = LIBDIR / "obj"
LIB = LIBDIR / "libnxsearch.so"

HOST_SRCS = [
    "hashmap.c", "json.c", "params.c", "results.c", "query.c", "tokenizer.c", "stem_en.c",
    "bkmirror.c", "index.c", "nxs.c", "search.c",
]
# libnxsb_tools.so: the synthetic-corpus generator, the bulk index-file writer
# and the query-front-end introspection the tests and benchmarks use
# (include/nxsb200_tools.h).  Its own library, so that the product library
# carries the search path only; it needs no CUDA.
TOOLS_LIB = LIBDIR / "libnxsb_tools.so"
TOOLS_SRCS = ["corpus.c", "querytools.c"]
TOOLS_SHARED = ["hashmap.c", "json.c", "params.c", "query.c", "tokenizer.c", "stem_en.c", "bkmirror.c"]
GPU_SRCS = ["engine.cu"]

CFLAGS = [
    "-std=gnu11", "-O2", "-g", "-fPIC", "-Wall", "-Wextra",
    "-fvisibility=hidden", "-D_GNU_SOURCE", f"-I{ROOT / 'include'}",
]
NVCCFLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
    f"-I{ROOT / 'include'}",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA engine cannot be built")


def _newer(src: Path, deps: list[Path], out: Path) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [src, *deps])


def _run(cmd: list[str]) -> None:
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"build failed: {' '.join(map(str, cmd))}\n{proc.stdout}\n{proc.stderr}")


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile what is out of date and return the path of libnxsearch.so."""
    OBJDIR.mkdir(parents=True, exist_ok=True)
    host_hdrs = list((CSRC / "host").glob("*.h")) + list((ROOT / "include").glob("*.h"))
    gpu_hdrs = list((CSRC / "gpu").glob("*.cuh")) + list((ROOT / "include").glob("*.h"))
    objs: list[Path] = []
    relink = force or not LIB.exists()

    for name in HOST_SRCS:
        src = CSRC / "host" / name
        obj = OBJDIR / (name + ".o")
        if force or _newer(src, host_hdrs, obj):
            if verbose:
                print("cc  ", name)
            _run(["gcc", *CFLAGS, "-c", str(src), "-o", str(obj)])
            relink = True
        objs.append(obj)
    for name in GPU_SRCS:
        src = CSRC / "gpu" / name
        obj = OBJDIR / (name + ".o")
        if force or _newer(src, gpu_hdrs, obj):
            if verbose:
                print("nvcc", name)
            _run([_nvcc(), *NVCCFLAGS, "-c", str(src), "-o", str(obj)])
            relink = True
        objs.append(obj)
    if relink:
        if verbose:
            print("link", LIB.name)
        _run([_nvcc(), "-shared", "-o", str(LIB), *map(str, objs),
              "-cudart", "static", "-lpthread", "-lm"])
    tools_objs: list[Path] = [OBJDIR / (n + ".o") for n in TOOLS_SHARED]
    relink_tools = force or relink or not TOOLS_LIB.exists()
    for name in TOOLS_SRCS:
        src = CSRC / "host" / name
        obj = OBJDIR / (name + ".o")
        if force or _newer(src, host_hdrs, obj):
            if verbose:
                print("cc  ", name)
            _run(["gcc", *CFLAGS, "-c", str(src), "-o", str(obj)])
            relink_tools = True
        tools_objs.append(obj)
    if relink_tools:
        if verbose:
            print("link", TOOLS_LIB.name)
        _run(["gcc", "-shared", "-Wl,--no-undefined", "-o", str(TOOLS_LIB), *map(str, tools_objs),
              "-lpthread", "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
