"""Build and run tests/c/nxs_caller.c: a plain C program compiled against
include/nxs.h and linked with libnxsearch.so, the way a user of the reference
links theirs (ref src/utils/benchmark.c).  Used by the GPU tests (results
mode) and by bench.py's single-query latency leg."""
from __future__ import annotations

import json
import os
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "c" / "nxs_caller.c"
OUT = ROOT / "tests" / "c" / "_build" / "nxs_caller"


def build() -> Path:
    from nxsearch_b200._lib import load_library  # noqa: F401  (builds libnxsearch.so if needed)
    libdir = ROOT / "nxsearch_b200" / "lib"
    if OUT.exists() and OUT.stat().st_mtime >= max(SRC.stat().st_mtime, (ROOT / "include" / "nxs.h").stat().st_mtime):
        return OUT
    OUT.parent.mkdir(parents=True, exist_ok=True)
    subprocess.run(["gcc", "-std=gnu11", "-O2", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(SRC),
                    "-o", str(OUT), f"-L{libdir}", "-lnxsearch", f"-Wl,-rpath,{libdir}"], check=True)
    return OUT


def run(basedir: str, index: str, algo: str, limit: int, queries: list[str], mode: str = "results",
        warmup: int = 16, device: int | None = None) -> str:
    exe = build()
    qfile = Path(basedir) / f"queries_{os.getpid()}.txt"
    qfile.write_text("".join(q + "\n" for q in queries))
    env = dict(os.environ)
    if device is not None:
        env["NXS_GPU_DEVICE"] = str(device)
    try:
        cmd = [str(exe), basedir, index, algo, str(limit), str(qfile), mode] + ([str(warmup)] if mode == "latency" else [])
        return subprocess.run(cmd, check=True, capture_output=True, text=True, env=env).stdout
    finally:
        qfile.unlink(missing_ok=True)


def results(basedir, index, algo, limit, queries, **kw) -> list[list[tuple[int, float]]]:
    out = []
    for line in run(basedir, index, algo, limit, queries, "results", **kw).splitlines():
        parts = line.split()
        assert parts and parts[0] != "error", line
        out.append([(int(p.split(":")[0]), float(p.split(":")[1])) for p in parts[1:]])
        assert len(out[-1]) == int(parts[0])
    return out


def latency(basedir, index, algo, limit, queries, **kw) -> dict:
    return json.loads(run(basedir, index, algo, limit, queries, "latency", **kw))
