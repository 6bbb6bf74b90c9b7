#!/usr/bin/env python
"""Incremental image refresh at the headline size: what the first
search after nxs_index_add / nxs_index_remove costs with delta segments, against
the full rebuild it replaces, and what a segmented image costs per batch.

    python scripts/refresh_probe.py [--docs 10000000] [--out profiles/x.json]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def doc_text(corpus, i):
    lo, hi = int(corpus.doc_off[i]), int(corpus.doc_off[i + 1])
    words = []
    for j in range(lo, hi):
        words += [corpus.term(int(corpus.pairs[2 * j]))] * int(corpus.pairs[2 * j + 1])
    return " ".join(words)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--adds", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    from nxsearch_b200 import capi, tools

    def make_queries(term_ids, n):
        """1..4-term OR queries as in bench.py."""
        out, pos = [], 0
        for i in range(n):
            nt = 1 + (i % 4)
            out.append((None, None, [int(t) for t in term_ids[pos:pos + nt]]))
            pos += nt
        return out

    corpus = tools.Corpus.generate(args.docs, args.vocab)
    extra = tools.Corpus.generate(2 * args.adds + 16, args.vocab, first_doc=args.docs)
    base = tempfile.mkdtemp(prefix="nxsb_refresh_")
    out = {"docs": args.docs, "vocab": args.vocab, "batch": args.batch}
    try:
        nxs = capi.Nxs(base)
        nxs.create_index("p").close()
        corpus.write(f"{base}/data/p/nxsterms", f"{base}/data/p/nxsdtmap")
        t0 = time.perf_counter()
        idx = nxs.open_index("p")
        out["open_s"] = time.perf_counter() - t0
        qt = corpus.query_terms(4 * args.batch * 8)
        qs = make_queries(qt, args.batch * 8)
        strings = [[" OR ".join(corpus.term(t) for t in leaves).encode() for _, _, leaves in qs[i * args.batch:(i + 1) * args.batch]]
                   for i in range(8)]
        arrays = [(C.c_char_p * len(s))(*s) for s in strings]

        def batch_ms(n=24):
            idx.search_batch_arrays(arrays[0], 10, algo="BM25")
            t0 = time.perf_counter()
            for i in range(n):
                idx.search_batch_arrays(arrays[i % 8], 10, algo="BM25")
            return 1000 * (time.perf_counter() - t0) / n

        t0 = time.perf_counter()
        idx.search_batch_arrays(arrays[0], 10, algo="BM25")
        out["first_search_full_build_s"] = time.perf_counter() - t0
        out["batch_ms_plain_image"] = batch_ms()

        nxt = 0
        def add(n):
            nonlocal nxt
            t0 = time.perf_counter()
            for _ in range(n):
                idx.add(int(extra.doc_ids[nxt]), doc_text(extra, nxt))
                nxt += 1
            return time.perf_counter() - t0

        out["add_s_per_doc"] = add(args.adds) / args.adds
        t0 = time.perf_counter()
        idx.search_batch_arrays(arrays[1], 10, algo="BM25")
        out["first_search_after_adds_s"] = time.perf_counter() - t0
        out["image_after_adds"] = idx.image_stats()
        out["batch_ms_one_delta_segment"] = batch_ms()

        for d in (3, 4_000, 77_777):
            idx.remove(int(corpus.doc_ids[d]))
        idx.remove(int(extra.doc_ids[5]))
        t0 = time.perf_counter()
        idx.search_batch_arrays(arrays[2], 10, algo="BM25")
        out["first_search_after_4_removes_s"] = time.perf_counter() - t0
        out["batch_ms_one_segment_4_dead"] = batch_ms()

        add(1)
        t0 = time.perf_counter()
        idx.search_batch_arrays(arrays[3], 10, algo="BM25")
        out["first_search_after_1_add_s"] = time.perf_counter() - t0
        out["image_now"] = idx.image_stats()

        # the behaviour this replaces: any change rebuilds the whole image
        os.environ["NXSB_REFRESH_FULL"] = "1"
        add(1)
        t0 = time.perf_counter()
        idx.search_batch_arrays(arrays[4], 10, algo="BM25")
        out["first_search_after_1_add_full_rebuild_s"] = time.perf_counter() - t0
        del os.environ["NXSB_REFRESH_FULL"]
        out["batch_ms_after_rebuild"] = batch_ms()
        # ... and the rebuild with the postings converted on the host (before SURVEY 8f N2)
        os.environ["NXSB_REFRESH_FULL"] = os.environ["NXSB_IMAGE_HOST_PAIRS"] = "1"
        add(1)
        t0 = time.perf_counter()
        idx.search_batch_arrays(arrays[5], 10, algo="BM25")
        out["first_search_after_1_add_full_rebuild_host_pairs_s"] = time.perf_counter() - t0
        del os.environ["NXSB_REFRESH_FULL"], os.environ["NXSB_IMAGE_HOST_PAIRS"]
        out["image_end"] = idx.image_stats()
        idx.close()
        nxs.close()
    finally:
        shutil.rmtree(base, ignore_errors=True)
    line = json.dumps(out)
    print(line)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(line + "\n")


if __name__ == "__main__":
    main()
