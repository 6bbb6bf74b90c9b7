"""Measured lines for the BASELINE.json configs that are not bench.py's
headline: C3 boolean TF-IDF top-100 and C4 fuzzy.

    python scripts/configs_bench.py c3 [--docs N]
    python scripts/configs_bench.py c4 [--vocab V --queries Q]

Each prints one JSON line (device times from CUDA events inside the engine,
a bounded CPU sample of the oracle port beside it, and an in-run parity check
of the sampled queries).  bench.py stays the driver-facing contract; these
lines are committed under profiles/.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from nxsearch_b200 import engine as eng, tools  # noqa: E402
import _oracle  # noqa: E402
from _oracle import OP_AND, OP_ANDNOT, OP_OR  # noqa: E402

PEAK = 6537.3
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def c3_queries(term_ids, n):
    """SURVEY 8d C3 templates: a AND b | (a OR b) AND c | a AND NOT b |
    (a OR b) AND (c OR d) AND NOT (e OR f); token order right-to-left."""
    shapes = ["ab&", "ab|c&", "ab-", "ab|cd|&ef|-"]
    out, pos = [], 0
    for i in range(n):
        shape = shapes[i % 4]
        nl = sum(ch.isalpha() for ch in shape)
        leaves = [int(t) for t in term_ids[pos:pos + nl]]
        pos += nl
        toks = []
        for t in reversed(leaves):
            if t not in toks:
                toks.append(t)
        slot = {t: s for s, t in enumerate(toks)}
        prog, li = [], 0
        for ch in shape:
            if ch == "&":
                prog.append(OP_AND)
            elif ch == "|":
                prog.append(OP_OR)
            elif ch == "-":
                prog.append(OP_ANDNOT)
            else:
                prog.append(slot[leaves[li]])
                li += 1
        out.append((toks, prog))
    return out


def run_c3(args):
    t0 = time.time()
    corpus = tools.Corpus.generate(args.docs, args.vocab)
    df = np.asarray(corpus.term_df)
    e = eng.Engine(0)
    e.load_corpus(corpus)
    nb = 6
    qt = corpus.query_terms(6 * args.batch * nb, seed=tools.SEED + 3)
    qs = c3_queries(qt, args.batch * nb)
    batches = [qs[i * args.batch:(i + 1) * args.batch] for i in range(nb)]
    hb = [eng.Batch.from_lists(eng.ALGO_TFIDF, 100, b) for b in batches]
    hs = [e.upload(b) for b in hb]
    for i in range(3):
        e.run(hs[i % nb])
    e.sync()
    l0 = e.launches
    for i in range(args.steps):
        e.run(hs[i % nb])
    e.sync()
    launches = e.launches - l0
    t = e.timings(args.steps)
    total_ms = sum(t.values()) / args.steps
    tile_ms = t.get("score_tiles", 0.0) / args.steps
    byts = np.mean([sum(8 * int(df[x - 1]) for toks, _ in b for x in toks) for b in batches])
    # bounded CPU sample + parity of the sampled queries
    ora = _oracle.OracleIndex(corpus)
    counts, ids, scores = e.search(hb[0])
    done, t1 = 0, time.perf_counter()
    for toks, prog in batches[0][:args.cpu_sample]:
        ora.search(_oracle.TFIDF, 100, toks, prog)
        done += 1
        if time.perf_counter() - t1 > args.cpu_budget:
            break
    cpu_dt = time.perf_counter() - t1
    checked = 0
    for i, (toks, prog) in enumerate(batches[0][:min(done, 12)]):
        all_ids, all_sc = ora.search_all(_oracle.TFIDF, toks, prog)
        _oracle.check_topk(ids[i, :counts[i]], scores[i, :counts[i]], all_ids, all_sc, 100, exact_scores=True)
        checked += 1
    line = {
        "config": "C3: %d docs, TF-IDF, nested AND/OR/NOT templates, top-100, batch %d" % (args.docs, args.batch),
        "metric": "queries/s", "value": args.batch / (total_ms / 1e3), "ms_per_batch": total_ms,
        "kernel_ms": {k: v / args.steps for k, v in t.items()}, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": byts / (tile_ms / 1e3) / 1e9 if tile_ms else None, "peak": PEAK,
                     "unit": "GB/s", "frac": byts / (tile_ms / 1e3) / 1e9 / PEAK if tile_ms else None,
                     "algorithmic_bytes_per_launch": float(byts)},
        "cpu_baseline": {"value": done / cpu_dt, "unit": "queries/s", "cores": 1, "kind": "port",
                         "sample": "first %d queries of one batch" % done, "parity_checked_queries": checked},
        "setup_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


def run_c4(args):
    t0 = time.time()
    corpus = tools.Corpus.generate(2000, args.vocab)        # only the vocabulary matters
    e = eng.Engine(0)
    e.load_corpus(corpus)
    parent, edge, rank = corpus.bk_mirror()
    e.load_vocab(corpus.term_blob, corpus.term_off, corpus.term_total, parent, edge, rank)
    qs = corpus.fuzzy_terms(args.queries)
    e.fuzzy(qs[:1000])
    l0 = e.launches
    t1 = time.perf_counter()
    term, dist, _ = e.fuzzy(qs)
    wall = time.perf_counter() - t1
    launches = e.launches - l0
    t = e.last_timings()
    dev_ms = sum(t.values())
    lens = np.diff(np.asarray(corpus.term_off)).astype(np.int64)
    pairs = float(len(qs)) * corpus.n_terms
    steps = float(len(qs)) * float(lens.sum())
    # bounded CPU sample: the reference's BK-tree walk (oracle port) + parity
    ora = _oracle.OracleIndex(corpus)
    done, t1 = 0, time.perf_counter()
    bad = 0
    for i, q in enumerate(qs[:args.cpu_sample]):
        tt, cands, dists, _ = ora.fuzzy(q)
        done += 1
        if tt != term[i]:
            bad += 1
        elif tt:
            s = corpus.term(int(tt)).encode()
            if dist[i] != _oracle.port().ora_levdist(q, len(q), s, len(s)):
                bad += 1
        if time.perf_counter() - t1 > args.cpu_budget:
            break
    cpu_dt = time.perf_counter() - t1
    assert bad == 0, f"{bad} fuzzy answers differ from the reference's BK-tree search"
    line = {
        "config": "C4: Levenshtein <= 2, %d query terms against a %d-term vocabulary" % (len(qs), corpus.n_terms),
        "metric": "lookups/s", "value": len(qs) / (dev_ms / 1e3), "device_ms": dev_ms, "wall_ms_host_call": wall * 1e3,
        "kernel_ms": t, "gpu_launches": int(launches),
        "pairs_per_s": pairs / (dev_ms / 1e3), "myers_steps_per_s": steps / (dev_ms / 1e3),
        "resolved": int((term != 0).sum()),
        "cpu_baseline": {"value": done / cpu_dt, "unit": "lookups/s", "cores": 1, "kind": "port",
                         "sample": "first %d query terms" % done, "parity_checked": done},
        "setup_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c3", "c4"])
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--queries", type=int, default=100_000)
    ap.add_argument("--cpu-sample", type=int, default=200)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    if args.config == "c3":
        args.cpu_sample = min(args.cpu_sample, 48)
        run_c3(args)
    else:
        run_c4(args)


if __name__ == "__main__":
    main()
