#!/usr/bin/env python
"""Headline benchmark: BM25 top-10 queries/s on a 10M-document synthetic index.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)

A step is one batch of 1024 synthetic OR-queries (1-4 Zipf terms each, SURVEY
section 8d "C2") PER GPU, scored against the whole index, top-10 per query.

Layouts for N > 1 (--layout, default auto):
  replica  the 10M-document image is 4.4 GB and fits one B200 forty times
           over, so every GPU holds the whole index and takes its own stream
           of query batches -- the reference's own concurrency model, one
           process with a private index per worker (ref docs/c-api.md:5-8).
           No data-path collective; weak scaling (N batches per step).
  shard    the index is document-sharded, every rank scores every query on
           its shard with global statistics, per-shard top-k lists are merged
           after an NCCL all-gather (strong scaling).  What a 100M-document
           index needs (--docs 100000000), and measured beside the replica
           line as `doc_sharded`: exact top-k pruning does not strong-scale,
           because each shard pays the threshold warm-up of a whole index.

One JSON line on stdout (rank 0): see the contract in the task description.
`value`  = device-resident throughput (batches already in HBM),
`e2e`    = the same through the reference-facing C API with host buffers
           (query strings in, result arrays out) on every rank,
`roofline` = the scoring kernel against the measured HBM peak,
`cpu_baseline` = the oracle port on this box's host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "BM25 top-10 queries/s, 10M-doc synthetic index"
UNIT = "queries/s"
E2E_REPS = 5       # repetitions of the K-step end-to-end measurement (median reported)
E2E_DEPTH = int(os.environ.get("NXSB_BENCH_DEPTH", "4"))      # batches in flight in the pipelined end-to-end leg (the library's maximum)

# BASELINE.json configs served by this file (--config): the query shape, the
# ranking algorithm and the limit.  c2 is the headline; c5 is c2 at 100M
# documents, document-sharded; c4 (fuzzy) has its own step, see run_c4().
CONFIGS = {
    "c2": dict(algo="BM25", limit=10, docs=10_000_000, shape="or"),
    "c3": dict(algo="TF-IDF", limit=100, docs=10_000_000, shape="bool"),
    "c5": dict(algo="BM25", limit=10, docs=100_000_000, shape="or"),
}


def algo_ids(args):
    """(engine constant, oracle constant, C API parameter value) of the run's algorithm."""
    import _oracle
    from nxsearch_b200 import engine as eng_mod
    return ((eng_mod.ALGO_BM25, _oracle.BM25, "BM25") if args.algo == "BM25"
            else (eng_mod.ALGO_TFIDF, _oracle.TFIDF, "TF-IDF"))


def metric_name(args) -> str:
    """BASELINE.json's metric; a run of another config or size says so."""
    if args.config == "c2" and args.docs == 10_000_000 and args.limit == 10:
        return METRIC
    docs = f"{args.docs // 1_000_000}M" if args.docs % 1_000_000 == 0 else str(args.docs)
    kind = "" if args.shape == "or" else " nested AND/OR/NOT"
    return f"{args.algo}{kind} top-{args.limit} queries/s, {docs}-doc synthetic index"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries loaded later (NCCL's
# version banner, for one) write to fd 1 directly, so fd 1 is pointed at
# stderr for the whole run and the line goes out through the saved descriptor.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


AUX: dict = {}      # side measurements of the e2e leg (index open, image build)


def emit(line: dict) -> None:
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


# --------------------------------------------------------------------------
# workload


def make_queries(term_ids: np.ndarray, n_queries: int, shape: str = "or"):
    """n_queries queries as (tokens, program, query-string template).

    shape "or": 1..4 terms (cycling) joined by OR; "bool":
    the four C3 templates.  Token order follows the reference: leaves
    right-to-left, duplicates merged (ref src/query/query.c:89-103,
    src/core/tokenizer.c:94-117).  The template holds {0}, {1} ... for the
    leaves' strings; query_string() fills it."""
    OP_AND, OP_OR, OP_ANDNOT = -2, -3, -4
    bool_shapes = [("ab&", "{0} AND {1}"), ("ab|c&", "({0} OR {1}) AND {2}"), ("ab-", "{0} AND NOT {1}"),
                   ("ab|cd|&ef|-", "({0} OR {1}) AND ({2} OR {3}) AND NOT ({4} OR {5})")]
    out, pos = [], 0
    for i in range(n_queries):
        if shape == "or":
            nt = 1 + (i % 4)
            postfix = "a" + "".join(chr(ord("a") + j) + "|" for j in range(1, nt))
            text = " OR ".join("{%d}" % j for j in range(nt))
        else:
            postfix, text = bool_shapes[i % 4]
            nt = sum(ch.isalpha() for ch in postfix)
        leaves = [int(t) for t in term_ids[pos:pos + nt]]
        pos += nt
        toks: list[int] = []
        for t in reversed(leaves):
            if t not in toks:
                toks.append(t)
        slot = {t: s for s, t in enumerate(toks)}
        prog = []
        for ch in postfix:
            prog.append({"&": OP_AND, "|": OP_OR, "-": OP_ANDNOT}[ch] if not ch.isalpha()
                        else slot[leaves[ord(ch) - ord("a")]])
        out.append((toks, prog, (text, leaves)))
    return out


def query_string(corpus, item) -> str:
    text, leaves = item[2]
    return text.format(*(corpus.term(t) for t in leaves))


def batch_bytes(batch_items, df) -> int:
    """Algorithmic bytes: 8 B x sum of df over the resolved tokens."""
    return int(sum(8 * int(df[t - 1]) for toks, _, _ in batch_items for t in toks))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                     "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self) -> dict:
        self._halt.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(s[3 + j].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(sm), "reasons": reasons}


def measured_peak() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(args, per_gpu_docs: int):
    """DRAM bytes per launch of the scoring kernel from the committed ncu
    capture (profiles/ncu_traffic.json) -- only when that capture is of this
    workload (same documents per GPU, batch, limit); else None."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    try:
        rec = json.loads(p.read_text())
        if (rec.get("docs_per_gpu"), rec.get("batch"), rec.get("limit")) == (per_gpu_docs, args.batch, args.limit):
            return rec
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------
# reference arm (CPU)


def run_reference(args, rank: int, world: int) -> None:
    """The reference's CPU path on this box's host cores, bounded sample.

    The compiled reference (oracle/_ref) needs ~60 s per million documents to
    open an index (measured; DESIGN.md "Oracle"), i.e. ~10 min at 10M, so the
    arm times the oracle port -- the restatement pinned bit-for-bit to the
    reference -- on the full-size index instead, one query per host thread."""
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from nxsearch_b200 import tools
    import _oracle

    if args.config == "c4":
        # The reference's BK-tree search (oracle port) over the same vocabulary, all host threads.
        corpus = tools.Corpus.generate(2000, args.vocab)
        ora = _oracle.OracleIndex(corpus)
        cores = len(os.sched_getaffinity(0))
        per_step = max(cores, min(args.ref_sample * 4, args.fuzzy_terms))
        n_steps = args.warmup + args.steps
        qs = corpus.fuzzy_terms(per_step * n_steps)
        ora.fuzzy(qs[0])
        times = []
        with ThreadPoolExecutor(max_workers=cores) as pool:
            for s in range(n_steps):
                t0 = time.perf_counter()
                list(pool.map(lambda q: ora.fuzzy(q)[0], qs[s * per_step:(s + 1) * per_step]))
                if s >= args.warmup:
                    times.append(time.perf_counter() - t0)
                if sum(times) > args.ref_budget and times:
                    break
        value = per_step * len(times) / sum(times)
        emit({"metric": f"fuzzy term lookups/s (Levenshtein <= 2), {args.vocab}-term vocabulary", "value": value,
              "unit": "lookups/s", "n_gpus": world, "steps": len(times), "warmup": args.warmup,
              "ms_per_step": 1000 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "u32", "data": "synthetic", "impl": "reference",
              "config": {"workload": f"C4: {per_step} query terms per step against {corpus.n_terms} vocabulary terms",
                         "vocab": corpus.n_terms, "terms_per_step": per_step},
              "cpu_baseline": {"value": value, "unit": "lookups/s", "cores": cores, "kind": "port",
                               "sample": f"{per_step} query terms/step x {len(times)} steps"},
              "e2e": {"value": value, "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
              "gpu_launches": 0})
        return

    t0 = time.time()
    corpus = tools.Corpus.generate(args.docs, args.vocab)
    log(f"[ref] corpus {corpus.n_docs} docs / {corpus.n_pairs} postings in {time.time() - t0:.1f}s")
    t0 = time.time()
    ora = _oracle.OracleIndex(corpus)
    log(f"[ref] oracle index in {time.time() - t0:.1f}s")
    cores = len(os.sched_getaffinity(0))
    per_step = max(cores, min(args.ref_sample, args.batch))
    n_steps = args.warmup + args.steps
    qt = corpus.query_terms(4 * per_step * n_steps)
    queries = make_queries(qt, per_step * n_steps, args.shape)
    _, ora_algo, _ = algo_ids(args)

    def one(q):
        toks, prog, _ = q
        return ora.search(ora_algo, args.limit, toks, prog)

    times = []
    with ThreadPoolExecutor(max_workers=cores) as pool:
        for s in range(n_steps):
            chunk = queries[s * per_step:(s + 1) * per_step]
            t0 = time.perf_counter()
            list(pool.map(one, chunk))
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
            if sum(times) > args.ref_budget and len(times) >= 1:
                break
    steps_done = len(times)
    total = sum(times)
    value = per_step * steps_done / total
    sample = f"{per_step} queries/step x {steps_done} steps of the same query stream, full {args.docs}-doc index"
    compiled = compiled_reference_check(args) if args.ref_real_docs > 0 else None
    line = {
        "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": steps_done,
        "warmup": args.warmup, "ms_per_step": 1000 * total / steps_done, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "impl": "reference",
        # our arm's config; what one timed step of THIS arm covers is `sample_queries_per_step`
        "config": {"workload": workload_name(args), "docs": args.docs, "vocab": args.vocab,
                   "batch": args.batch, "limit": args.limit, "sample_queries_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if compiled:
        line["compiled_reference"] = compiled
    emit(line)


def compiled_reference_check(args):
    """How the port relates to the real thing: the reference's own C files
    (oracle/_ref) and the port on the SAME reduced index -- the compiled
    reference needs about a minute per million documents to open one, so it
    cannot serve the 10M-document config inside a benchmark run -- one core
    each, the same queries.  Supplementary; not the line's value."""
    import shutil
    import _oracle
    from nxsearch_b200 import capi, tools

    ref = _oracle.ref()
    if ref is None:
        return {"unavailable": "oracle/_ref was not built (no /root/reference at build time)"}
    base = tempfile.mkdtemp(prefix="nxsb_ref_", dir=args.tmpdir)
    try:
        corpus = tools.Corpus.generate(args.ref_real_docs, args.vocab)
        boot = capi.Nxs(base)                      # index directory + params.db
        boot.create_index("r").close()
        boot.close()
        corpus.write(f"{base}/data/r/nxsterms", f"{base}/data/r/nxsdtmap")
        t0 = time.perf_counter()
        rn = capi.Nxs(base, lib=ref)
        ridx = rn.open_index("r")
        open_s = time.perf_counter() - t0
        qt = corpus.query_terms(4 * 64)
        queries = make_queries(qt, 64, args.shape)
        strings = [query_string(corpus, q) for q in queries]
        _, ora_algo, api_algo = algo_ids(args)
        done, t0 = 0, time.perf_counter()
        for q in strings:
            ridx.search(q, limit=args.limit, algo=api_algo)
            done += 1
            if time.perf_counter() - t0 > 10:
                break
        ref_qps = done / (time.perf_counter() - t0)
        ridx.close()
        rn.close()
        ora = _oracle.OracleIndex(corpus)
        t0 = time.perf_counter()
        for toks, prog, _ in queries[:done]:
            ora.search(ora_algo, args.limit, toks, prog)
        port_qps = done / (time.perf_counter() - t0)
        ora.close()
        return {"docs": args.ref_real_docs, "index_open_s": round(open_s, 1), "queries": done,
                "reference_queries_per_s_1core": ref_qps, "port_queries_per_s_1core": port_qps,
                "note": "same reduced index and queries for both; the line's value is the port on the full index"}
    finally:
        shutil.rmtree(base, ignore_errors=True)


def workload_name(args) -> str:
    what = ("BM25 OR-queries (1-4 terms)" if args.shape == "or" and args.algo == "BM25"
            else f"{args.algo} OR-queries (1-4 terms)" if args.shape == "or"
            else f"{args.algo} queries of the templates a AND b | (a OR b) AND c | a AND NOT b | "
                 "(a OR b) AND (c OR d) AND NOT (e OR f)")
    return (f"{args.config.upper()}: {args.docs} synthetic docs (Zipf(1.0) over {args.vocab} terms, 16-111 tokens), "
            f"batches of {args.batch} {what}, top-{args.limit}")


# --------------------------------------------------------------------------
# our arm


def pick_layout(args, world: int) -> str:
    """replica while the whole image fits one GPU with room to spare, else shard."""
    if world == 1:
        return "replica"
    if args.layout != "auto":
        return args.layout
    image_gb = args.docs * 64 * 8 / 1e9 * 1.6       # postings + block arrays + tables, generous
    return "replica" if image_gb < 90 else "shard"


class LaneRunner:
    """Resident batches straight on the engine's own two streams (no collective
    follows in the replica layout): consecutive handles alternate lanes, so two
    launches are in flight on the GPU at any time."""

    def __init__(self, engine):
        self.engine = engine

    def run(self, handle, n_q, k):
        self.engine.run(handle)


def timed_value_leg(args, torch, dist, world, dev, stream, engine, searcher, handles, bytes_local, barrier, rot=0,
                    join=None):
    """W warm-up steps, then K timed steps bracketed by barrier + synchronize,
    device time from CUDA events on the engine's stream, max over ranks."""
    n_distinct = len(handles)
    # rot: replicas draw from the same batches, each starting `rank` batches in.
    for s in range(args.warmup):
        searcher.run(handles[(s + rot) % n_distinct], args.batch, args.limit)
    barrier()
    launches0 = engine.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for s in range(args.steps):
        searcher.run(handles[(args.warmup + s + rot) % n_distinct], args.batch, args.limit)
    if join:
        join()                       # lane 0 waits for the other lane: the event below closes both
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = engine.launches - launches0
    runs_timed = min(args.steps, 256)
    kern = engine.timings(runs_timed)
    timed_bytes = sum(bytes_local[(args.warmup + s + rot) % n_distinct] for s in range(max(0, args.steps - 256), args.steps))
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    return elapsed_ms, launches, kern, timed_bytes, runs_timed


def run_ours(args, rank: int, world: int, local_rank: int) -> None:
    import torch
    import torch.distributed as dist
    from nxsearch_b200 import capi, dist as nxdist, engine as eng_mod, tools

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    layout = pick_layout(args, world)
    # A dedicated (non-default) stream shared by the engine's kernels, the NCCL
    # collectives and the timing events.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def load(lo, hi):
        t0 = time.time()
        corpus = tools.Corpus.generate(hi - lo, args.vocab, first_doc=lo)
        log(f"[{rank}] docs [{lo},{hi}) {corpus.n_pairs} postings generated in {time.time() - t0:.1f}s")
        if hi - lo == args.docs:
            df, tokens, ndocs = np.asarray(corpus.term_df).astype(np.uint32), corpus.token_count, corpus.n_docs
        else:
            df, tokens, ndocs = nxdist.allreduce_stats(np.asarray(corpus.term_df), corpus.token_count,
                                                       corpus.n_docs, device=dev)
        t0 = time.time()
        engine = eng_mod.Engine(local_rank)
        engine.load_corpus(corpus, df=df, token_count=tokens, doc_count=ndocs)
        if hi - lo != args.docs:
            engine.set_stream(stream.cuda_stream)        # shards: one stream with the NCCL collective
        log(f"[{rank}] HBM image built in {time.time() - t0:.1f}s")
        return corpus, engine, df

    def stage(corpus, engine, df, seed):
        """The batches of the run: queries, their algorithmic bytes on this GPU, resident handles."""
        n_distinct = min(args.warmup + args.steps, 48)
        qt = tools.query_terms(corpus.n_terms, df, 4 * args.batch * n_distinct, seed=seed)
        queries = make_queries(qt, args.batch * n_distinct, args.shape)
        batches = [queries[i * args.batch:(i + 1) * args.batch] for i in range(n_distinct)]
        local_df = np.asarray(corpus.term_df)
        bytes_local = [batch_bytes(b, local_df) for b in batches]
        host_batches = [eng_mod.Batch.from_lists(algo_ids(args)[0], args.limit, [(t, p) for t, p, _ in b]) for b in batches]
        handles = [engine.upload(hb) for hb in host_batches]
        return batches, bytes_local, host_batches, handles

    doc_sharded = None
    if layout == "replica":
        # ---- every GPU holds the whole index and takes its own query stream
        corpus, engine, df = load(0, args.docs)
        # Every replica serves the same query distribution: the same batches,
        # rank r starting r batches into the cycle.
        batches, bytes_local, host_batches, handles = stage(corpus, engine, df, tools.SEED + 1)
        searcher = LaneRunner(engine)
        lane0 = torch.cuda.ExternalStream(engine.lane_stream(0), device=dev)
        sampler = ClockSampler(local_rank)
        sampler.start()
        elapsed_ms, launches, kern, timed_bytes, runs_timed = timed_value_leg(
            args, torch, dist, world, dev, lane0, engine, searcher, handles, bytes_local, barrier, rot=rank,
            join=engine.lanes_join)
        clocks = sampler.stop()
        value = world * args.batch * args.steps / (elapsed_ms / 1000)
        e2e, h2d, d2h, e2e_note = e2e_capi(args, corpus, batches, capi, rank, world, local_rank, dist, barrier, dev)
        if world > 1 and not args.no_shard_leg:
            doc_sharded = shard_leg(args, rank, world, torch, dist, dev, stream, barrier, load, stage, nxdist, tools)
        parallelism = (f"replica x{world}: whole index on every GPU, {world} batches of {args.batch} per step "
                       "(the same cycle of batches on every GPU, rank r starting r batches in)")
        scaling = "weak"
    else:
        # ---- document shards + NCCL all-gather merge
        lo, hi = nxdist.shard_range(args.docs, rank, world)
        corpus, engine, df = load(lo, hi)
        batches, bytes_local, host_batches, handles = stage(corpus, engine, df, tools.SEED + 1)
        searcher = nxdist.ShardedSearcher(engine, rank, world)
        sampler = ClockSampler(local_rank)
        sampler.start()
        elapsed_ms, launches, kern, timed_bytes, runs_timed = timed_value_leg(
            args, torch, dist, world, dev, stream, engine, searcher, handles, bytes_local, barrier)
        clocks = sampler.stop()
        value = args.batch * args.steps / (elapsed_ms / 1000)
        e2e, h2d, d2h, e2e_note = e2e_sharded(args, engine, searcher, host_batches, barrier, dist, dev)
        parallelism = f"doc-shard x{world}: every query on every shard, NCCL all-gather + merge"
        scaling = "strong"

    # ---- roofline of the dominant kernel (this rank's GPU)
    peak, peak_src = measured_peak()
    tile_ms = kern.get("score_tiles", 0.0)
    lanes = 2 if layout == "replica" and os.environ.get("NXSB_LANES", "2") != "1" else 1
    if lanes > 1:
        # Two launches are in flight on two streams: a launch's event-bracketed
        # time on its own stream covers the sharing, so the rate is taken over
        # the timed region as a whole (the small kernels count against it).
        achieved = (timed_bytes * args.steps / runs_timed) / (elapsed_ms / 1000) / 1e9
    else:
        achieved = timed_bytes / (tile_ms / 1000) / 1e9 if tile_ms > 0 else 0.0
    rec = recorded_traffic(args, corpus.n_docs)
    roofline = {
        "bound": "hbm",
        "kernel": ("score_bmw_kernel<%s, 32-document blocks> (exact top-k with block-max pruning)" % args.algo
                   if args.shape == "or" and args.limit <= 128 else
                   "score_bmw_kernel<LOGIC, %s> (block bounds + present-token masks + a membership byte per document) for "
                   "queries with <= 3 positive terms, score_stream_kernel<LOGIC> (every posting streamed) for the rest; "
                   "kernel time is their sum" % args.algo),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": rec["dram_bytes_per_launch"] if rec else None, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": timed_bytes / max(runs_timed, 1),
        "kernel_ms_per_launch": tile_ms / max(runs_timed, 1),
        "launches_in_flight": lanes,
        "kernel_share_of_step": ((tile_ms / runs_timed) / (elapsed_ms / args.steps) / lanes) if tile_ms else None,
        "other_kernels_ms_per_step": {k: v / runs_timed for k, v in kern.items() if k != "score_tiles"},
        "note": ("achieved = SURVEY 8d's algorithmic bytes (8 B x sum of df over query tokens: what the reference's "
                 "exhaustive loop and the round-1 streaming kernel touch) / kernel time, so frac > 1 measures work "
                 "AVOIDED by pruning, not HBM utilisation; see `physical` for what the kernel really moves"
                 + ("; two launches overlap on two streams, so achieved = algorithmic bytes of the timed launches / "
                    "the timed region and kernel_ms_per_launch is a launch's time on its own stream while sharing "
                    "the GPU with the other" if lanes > 1 else "")
                 if args.shape == "or" else
                 "achieved = SURVEY 8d's algorithmic bytes (8 B x sum of df over query tokens) / kernel time; "
                 "head-term slices are re-read from L2 across the queries of a batch, so DRAM traffic is lower"),
    }
    if rec and tile_ms > 0:
        # DRAM rate of the scoring launches: per-launch traffic of the capture over the launch rate of this run
        ms = (elapsed_ms / args.steps) if lanes > 1 else tile_ms / max(runs_timed, 1)
        roofline["physical"] = {
            "dram_bytes_per_launch": rec["dram_bytes_per_launch"],
            "dram_gbs": rec["dram_bytes_per_launch"] / (ms / 1000) / 1e9,
            "hbm_frac": rec["dram_bytes_per_launch"] / (ms / 1000) / 1e9 / peak,
            "issue_slots_busy_pct": rec.get("issue_slots_busy_pct"),
            "l2_hit_pct": rec.get("l2_hit_pct"),
            "limiter": rec.get("limiter"),
            "source": rec.get("source"),
        }

    line = {
        "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "docs": args.docs, "vocab": args.vocab,
                   "batch": args.batch, "global_batch": args.batch * (world if layout == "replica" else 1),
                   "limit": args.limit, "layout": layout, "parallelism": parallelism,
                   "distinct_batches": len(handles),
                   "l2": "inputs larger than L2: %.1f GB of postings per GPU, a different batch every step"
                         % (corpus.n_pairs * 8 / 1e9)},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "path": e2e_note},
        "gpu_launches": int(launches) * (world if layout == "replica" else 1),
        "roofline": roofline,
    }
    if doc_sharded:
        line["doc_sharded"] = doc_sharded
    if "latency" in AUX:
        line["latency"] = AUX.pop("latency")
    if "e2e_single_process" in AUX:
        line["e2e_single_process"] = AUX.pop("e2e_single_process")
    if "e2e_single_process_shards" in AUX:
        line["e2e_single_process_shards"] = AUX.pop("e2e_single_process_shards")
    if "e2e_fuzzymatch_default" in AUX:
        line["e2e_fuzzymatch_default"] = AUX.pop("e2e_fuzzymatch_default")
    if AUX:
        line["index_load"] = dict(AUX, note="10M-document index through the public C API: nxs_index_open of the "
                                            "reference-format files, then the HBM image build inside the first search")
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        i = args.warmup % len(handles)
        line["cpu_baseline"] = cpu_baseline(args, corpus, batches[i], engine, host_batches[i])
    if rank == 0:
        emit(line)
    for h in handles:
        engine.release(h)
    engine.close()
    if world > 1:
        dist.destroy_process_group()


def run_c4(args, rank: int, world: int, local_rank: int) -> None:
    """BASELINE config 4: a step = one call resolving --fuzzy-terms misspelt
    query terms against a 1M-term vocabulary (Levenshtein <= 2, the term the
    reference's idxterm_fuzzysearch picks).  N > 1: every GPU holds the
    vocabulary and takes its own call per step (no collective)."""
    import torch
    import torch.distributed as dist
    import _oracle
    from nxsearch_b200 import engine as eng_mod, tools

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    corpus = tools.Corpus.generate(2000, args.vocab)           # only the vocabulary matters
    e = eng_mod.Engine(local_rank)
    e.load_corpus(corpus)
    parent, edge, rank_bfs = corpus.bk_mirror()
    e.load_vocab(corpus.term_blob, corpus.term_off, corpus.term_total, parent, edge, rank_bfs)
    n = args.fuzzy_terms
    qs = corpus.fuzzy_terms(n, seed=tools.SEED + 2 + 7919 * rank)
    for _ in range(args.warmup):
        e.fuzzy(qs)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0, dev_ms, kern = e.launches, 0.0, {}
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        term, dist_, _ = e.fuzzy(qs)                            # host strings in, host arrays out, synchronous
        for k, v in e.last_timings().items():
            kern[k] = kern.get(k, 0.0) + v
            dev_ms += v
    wall = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()
    launches = e.launches - l0
    if world > 1:
        t = torch.tensor([dev_ms, wall], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall = float(t[0].item()), float(t[1].item())
    lens = np.diff(np.asarray(corpus.term_off)).astype(np.int64)
    value = world * n * args.steps / (dev_ms / 1e3)
    steps_per_s = world * float(n) * float(lens.sum()) * args.steps / (dev_ms / 1e3)
    rec = None
    try:
        rec = json.loads((ROOT / "profiles" / "ncu_fuzzy.json").read_text())
    except Exception:
        pass
    blob_bytes = sum(len(q) for q in qs)
    line = {
        "metric": f"fuzzy term lookups/s (Levenshtein <= 2), {args.vocab}-term vocabulary", "value": value,
        "unit": "lookups/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"C4: {n} query terms (a vocabulary term with 1-2 byte edits) per step against "
                               f"{corpus.n_terms} vocabulary terms, the reference's BK-tree choice reproduced",
                   "vocab": corpus.n_terms, "terms_per_step": n, "layout": "replica",
                   "l2": "the vocabulary image (slots + signatures, ~40 MB) is L2-resident by design; every step "
                         "re-reads all of it for every query term"},
        "clocks": clocks,
        "e2e": {"value": world * n * args.steps / wall, "unit": "lookups/s", "h2d_bytes_per_step": blob_bytes + 4 * (n + 1),
                "d2h_bytes_per_step": 8 * n,
                "path": "nxsb_engine_fuzzy (engine C ABI: query strings in host memory in, term ids and distances "
                        "out to host arrays, synchronous) -- what nxs_index_search calls for unresolved tokens"},
        "gpu_launches": int(launches) * world,
        "roofline": {"bound": "int-alu", "kernel": "fuzzy_scan_kernel<u32> (signature prefilter + Myers bit-parallel)",
                     "achieved": steps_per_s / 1e9, "peak": None, "unit": "G Myers column steps/s (SURVEY 8d: sum of "
                     "len(vocabulary term) per query term)", "frac": None,
                     "pairs_per_s": world * float(n) * corpus.n_terms * args.steps / (dev_ms / 1e3),
                     "kernel_ms_per_launch": kern.get("fuzzy_scan", dev_ms) / args.steps,
                     "int_pipe": rec, "traffic": rec.get("dram_bytes_per_launch") if rec else None,
                     "note": "integer-pipe bound (SURVEY 8d); the signature tests skip most Myers steps the "
                             "algorithmic figure counts, so steps/s is a throughput score, not pipe utilisation"},
        "resolved_terms": int((term != 0).sum()),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from concurrent.futures import ThreadPoolExecutor
        ora = _oracle.OracleIndex(corpus)
        ora.fuzzy(qs[0])                                        # builds the BK-tree
        done, t1 = 0, time.perf_counter()
        for q in qs[:args.cpu_sample * 8]:
            ora.fuzzy(q)
            done += 1
            if time.perf_counter() - t1 > args.cpu_budget:
                break
        cpu_dt = time.perf_counter() - t1
        sample = list(range(0, n, max(1, n // 2000)))[:2000]
        with ThreadPoolExecutor(max_workers=max(1, min(16, len(os.sched_getaffinity(0))))) as pool:
            ref = list(pool.map(lambda i: ora.fuzzy(qs[i])[0], sample))
        lev = _oracle.port().ora_levdist
        for i, t_ref in zip(sample, ref):
            assert term[i] == t_ref, (qs[i], int(term[i]), t_ref)
            if t_ref:
                s_ = corpus.term(int(t_ref)).encode()
                assert dist_[i] == lev(qs[i], len(qs[i]), s_, len(s_))
        ora.close()
        line["cpu_baseline"] = {"value": done / cpu_dt, "unit": "lookups/s", "cores": 1, "kind": "port",
                                "sample": f"first {done} query terms, BK-tree walk over the same vocabulary",
                                "parity_checked_terms": len(sample),
                                "parity": "chosen term id and edit distance exact on %d terms" % len(sample)}
    if rank == 0:
        emit(line)
    e.close()
    if world > 1:
        dist.destroy_process_group()


def shard_leg(args, rank, world, torch, dist, dev, stream, barrier, load, stage, nxdist, tools):
    """The same index document-sharded over the ranks (strong scaling): every
    rank scores every query of a 1024-query batch on its shard, NCCL
    all-gather, merge.  Reported beside the replica line."""
    lo, hi = nxdist.shard_range(args.docs, rank, world)
    corpus, engine, df = load(lo, hi)
    batches, bytes_local, host_batches, handles = stage(corpus, engine, df, tools.SEED + 1)
    searcher = nxdist.ShardedSearcher(engine, rank, world)
    elapsed_ms, launches, kern, timed_bytes, runs_timed = timed_value_leg(
        args, torch, dist, world, dev, stream, engine, searcher, handles, bytes_local, barrier)
    out = {"value": args.batch * args.steps / (elapsed_ms / 1000), "unit": UNIT, "scaling": "strong",
           "ms_per_step": elapsed_ms / args.steps,
           "kernel_ms_per_launch": kern.get("score_tiles", 0.0) / max(runs_timed, 1),
           "parallelism": f"doc-shard x{world}: {hi - lo} documents per GPU, every query on every shard, "
                          "NCCL all-gather of the per-shard top-k + merge_topk_kernel",
           "note": "device-resident batches, same protocol as `value`"}
    for h in handles:
        engine.release(h)
    engine.close()
    corpus.close()
    return out


def e2e_capi(args, corpus, batches, capi, rank=0, world=1, local_rank=0, dist=None, barrier=None, dev=None):
    """The public C API on every rank, query strings in -> result arrays out.
    One process per GPU, each with its own nxs_t on the SAME index files
    (rank 0 writes them), the reference's multi-process model."""
    import shutil

    if rank == 0:
        base = tempfile.mkdtemp(prefix="nxsb_bench_", dir=args.tmpdir)
    else:
        base = None
    if world > 1:
        box = [base]
        dist.broadcast_object_list(box, src=0)
        base = box[0]
    os.environ["NXS_GPU_DEVICE"] = str(local_rank)
    try:
        if rank == 0:
            boot = capi.Nxs(base)
            boot.create_index("bench").close()
            boot.close()
            t0 = time.time()
            corpus.write(f"{base}/data/bench/nxsterms", f"{base}/data/bench/nxsdtmap")
            AUX["index_files_write_s"] = round(time.time() - t0, 2)
        if world > 1:
            barrier()
        t1 = time.time()
        nxs = capi.Nxs(base)
        idx = nxs.open_index("bench")
        AUX["nxs_index_open_s"] = round(time.time() - t1, 2)
        log(f"[{rank}] index opened through nxs_index_open in {time.time() - t1:.1f}s")
        import ctypes as C
        strings = [[query_string(corpus, q).encode() for q in b] for b in batches]
        # Host buffers as a C caller holds them: an array of C strings in, and
        # every result drained through the public iterator into host arrays
        # (nxsb_resp_collect) -- no Python objects per result.
        arrays = [(C.c_char_p * len(s))(*s) for s in strings]
        params = dict(algo=algo_ids(args)[2], fuzzymatch=False)
        t0 = time.time()
        idx.search_batch(strings[0][:8], limit=args.limit, **params)          # builds the HBM image
        AUX["first_search_image_build_s"] = round(time.time() - t0, 2)
        log(f"[{rank}] first search (image build) {time.time() - t0:.1f}s")
        n = len(batches)
        arrays = arrays[rank % n:] + arrays[:rank % n]          # rank r starts r batches into the cycle
        strings = strings[rank % n:] + strings[:rank % n]
        for s in range(args.warmup):
            idx.search_batch_arrays(arrays[s % n], args.limit, **params)
        # K steps per measurement are a few tens of milliseconds here, so one
        # scheduling hiccup on one rank would decide the max over ranks: each leg
        # is measured E2E_REPS times (every repetition is K steps between
        # barriers, max over ranks) and the median repetition is reported.
        def over_ranks(dt):
            if world > 1:
                import torch
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item())
            return dt

        serial_reps, piped_reps = [], []
        for rep in range(E2E_REPS):
            # (1) one synchronous nxs_index_search_batch call per step
            if barrier:
                barrier()
            t0 = time.perf_counter()
            for s in range(args.steps):
                counts, ids, scores = idx.search_batch_arrays(arrays[(args.warmup + s) % n], args.limit, **params)
            serial_reps.append(over_ranks(time.perf_counter() - t0))
            # (2) the same calls split in begin/end, E2E_DEPTH batches in flight: the
            # host parses the next batches while the GPU scores two at a time.  Every step still
            # takes its strings from host memory, copies its descriptors to the
            # device and drains its results into host arrays.
            if barrier:
                barrier()
            t0 = time.perf_counter()
            flight, issued = [], 0
            for s in range(args.steps):
                while issued < args.steps and len(flight) < E2E_DEPTH:
                    flight.append(idx.search_batch_begin(arrays[(args.warmup + issued) % n], args.limit, **params))
                    issued += 1
                counts, ids, scores = idx.search_batch_end_arrays(flight.pop(0))
            piped_reps.append(over_ranks(time.perf_counter() - t0))
        dt_serial = sorted(serial_reps)[len(serial_reps) // 2]
        dt = sorted(piped_reps)[len(piped_reps) // 2]
        log(f"[{rank}] e2e: synchronous {world * args.batch * args.steps / dt_serial:.0f} q/s, "
            f"pipelined ({E2E_DEPTH} in flight) {world * args.batch * args.steps / dt:.0f} q/s")
        assert len(counts) == args.batch and int(counts.max()) <= args.limit and int(counts.sum()) > 0
        # the drained arrays are what the list-building wrapper returns
        last = (args.warmup + args.steps - 1) % n
        ref = idx.search_batch(strings[last][:16], limit=args.limit, **params)
        for i, r in enumerate(ref):
            assert [d for d, _ in r] == [int(x) for x in ids[i, :counts[i]]]
        if world == 1 and not args.no_fuzzy_leg:
            # The same pipelined calls with the reference's DEFAULT parameters
            # (fuzzymatch on) and one query in twenty carrying a misspelt term:
            # each such batch resolves ~51 terms by a vocabulary scan on the GPU
            # (nxsb_engine_fuzzy inside nxs_index_search_batch_begin).
            fz = corpus.fuzzy_terms(sum(len(b) for b in batches) // 20 + 8)
            k, arrays_fz = 0, []
            for b in batches:
                row = []
                for qi, item in enumerate(b):
                    text, leaves = item[2]
                    names = [corpus.term(t) for t in leaves]
                    if qi % 20 == 7:
                        names[0] = fz[k].decode()
                        k += 1
                    row.append(text.format(*names).encode())
                arrays_fz.append((C.c_char_p * len(row))(*row))
            pf = dict(algo=algo_ids(args)[2])
            for s in range(args.warmup):
                idx.search_batch_arrays(arrays_fz[s % n], args.limit, **pf)
            reps = []
            for rep in range(E2E_REPS):
                t0 = time.perf_counter()
                flight, issued = [], 0
                for s in range(args.steps):
                    while issued < args.steps and len(flight) < E2E_DEPTH:
                        flight.append(idx.search_batch_begin(arrays_fz[(args.warmup + issued) % n], args.limit, **pf))
                        issued += 1
                    fc, fi, fs = idx.search_batch_end_arrays(flight.pop(0))
                reps.append(time.perf_counter() - t0)
            dtf = sorted(reps)[len(reps) // 2]
            AUX["e2e_fuzzymatch_default"] = {
                "value": args.batch * args.steps / dtf, "unit": UNIT,
                "workload": "the same batches, fuzzymatch at its default (on), every 20th query with one "
                            "misspelt term (1-2 byte edits of a vocabulary term) resolved by the GPU vocabulary scan",
                "fuzzy_lookups_per_batch": args.batch // 20}
            log(f"[0] e2e with fuzzymatch on and 5% of the queries misspelt: {args.batch * args.steps / dtf:.0f} q/s")
        ntok = np.mean([sum(len(t) for t, _, _ in b) for b in batches])
        nprog = np.mean([sum(len(p) for _, p, _ in b) for b in batches])
        h2d = int(args.batch * 16 + 4 * ntok + 4 * nprog + 8 * args.batch) * world
        d2h = int(args.batch * args.limit * 16 + 4 * args.batch) * world
        idx.close()
        nxs.close()
        if rank == 0 and not args.no_latency:
            # The call every caller of the reference makes: ONE query per
            # nxs_index_search (ref src/core/lua.c:341-367, src/utils/benchmark.c:204),
            # from a C program linked with libnxsearch.so, its own process and nxs_t.
            import _ccaller
            try:
                singles = [q.decode() for q in strings[0][:args.latency_queries]]
                t0 = time.time()
                lat = _ccaller.latency(base, "bench", algo_ids(args)[2], args.limit, singles, device=local_rank)
                lat["path"] = ("nxs_index_search, one query per call, from tests/c/nxs_caller.c (C, linked with "
                               "libnxsearch.so); includes parsing, term lookup, H2D, kernels, D2H, response")
                AUX["latency"] = lat
                log(f"[0] single-query nxs_index_search: p50 {lat['p50_us']:.0f} us, p99 {lat['p99_us']:.0f} us, "
                    f"{lat['serial_queries_per_s']:.0f} q/s serial ({time.time() - t0:.1f}s incl. open + image build)")
            except Exception as exc:            # the leg is informational
                log(f"[0] latency leg failed: {exc}")
        # The other ranks must not wait inside an NCCL barrier here (its kernel
        # spins on their GPUs, which this leg uses): a gloo barrier on the CPU.
        cpu_group = None
        if world > 1 and not args.no_single_process_leg:
            cpu_group = dist.new_group(backend="gloo")
            barrier()
            dist.barrier(group=cpu_group)
        if rank == 0 and world > 1 and not args.no_single_process_leg:
            # ONE process, one nxs_t, all the GPUs: NXS_GPU_DEVICES makes the library keep a
            # replica of the image per device and split each batch between them.  The other
            # ranks are idle (they wait at the barrier below); a call carries `world` batches.
            os.environ["NXS_GPU_DEVICES"] = f"0-{world - 1}"
            try:
                big = [(C.c_char_p * (world * args.batch))(*sum((strings[(j * world + r) % n] for r in range(world)), []))
                       for j in range(min(n, 8))]
                seen = None
                # ... and the same calls with NXS_GPU_LAYOUT=shards: a range of the documents per
                # device, every device scores every query, the lists merge on the first device
                # (what an index larger than one GPU needs; the answers must be the same).
                for layout in ("replicas", "shards"):
                    os.environ["NXS_GPU_LAYOUT"] = layout
                    t0 = time.time()
                    nxs2 = capi.Nxs(base)
                    idx2 = nxs2.open_index("bench")
                    first = idx2.search_batch_arrays(big[0], args.limit, **params)  # builds the images
                    built = time.time() - t0
                    if seen is None:
                        seen = [a.copy() for a in first]
                    else:
                        assert all(np.array_equal(a, b) for a, b in zip(seen, first)), \
                            "shards and replicas disagree"
                    for s in range(args.warmup):
                        idx2.search_batch_arrays(big[s % len(big)], args.limit, **params)
                    reps = []
                    for rep in range(E2E_REPS):
                        t0 = time.perf_counter()
                        flight, issued = [], 0
                        for s in range(args.steps):
                            while issued < args.steps and len(flight) < E2E_DEPTH:
                                flight.append(idx2.search_batch_begin(big[issued % len(big)], args.limit, **params))
                                issued += 1
                            idx2.search_batch_end_arrays(flight.pop(0))
                        reps.append(time.perf_counter() - t0)
                    dts = sorted(reps)[len(reps) // 2]
                    how = (f"split over {world} replicas inside the library" if layout == "replicas" else
                           f"every query scored on all {world} document ranges, top-{args.limit} lists merged on "
                           f"device 0 over NVLink; results equal to the replicas' bit for bit")
                    AUX["e2e_single_process" + ("" if layout == "replicas" else "_shards")] = {
                        "value": world * args.batch * args.steps / dts, "unit": UNIT,
                        "path": f"one process, one nxs_t, NXS_GPU_DEVICES=0-{world - 1}"
                                f"{'' if layout == 'replicas' else ' NXS_GPU_LAYOUT=shards'}: "
                                f"nxs_index_search_batch_begin/_end with {world * args.batch} queries per call, {how}",
                        "open_and_build_s": round(built, 2)}
                    log(f"[0] e2e, one process driving {world} GPUs ({layout}): "
                        f"{world * args.batch * args.steps / dts:.0f} q/s")
                    idx2.close()
                    nxs2.close()
            except Exception as exc:
                log(f"[0] single-process leg failed: {exc}")
            finally:
                os.environ.pop("NXS_GPU_DEVICES", None)
                os.environ.pop("NXS_GPU_LAYOUT", None)
        if cpu_group is not None:
            dist.barrier(group=cpu_group)
        if world > 1:
            barrier()
        return (world * args.batch * args.steps / dt, h2d, d2h,
                "nxs_index_search_batch_begin/_end (C API: C strings in, results drained via "
                "nxs_resp_iter_result into host arrays; four batches in flight, two at a time on the GPU)%s; median of 5 K-step "
                "measurements; synchronous nxs_index_search_batch: %.0f queries/s"
                % (f" on each of {world} processes, one GPU each, sharing the index files" if world > 1 else "",
                   world * args.batch * args.steps / dt_serial))
    finally:
        if rank == 0:
            shutil.rmtree(base, ignore_errors=True)


def e2e_sharded(args, engine, searcher, host_batches, barrier, dist, dev):
    """N > 1: engine C ABI with host descriptor arrays + all-gather merge + D2H."""
    import torch

    n = len(host_batches)

    for s in range(args.warmup):
        searcher.collect(searcher.submit(host_batches[s % n]))
    barrier()
    # Two batches in flight per rank: batch s+1 is submitted (descriptors
    # H2D, scoring, all-gather, merge, records D2H) before batch s is awaited.
    t0 = time.perf_counter()
    ticket = searcher.submit(host_batches[args.warmup % n])
    for s in range(args.steps):
        nxt = searcher.submit(host_batches[(args.warmup + s + 1) % n]) if s + 1 < args.steps else None
        recs = searcher.collect(ticket)
        ticket = nxt
    barrier()
    dt = time.perf_counter() - t0
    assert recs.shape == (args.batch, args.limit) and int(recs["valid"].sum()) > 0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    hb = host_batches[0]
    h2d = int(hb.queries.nbytes + hb.tokens.nbytes + hb.prog.nbytes)
    d2h = int(args.batch * args.limit * 16)
    return args.batch * args.steps / float(t.item()), h2d, d2h, "engine C ABI nxsb_engine_search_begin_dev/_end (host descriptors in, merged records out to pinned host memory) + NCCL all-gather + merge_topk_kernel; two batches in flight"


def cpu_baseline(args, corpus, batch_items, engine, host_batch):
    """Oracle port on a bounded sample of one timed batch, 1 core; the same
    queries double as an in-run parity check of the GPU results."""
    import _oracle

    _, ora_algo, _ = algo_ids(args)
    t0 = time.time()
    ora = _oracle.OracleIndex(corpus)
    log(f"[0] oracle index built in {time.time() - t0:.1f}s")
    counts, ids, scores = engine.search(host_batch)
    done, checked = 0, 0
    t0 = time.perf_counter()
    for i, (toks, prog, _) in enumerate(batch_items[:args.cpu_sample]):
        ora.search(ora_algo, args.limit, toks, prog)
        done += 1
        if time.perf_counter() - t0 > args.cpu_budget:
            break
    dt = time.perf_counter() - t0
    from concurrent.futures import ThreadPoolExecutor
    n_check = min(len(batch_items), args.parity_queries)
    with ThreadPoolExecutor(max_workers=max(1, min(16, len(os.sched_getaffinity(0))))) as pool:
        truth = list(pool.map(lambda q: ora.search_all(ora_algo, q[0], q[1]), batch_items[:n_check]))
    for i, (all_ids, all_sc) in enumerate(truth):
        _oracle.check_topk(ids[i, :counts[i]], scores[i, :counts[i]], all_ids, all_sc, args.limit,
                           exact_scores=(args.algo != "BM25"))
        checked += 1
    ora.close()
    return {"value": done / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {done} queries of one timed batch, full {args.docs}-doc index",
            "parity_checked_queries": checked,
            "parity": "ids exact, scores %s, order modulo ties -- every match scored by the port on %d queries"
                      % ("bit-equal" if args.algo != "BM25" else "within 1e-5 relative", checked)}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE.json config: c2 BM25 OR top-10 (headline), c3 TF-IDF boolean top-100, "
                         "c4 fuzzy lookups, c5 = c2 at 100M documents, document-sharded")
    ap.add_argument("--docs", type=int, default=None)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--limit", type=int, default=None)
    ap.add_argument("--algo", default=None, choices=["BM25", "TF-IDF"])
    ap.add_argument("--parity-queries", type=int, default=64, help="queries of one batch checked against the oracle")
    ap.add_argument("--fuzzy-terms", type=int, default=100_000, help="c4: query terms per step")
    ap.add_argument("--cpu-sample", type=int, default=64)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-sample", type=int, default=64, help="queries per step of the reference arm")
    ap.add_argument("--ref-budget", type=float, default=60.0, help="seconds of timed CPU work, reference arm")
    ap.add_argument("--ref-real-docs", type=int, default=500_000,
                    help="reference arm: also time the compiled reference on an index of this many documents (0: skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-query nxs_index_search leg")
    ap.add_argument("--no-fuzzy-leg", action="store_true", help="skip the fuzzymatch-on end-to-end leg")
    ap.add_argument("--no-single-process-leg", action="store_true",
                    help="N > 1: skip the one-process-all-GPUs leg (NXS_GPU_DEVICES)")
    ap.add_argument("--latency-queries", type=int, default=1024)
    ap.add_argument("--layout", default="auto", choices=["auto", "replica", "shard"],
                    help="N > 1: whole index on every GPU with its own query stream, or document shards + NCCL merge")
    ap.add_argument("--no-shard-leg", action="store_true", help="replica layout: skip the doc_sharded side measurement")
    ap.add_argument("--tmpdir", default=None)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = CONFIGS.get(args.config, CONFIGS["c2"])
    args.docs = args.docs or cfg["docs"]
    args.limit = args.limit or cfg["limit"]
    args.algo = args.algo or cfg["algo"]
    args.shape = cfg["shape"]
    if args.config == "c5" and args.layout == "auto":
        args.layout = "shard"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world != args.gpus:
            log(f"note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world} rank(s)")
        if args.config == "c4":
            run_c4(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
