"""Document-sharded search over torch.distributed (one process per GPU).

The index is split into contiguous document ranges, one per rank.  Every rank
scores each query against its own shard with the whole-index statistics
(N, token count, df[] -- exchanged once at load with an all-reduce, because
BM25 / TF-IDF use global counts: ref src/algo/ranking.c:77-78,149-150,163),
writes its per-query top-k records straight into the all-gather send buffer,
and a merge kernel picks the global top-k on every rank.  The reference has no
counterpart: its only concurrency is independent processes with private
copies of the whole index (ref docs/c-api.md:5-8).

torch is used for what it is good at here -- device buffers, streams and the
NCCL plumbing; scoring, top-k and merge are the engine's own kernels.
"""
from __future__ import annotations

import numpy as np

REC_BYTES = 16  # { u64 doc_id; f32 score; u32 valid }


def shard_range(n_docs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous document range [lo, hi) of a rank (dense index = ascending id)."""
    return n_docs * rank // world, n_docs * (rank + 1) // world


def allreduce_stats(local_df: np.ndarray, local_tokens: int, local_docs: int, *, device=None):
    """Sum df[], token count and doc count over all ranks (any backend)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_df.astype(np.uint32), int(local_tokens), int(local_docs)
    dev = device or "cpu"
    df = torch.from_numpy(local_df.astype(np.int64)).to(dev)
    cnt = torch.tensor([int(local_tokens), int(local_docs)], dtype=torch.int64, device=dev)
    dist.all_reduce(df)
    dist.all_reduce(cnt)
    cnt = cnt.cpu()
    return df.cpu().numpy().astype(np.uint32), int(cnt[0]), int(cnt[1])


def merge_topk_host(recs: np.ndarray, limit: int) -> np.ndarray:
    """Reference semantics of the merge kernel on host arrays (used by the CPU
    tests): recs[g, q, r] structured (doc_id, score, valid) sorted per shard;
    ties prefer the higher shard, then the earlier position."""
    g, q, k = recs.shape
    out = np.zeros((q, limit), dtype=recs.dtype)
    for qi in range(q):
        items = []
        for gi in range(g):
            for r in range(k):
                rec = recs[gi, qi, r]
                if rec["valid"]:
                    items.append((-float(rec["score"]), -gi, r, rec))
        items.sort(key=lambda x: x[:3])
        for j, it in enumerate(items[:limit]):
            out[qi, j] = it[3]
    return out


REC_DTYPE = np.dtype([("doc_id", np.uint64), ("score", np.float32), ("valid", np.uint32)])


class ShardedSearcher:
    """One rank's shard plus the collective top-k merge."""

    def __init__(self, engine, rank: int, world: int):
        import torch

        self.engine, self.rank, self.world = engine, rank, world
        self.torch = torch
        # The engine's kernels, the NCCL collective and the merge must be on ONE
        # stream (the all-gather reads what the scorer wrote): bind the engine to
        # torch's current stream here rather than trust the caller to have done it.
        if torch.cuda.is_available():
            engine.set_stream(torch.cuda.current_stream().cuda_stream)
        self._bufs: dict[tuple[int, int], tuple] = {}
        self._slots: dict[tuple[int, int, int], list] = {}
        self._next: dict[tuple[int, int, int], int] = {}

    def _buffers(self, n_q: int, k: int):
        torch = self.torch
        key = (n_q, k)
        if key not in self._bufs:
            dev = torch.device("cuda", torch.cuda.current_device())
            local = torch.zeros(n_q * k * REC_BYTES, dtype=torch.uint8, device=dev)
            gathered = torch.zeros(self.world * n_q * k * REC_BYTES, dtype=torch.uint8, device=dev)
            merged = torch.zeros(n_q * k * REC_BYTES, dtype=torch.uint8, device=dev)
            self._bufs[key] = (local, gathered, merged)
        return self._bufs[key]

    def run(self, handle: int, n_q: int, k: int):
        """Score a resident batch on this shard and merge across ranks; returns
        the device tensor holding the merged records (uint8 view)."""
        import torch.distributed as dist

        local, gathered, merged = self._buffers(n_q, k)
        if self.world == 1:
            self.engine.run(handle, local.data_ptr())
            return local
        # The scoring kernels write this rank's top-k directly into the
        # all-gather send buffer; NCCL then exchanges world * n_q * k * 16 B.
        self.engine.run(handle, local.data_ptr())
        dist.all_gather_into_tensor(gathered, local)
        self.engine.merge_topk(gathered.data_ptr(), self.world, n_q, k, merged.data_ptr())
        return merged

    # ---- host batches in, host records out, two batches in flight

    def submit(self, host_batch, depth: int = 2):
        """Descriptor copy, scoring, all-gather, merge and the copy of the
        merged records into pinned host memory, all enqueued on the engine's
        stream; returns a ticket for collect().  Slots rotate, so at most
        `depth` batches may be outstanding."""
        import torch.distributed as dist

        torch = self.torch
        n_q, k = len(host_batch.queries), host_batch.limit
        key = (n_q, k, depth)
        if key not in self._slots:
            dev = torch.device("cuda", torch.cuda.current_device())
            mk = lambda n: torch.zeros(n * n_q * k * REC_BYTES, dtype=torch.uint8, device=dev)
            self._slots[key] = [dict(local=mk(1), gathered=mk(self.world), merged=mk(1),
                                     host=torch.zeros(n_q * k * REC_BYTES, dtype=torch.uint8).pin_memory(),
                                     done=torch.cuda.Event()) for _ in range(depth)]
            self._next[key] = 0
        sl = self._slots[key][self._next[key] % depth]
        self._next[key] += 1
        h = self.engine.search_begin(host_batch, sl["local"].data_ptr())
        out = sl["local"]
        if self.world > 1:
            dist.all_gather_into_tensor(sl["gathered"], sl["local"])
            self.engine.merge_topk(sl["gathered"].data_ptr(), self.world, n_q, k, sl["merged"].data_ptr())
            out = sl["merged"]
        sl["host"].copy_(out, non_blocking=True)
        sl["done"].record()
        return (h, sl, n_q, k)

    def collect(self, ticket) -> np.ndarray:
        h, sl, n_q, k = ticket
        sl["done"].synchronize()
        self.engine.search_end(h, discard=True)
        return sl["host"].numpy().view(REC_DTYPE).reshape(n_q, k)

    def to_host(self, recs_u8, n_q: int, k: int) -> np.ndarray:
        return recs_u8.cpu().numpy().view(REC_DTYPE).reshape(n_q, k)
