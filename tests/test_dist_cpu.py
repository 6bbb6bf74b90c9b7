"""Host logic of the sharded path on CPU: two gloo ranks."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

from nxsearch_b200 import dist as nxdist

ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["NXSB_ROOT"])
import numpy as np, torch.distributed as dist
from nxsearch_b200 import dist as nxdist, tools
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, V = 6000, 5000
lo, hi = nxdist.shard_range(N, rank, world)
shard = tools.Corpus.generate(hi - lo, V, first_doc=lo)
df, tokens, docs = nxdist.allreduce_stats(np.asarray(shard.term_df), shard.token_count, shard.n_docs)
whole = tools.Corpus.generate(N, V)
assert docs == N and tokens == whole.token_count
assert np.array_equal(df, whole.term_df)
# a shard is exactly the corresponding slice of the whole corpus
assert np.array_equal(shard.doc_ids, whole.doc_ids[lo:hi])
s, e = int(whole.doc_off[lo]), int(whole.doc_off[hi])
assert np.array_equal(shard.pairs, whole.pairs[2 * s: 2 * e])
# every rank draws the same query stream from the global df
q = tools.query_terms(V, df, 64)
ref = [None] * world
dist.all_gather_object(ref, q.tolist())
assert all(r == ref[0] for r in ref)
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_shards_and_stats(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, NXSB_ROOT=str(ROOT))
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
        capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_shard_ranges_partition_the_index():
    for n in (0, 1, 7, 10_000_000):
        for w in (1, 2, 3, 8):
            edges = [nxdist.shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))


def test_merge_model_orders_ties_by_shard_then_position():
    recs = np.zeros((2, 1, 3), dtype=nxdist.REC_DTYPE)
    recs[0, 0] = [(5, 2.0, 1), (4, 1.0, 1), (0, 0, 0)]
    recs[1, 0] = [(9, 2.0, 1), (8, 2.0, 1), (7, 0.5, 1)]
    out = nxdist.merge_topk_host(recs, 4)
    assert out[0]["doc_id"].tolist() == [9, 8, 5, 4]
