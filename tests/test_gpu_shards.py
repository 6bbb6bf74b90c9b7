"""Shard invariance: scoring G document shards with global
statistics and merging the per-shard top-k equals scoring the whole index."""
import numpy as np
import pytest

from _oracle import BM25, TFIDF, check_topk

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_sharded_equals_whole(c1_corpus, c1_oracle, n_shards):
    import torch
    from nxsearch_b200 import dist as nxdist, engine
    from test_gpu_engine import c1_queries

    qs = c1_queries(c1_corpus, 400)
    whole = engine.Engine(0)
    whole.load_corpus(c1_corpus)
    shards = []
    for g in range(n_shards):
        lo, hi = nxdist.shard_range(c1_corpus.n_docs, g, n_shards)
        e = engine.Engine(0)
        e.load_corpus(c1_corpus, lo=lo, hi=hi, df=c1_corpus.term_df,
                      token_count=c1_corpus.token_count, doc_count=c1_corpus.doc_count)
        shards.append(e)
    for algo in (BM25, TFIDF):
        for k in (10, 100):
            batch = engine.Batch.from_lists(algo, k, qs)
            counts, ids, scores = whole.search(batch)
            gathered = torch.zeros(n_shards * len(qs) * k * nxdist.REC_BYTES, dtype=torch.uint8, device="cuda")
            per = len(qs) * k * nxdist.REC_BYTES
            for g, e in enumerate(shards):
                h = e.upload(batch)
                e.run(h, gathered.data_ptr() + g * per)
                e.sync()
                e.release(h)
            merged = torch.zeros(per, dtype=torch.uint8, device="cuda")
            shards[0].merge_topk(gathered.data_ptr(), n_shards, len(qs), k, merged.data_ptr())
            shards[0].sync()
            recs = merged.cpu().numpy().view(nxdist.REC_DTYPE).reshape(len(qs), k)
            host = nxdist.merge_topk_host(
                gathered.cpu().numpy().view(nxdist.REC_DTYPE).reshape(n_shards, len(qs), k), k)
            for i in range(len(qs)):
                n = int(recs[i]["valid"].sum())
                assert n == counts[i]
                # bit-for-bit the single-GPU answer, and the host model of the merge
                assert np.array_equal(recs[i]["doc_id"][:n], ids[i, :n])
                assert np.array_equal(recs[i]["score"][:n], scores[i, :n])
                assert np.array_equal(recs[i]["doc_id"], host[i]["doc_id"])
    for e in shards:
        e.close()
    whole.close()
