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


def test_submit_collect_pipeline_equals_search(c1_corpus):
    """ShardedSearcher.submit/collect (nxsb_engine_search_begin_dev/_end, two
    batches in flight) returns the records of the synchronous search."""
    import torch
    from nxsearch_b200 import dist as nxdist, engine
    from test_gpu_engine import c1_queries

    e = engine.Engine(0)
    e.load_corpus(c1_corpus)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    e.set_stream(stream.cuda_stream)
    s = nxdist.ShardedSearcher(e, 0, 1)
    qs = c1_queries(c1_corpus, 600)
    batches = [engine.Batch.from_lists(BM25, 10, qs[i * 200:(i + 1) * 200]) for i in range(3)]
    want = [e.search(b) for b in batches]
    order = [0, 1, 2, 0, 2, 1, 1]
    tickets = [s.submit(batches[order[0]])]
    got = []
    for j in range(len(order)):
        if j + 1 < len(order):
            tickets.append(s.submit(batches[order[j + 1]]))
        got.append(s.collect(tickets[j]).copy())
    for j, b in enumerate(order):
        counts, ids, scores = want[b]
        for i in range(200):
            n = int(got[j][i]["valid"].sum())
            assert n == counts[i]
            assert np.array_equal(got[j][i]["doc_id"][:n], ids[i, :n])
            assert np.array_equal(got[j][i]["score"][:n], scores[i, :n])
    # the host-array flavour of the same two calls
    h0, h1 = e.search_begin(batches[0]), e.search_begin(batches[1])
    r1, r0 = e.search_end(h1, 200, 10), e.search_end(h0, 200, 10)
    for r, w in ((r0, want[0]), (r1, want[1])):
        assert all(np.array_equal(a, b) for a, b in zip(r, w))
    torch.cuda.set_stream(torch.cuda.default_stream())
    e.close()
