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


def _devices(n):
    """n device ordinals: every visible device first, wrapping on a small box."""
    from nxsearch_b200 import engine

    ndev = engine.device_count()
    return [d % ndev for d in range(max(n, ndev))]


@pytest.mark.parametrize("n_shards", [2, 5])
def test_one_sharded_engine_equals_one_engine(c1_corpus, n_shards):
    """nxsb_engine_create_sharded: one engine object, a range of the documents per
    device, every device scores the whole batch, lists merged on the first
    device.  Counts, ids and scores equal the single-engine answer bit for bit,
    from pairs and from dtmap bytes, synchronously and with four in flight."""
    from nxsearch_b200 import engine
    from test_gpu_engine import c1_queries
    from test_gpu_segments import assert_same, boolean_queries

    c = c1_corpus
    devs = _devices(n_shards)
    whole = engine.Engine(0)
    whole.load_corpus(c)
    sh = engine.Engine(devices=devs, layout="shards")
    # no df handed over: the dispatcher sums the shards' own (the all-reduce of dist.py)
    sh.load_corpus(c)
    assert np.array_equal(sh.get_df(c.n_terms), whole.get_df(c.n_terms))
    qs = c1_queries(c, 300) + boolean_queries(c, 90)
    for algo in (BM25, TFIDF):
        for k in (1, 10, 100, 200):
            batch = engine.Batch.from_lists(algo, k, qs)
            want = whole.search(batch)
            assert_same(sh.search(batch), want)
            hs = [sh.search_begin(batch) for _ in range(4)]
            for h in reversed(hs):
                assert_same(sh.search_end(h, len(qs), k), want)
    # a batch of one query, and a query nobody matches
    one = engine.Batch.from_lists(BM25, 10, qs[:1])
    assert_same(sh.search(one), whole.search(one))
    # with whole-index df given, as the host library does
    sh.load_corpus(c, df=c.term_df)
    batch = engine.Batch.from_lists(BM25, 10, qs)
    assert_same(sh.search(batch), whole.search(batch))
    with pytest.raises(RuntimeError, match="replicated|sharded"):
        sh.upload(batch)
    sh.close()
    whole.close()


def test_sharded_engine_segments_and_removals(c1_corpus):
    """Delta segments go whole to one device each (round robin) and removal
    notes to whoever holds the document: the sharded engine still answers like
    one image rebuilt from the live documents."""
    from nxsearch_b200 import engine
    from test_gpu_engine import c1_queries
    from test_gpu_segments import assert_same, boolean_queries, stats, subset

    c = c1_corpus
    N = c.n_docs
    rng = np.random.default_rng(11)
    # ids of the base and of four deltas interleave: ties are ordered by id, not by device
    part = rng.integers(0, 20, N)
    seg_docs = [np.flatnonzero(part < 14)] + [np.flatnonzero(part == 14 + i) for i in range(4)]
    seg_docs.append(np.flatnonzero(part >= 18))
    dead = [rng.choice(s, size=min(k, len(s)), replace=False) for s, k in zip(seg_docs, (50, 7, 0, 3, 1, 9))]
    dead_all = np.concatenate(dead)
    live = np.setdiff1d(np.arange(N), dead_all)
    df, tokens, ndocs = stats(c, live)
    truth = engine.Engine(0)
    truth.load_docs(*subset(c, live), c.n_terms, tokens, ndocs, df)

    sh = engine.Engine(devices=_devices(3), layout="shards")
    sh.load_docs(*subset(c, seg_docs[0]), c.n_terms, c.token_count, c.doc_count, c.term_df)
    for g, s in enumerate(seg_docs[1:], 1):
        sh.load_docs(*subset(c, s), c.n_terms, c.token_count, c.doc_count, c.term_df, segment=True)
        assert sh.segment_count() == g
    for g, d in enumerate(dead):
        sh.set_dead(g, c.doc_ids[d])
    sh.set_global_stats(df, tokens, ndocs)
    qs = c1_queries(c, 300) + boolean_queries(c, 90)
    dead_ids = set(int(x) for x in c.doc_ids[dead_all])
    for algo in (BM25, TFIDF):
        for k in (1, 10, 100):
            batch = engine.Batch.from_lists(algo, k, qs)
            want = truth.search(batch)
            got = sh.search(batch)
            assert_same(got, want)
            assert not dead_ids.intersection(int(x) for x in got[1].ravel())
    # notes are replaced, not accumulated: clearing them brings the documents back
    for g in range(len(dead)):
        sh.set_dead(g, [])
    df_all, tok_all, n_all = stats(c, np.arange(N))
    sh.set_global_stats(df_all, tok_all, n_all)
    truth.load_corpus(c)
    batch = engine.Batch.from_lists(BM25, 10, qs)
    assert_same(sh.search(batch), truth.search(batch))
    sh.segments_drop()
    assert sh.segment_count() == 0
    sh.close()
    truth.close()


def test_sharded_engine_with_fewer_documents_than_devices():
    """Three documents over five ranges: the empty ranges answer nothing, the
    merge still returns what one engine returns; then an index with no
    documents at all, and a delta segment next to empty base ranges."""
    from nxsearch_b200 import engine
    from test_gpu_segments import assert_same

    ids, lens = [7, 9, 12], [13, 5, 4]
    off, pairs = [0, 2, 3, 5], [1, 10, 2, 3, 2, 5, 1, 1, 3, 3]
    qs = [([1], [0]), ([2, 1], [1, 0, -3]), ([3], [0]), ([2, 1], [1, 0, -2]), ([3, 2], [1, 0, -4])]
    one = engine.Engine(0)
    sh = engine.Engine(devices=_devices(5), layout="shards")
    for e in (one, sh):
        e.load_docs(ids, lens, off, pairs, 3, sum(lens), 3)
    for algo in (BM25, TFIDF):
        for k in (1, 2, 10):
            batch = engine.Batch.from_lists(algo, k, qs)
            assert_same(sh.search(batch), one.search(batch))
    # a delta segment and a removal; statistics move with them
    df = np.array([3, 2, 2], dtype=np.uint32)
    for e in (one, sh):
        e.load_docs([20], [6], [0, 2], [1, 2, 3, 4], 3, sum(lens) + 6, 4, df, segment=True)
        e.set_dead(0, [9])
        e.set_global_stats(np.array([3, 1, 2], dtype=np.uint32), sum(lens) + 6 - 5, 3)
    batch = engine.Batch.from_lists(BM25, 10, qs)
    got = sh.search(batch)
    assert_same(got, one.search(batch))
    assert 9 not in got[1] and 20 in got[1]
    for e in (one, sh):
        e.load_docs([], [], [0], [], 3, 0, 0)
    assert list(sh.search(batch)[0]) == [0] * len(qs)
    one.close()
    sh.close()
