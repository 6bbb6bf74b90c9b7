"""Incremental image refresh: a base shard + delta segments +
removal notes must answer exactly like one image rebuilt from the live
documents -- the reference applies every appended dtmap block to its in-memory
index on the next search (ref src/index/dtmap.c:357-441), so there is no
"stale" state to be compatible with, only the rebuilt one."""
import numpy as np
import pytest

from _oracle import BM25, TFIDF

pytestmark = pytest.mark.gpu


def subset(corpus, idx):
    """(ids, lens, doc_off, pairs) of the documents idx (ascending) of a tools.Corpus."""
    idx = np.asarray(idx, dtype=np.int64)
    n = (corpus.doc_off[idx + 1] - corpus.doc_off[idx]).astype(np.int64)
    off = np.zeros(len(idx) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(n)
    pairs = np.concatenate([corpus.pairs[2 * int(corpus.doc_off[i]): 2 * int(corpus.doc_off[i + 1])]
                            for i in idx]) if len(idx) else np.zeros(0, np.uint32)
    return corpus.doc_ids[idx], corpus.doc_len[idx], off, pairs


def stats(corpus, live):
    """(df, token_count, doc_count) over the live documents."""
    _, lens, _, pairs = subset(corpus, live)
    df = np.bincount(pairs[0::2].astype(np.int64) - 1, minlength=corpus.n_terms).astype(np.uint32)
    return df, int(lens.sum()), len(live)


def boolean_queries(corpus, n):
    from _oracle import OP_AND, OP_OR, OP_ANDNOT
    qt = [int(t) for t in corpus.query_terms(4 * n, seed=77)]
    out = []
    for i in range(n):
        a, b, c, d = qt[4 * i: 4 * i + 4]
        if len({a, b, c, d}) < 4:
            continue
        if i % 3 == 0:
            out.append(([b, a], [1, 0, OP_AND]))
        elif i % 3 == 1:
            out.append(([c, b, a], [2, 1, OP_OR, 0, OP_AND]))
        else:
            out.append(([b, a], [1, 0, OP_ANDNOT]))
    return out


def assert_same(a, b):
    (ca, ia, sa), (cb, ib, sb) = a, b
    assert np.array_equal(ca, cb)
    for q in range(len(ca)):
        n = int(ca[q])
        assert np.array_equal(ia[q, :n], ib[q, :n]), q
        assert np.array_equal(sa[q, :n], sb[q, :n]), q


@pytest.mark.parametrize("layout", ["interleaved", "appended"])
def test_segments_and_removals_equal_a_rebuild(c1_corpus, layout):
    from nxsearch_b200 import engine
    from test_gpu_engine import c1_queries

    c = c1_corpus
    N = c.n_docs
    rng = np.random.default_rng(5)
    if layout == "interleaved":
        # ids of the three segments interleave: ties must be ordered by id, not by segment
        part = rng.integers(0, 10, N)
        seg_docs = [np.flatnonzero(part < 7), np.flatnonzero((part >= 7) & (part < 9)), np.flatnonzero(part == 9)]
    else:
        seg_docs = [np.arange(0, 7000), np.arange(7000, 9990), np.arange(9990, N)]
    dead = [rng.choice(s, size=min(k, len(s)), replace=False) for s, k in zip(seg_docs, (40, 9, 2))]
    dead_all = np.concatenate(dead)
    live = np.setdiff1d(np.arange(N), dead_all)
    df, tokens, ndocs = stats(c, live)

    truth = engine.Engine(0)
    truth.load_docs(*subset(c, live), c.n_terms, tokens, ndocs, df)

    seg = engine.Engine(0)
    # built before the removals happened: statistics at load time are stale on purpose
    seg.load_docs(*subset(c, seg_docs[0]), c.n_terms, c.token_count, c.doc_count, c.term_df)
    for s in seg_docs[1:]:
        seg.load_docs(*subset(c, s), c.n_terms, c.token_count, c.doc_count, c.term_df, segment=True)
    assert seg.segment_count() == 2
    for g, d in enumerate(dead):
        seg.set_dead(g, c.doc_ids[d])
    seg.set_global_stats(df, tokens, ndocs)

    qs = c1_queries(c, 400) + boolean_queries(c, 120)
    for algo in (BM25, TFIDF):
        for k in (1, 10, 100):
            batch = engine.Batch.from_lists(algo, k, qs)
            want = truth.search(batch)
            assert_same(seg.search(batch), want)
            # the pipelined entry points take the same path
            h1 = seg.search_begin(batch)
            h2 = seg.search_begin(batch)
            assert_same(seg.search_end(h1, len(qs), k), want)
            assert_same(seg.search_end(h2, len(qs), k), want)
            dead_ids = set(int(x) for x in c.doc_ids[dead_all])
            assert not dead_ids.intersection(int(x) for x in want[1].ravel())

    # resident batches are refused on a segmented image, loudly
    with pytest.raises(RuntimeError, match="segmented"):
        h = seg.upload(engine.Batch.from_lists(BM25, 10, qs[:4]))
        seg.run(h)

    # dropping the segments and notes gives back the plain base image
    seg.segments_drop()
    assert seg.segment_count() == 0
    base_only = engine.Engine(0)
    base_only.load_docs(*subset(c, seg_docs[0]), c.n_terms, tokens, ndocs, df)
    batch = engine.Batch.from_lists(BM25, 10, qs)
    assert_same(seg.search(batch), base_only.search(batch))
    for e in (truth, seg, base_only):
        e.close()


def test_segment_with_a_longer_vocabulary(c1_corpus):
    """Terms only grow: a delta segment may know terms the base was built without."""
    from nxsearch_b200 import engine

    c = c1_corpus
    v_base = 30_000                       # the base image was built when the vocabulary ended here
    top = np.array([c.pairs[2 * int(c.doc_off[i + 1]) - 2] for i in range(c.n_docs)])   # pairs ascend by term
    base_docs = np.flatnonzero(top <= v_base)
    rest = np.flatnonzero(top > v_base)
    assert len(base_docs) > 100 and len(rest) > 100
    df, tokens, ndocs = stats(c, np.arange(c.n_docs))
    seg = engine.Engine(0)
    seg.load_docs(*subset(c, base_docs), v_base, tokens, ndocs, df[:v_base])
    seg.load_docs(*subset(c, rest), c.n_terms, tokens, ndocs, df, segment=True)
    seg.set_global_stats(df, tokens, ndocs)
    truth = engine.Engine(0)
    truth.load_corpus(c)
    # queries over terms the base image has never heard of, and ordinary ones
    _, _, _, rp = subset(c, rest)
    new_terms = [int(t) for t in np.unique(rp[0::2]) if t > v_base][:64]
    assert new_terms
    common = [int(t) for t in c.query_terms(64)]
    qs = [([t], [0]) for t in new_terms] + [([u, t], [1, 0, -3]) for t, u in zip(new_terms, common) if t != u]
    for algo in (BM25, TFIDF):
        batch = engine.Batch.from_lists(algo, 10, qs)
        assert_same(seg.search(batch), truth.search(batch))
    seg.close()
    truth.close()


# --------------------------------------------------------------------------
# through the public C API: nxs_index_add / nxs_index_remove between searches


def _doc_text(corpus, i):
    lo, hi = int(corpus.doc_off[i]), int(corpus.doc_off[i + 1])
    words = []
    for j in range(lo, hi):
        words += [corpus.term(int(corpus.pairs[2 * j]))] * int(corpus.pairs[2 * j + 1])
    return " ".join(words)


def _api_queries(corpus, n):
    qt = [corpus.term(int(t)) for t in corpus.query_terms(4 * n, seed=99)]
    qs = []
    for i in range(n):
        a, b, c, d = qt[4 * i: 4 * i + 4]
        qs.append([a, f"{a} OR {b}", f"{a} OR {b} OR {c} OR {d}", f"({a} OR {b}) AND {c}",
                   f"{a} AND NOT {b}", f"{a} AND {b}"][i % 6])
    return qs


def _same_as_fresh_open(nxs, idx, name, queries):
    """The long-lived handle (incrementally refreshed image) against a fresh
    instance over the same files (image built from scratch): bit-equal."""
    from nxsearch_b200 import capi
    other = capi.Nxs(nxs.base)
    fresh = other.open_index(name)
    try:
        for algo in ("BM25", "TF-IDF"):
            for limit in (10, 100):
                got = idx.search_batch(queries, limit=limit, algo=algo)
                want = fresh.search_batch(queries, limit=limit, algo=algo)
                assert got == want, (algo, limit)
        assert fresh.image_stats()["full_builds"] == 1 and fresh.image_stats()["segments"] == 0
    finally:
        fresh.close()
        other.close()


def test_c_api_incremental_refresh(c1_corpus):
    import shutil
    import tempfile
    import _oracle
    from nxsearch_b200 import capi, tools
    from test_gpu_capi import same_results

    c = c1_corpus
    n_base = 9000
    base = tempfile.mkdtemp(prefix="nxsb_seg_")
    nxs = capi.Nxs(base)
    nxs.base = base
    try:
        nxs.create_index("inc", filters=["normalizer"]).close()
        # 9000 documents arrive as files (bulk writer), the rest through nxs_index_add
        part = tools.Corpus.generate(n_base, c.n_terms)
        part.write(f"{base}/data/inc/nxsterms", f"{base}/data/inc/nxsdtmap")
        idx = nxs.open_index("inc")
        queries = _api_queries(c, 120)
        idx.search_batch(queries[:4], limit=10)
        assert idx.image_stats() == dict(full_builds=1, delta_builds=0, consolidations=0, segments=0,
                                         dead_noted=0, live=n_base, pending=0)

        nxt = n_base
        # three rounds of additions, each picked up by the next search as one delta segment
        for rnd in range(3):
            for _ in range(40):
                idx.add(int(c.doc_ids[nxt]), _doc_text(c, nxt))
                nxt += 1
            assert idx.image_stats()["pending"] == 40
            _same_as_fresh_open(nxs, idx, "inc", queries)
            st = idx.image_stats()
            assert (st["full_builds"], st["delta_builds"], st["segments"], st["pending"]) == (1, rnd + 1, rnd + 1, 0)

        # removals from the base and from a delta: noted, not rebuilt
        for d in (5, 77, 4321, n_base + 3, n_base + 41):
            idx.remove(int(c.doc_ids[d]))
        _same_as_fresh_open(nxs, idx, "inc", queries)
        st = idx.image_stats()
        assert (st["full_builds"], st["dead_noted"], st["live"]) == (1, 5, nxt - 5)
        # an id removed and added again lives in a newer segment than its dead twin
        idx.add(int(c.doc_ids[77]), _doc_text(c, 77))
        _same_as_fresh_open(nxs, idx, "inc", queries)

        # another process appends and removes; this handle follows
        other = capi.Nxs(base)
        oidx = other.open_index("inc")
        for _ in range(25):
            oidx.add(int(c.doc_ids[nxt]), _doc_text(c, nxt))
            nxt += 1
        oidx.remove(int(c.doc_ids[100]))
        oidx.close()
        other.close()
        _same_as_fresh_open(nxs, idx, "inc", queries)
        assert idx.image_stats()["full_builds"] == 1

        # more refreshes than delta slots: the deltas are consolidated, the base stays
        for rnd in range(8):
            idx.add(int(c.doc_ids[nxt]), _doc_text(c, nxt))
            nxt += 1
            idx.search(queries[0])
        st = idx.image_stats()
        assert st["full_builds"] == 1 and st["consolidations"] >= 1 and st["segments"] <= 8
        _same_as_fresh_open(nxs, idx, "inc", queries)

        # too many dead documents in the base: rebuilt in full
        for d in range(200, 270):
            idx.remove(int(c.doc_ids[d]))
        _same_as_fresh_open(nxs, idx, "inc", queries)
        st = idx.image_stats()
        assert (st["full_builds"], st["segments"], st["dead_noted"]) == (2, 0, 0)

        # and the compiled reference reads the same files the same way
        ref = _oracle.ref()
        if ref is not None:
            rn = capi.Nxs(base, lib=ref)
            ridx = rn.open_index("inc")
            for q in queries[:40]:
                for algo in ("BM25", "TF-IDF"):
                    same_results(idx.search(q, limit=20, algo=algo), ridx.search(q, limit=20, algo=algo))
            ridx.close()
            rn.close()
        idx.close()
    finally:
        nxs.close()
        shutil.rmtree(base, ignore_errors=True)
