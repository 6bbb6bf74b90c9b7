"""GPU parity across tile boundaries: the stream kernel (stream.cuh) on a
multi-tile shard against the CPU oracle, and against the older tile kernel.

The C1 corpus of test_gpu_engine.py is a single 16384-document tile; the
threshold pruning between tiles, the sparse/dense epilogues, the overflow
rounds and the per-(query, tile) candidate cells only show with many tiles.
"""
import os

import numpy as np
import pytest

import _oracle
from _oracle import BM25, TFIDF, OP_AND, OP_OR, OP_ANDNOT, check_topk

pytestmark = pytest.mark.gpu

N_DOCS = 300_000      # 19 tiles
N_TERMS = 50_000


@pytest.fixture(scope="module")
def corpus():
    from nxsearch_b200 import tools

    return tools.Corpus.generate(N_DOCS, N_TERMS)


@pytest.fixture(scope="module")
def oracle(corpus):
    o = _oracle.OracleIndex(corpus)
    yield o
    o.close()


@pytest.fixture(scope="module")
def eng(corpus):
    from nxsearch_b200 import engine

    e = engine.Engine(0)
    e.load_corpus(corpus)
    yield e
    e.close()


def or_queries(corpus, n, seed_off=0):
    from nxsearch_b200 import tools

    qt = corpus.query_terms(4 * n, seed=tools.SEED + 11 + seed_off)
    out, pos = [], 0
    for i in range(n):
        nt = 1 + (i % 4)
        leaves = [int(t) for t in qt[pos:pos + nt]]
        pos += nt
        toks = []
        for t in reversed(leaves):
            if t not in toks:
                toks.append(t)
        slot = {t: s for s, t in enumerate(toks)}
        prog = [slot[leaves[0]]]
        for t in leaves[1:]:
            prog += [slot[t], OP_OR]
        out.append((toks, prog))
    return out


def bool_queries(corpus, n):
    from nxsearch_b200 import tools

    shapes = ["ab&", "ab|c&", "ab-", "ab|cd|&ef|-", "abc&&", "ab&c|"]
    qt = corpus.query_terms(6 * n, seed=tools.SEED + 23)
    out, pos = [], 0
    for i in range(n):
        shape = shapes[i % len(shapes)]
        nl = sum(ch.isalpha() for ch in shape)
        leaves = [int(t) for t in qt[pos:pos + nl]]
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


@pytest.mark.parametrize("algo,limit", [(BM25, 10), (TFIDF, 100), (BM25, 1), (TFIDF, 128)])
def test_or_queries_multi_tile(corpus, oracle, eng, algo, limit):
    from nxsearch_b200 import engine

    qs = or_queries(corpus, 256)
    counts, ids, scores = eng.search(engine.Batch.from_lists(algo, limit, qs))
    for i, (toks, prog) in enumerate(qs):
        all_ids, all_sc = oracle.search_all(algo, toks, prog)
        check_topk(ids[i, :counts[i]], scores[i, :counts[i]], all_ids, all_sc, limit,
                   exact_scores=(algo == TFIDF))
        if algo == TFIDF:
            order = np.lexsort((-all_ids.astype(np.int64), -all_sc.astype(np.float64)))[:limit]
            assert np.array_equal(all_ids[order], ids[i, :counts[i]])


@pytest.mark.parametrize("algo,limit", [(TFIDF, 100), (BM25, 10)])
def test_boolean_queries_multi_tile(corpus, oracle, eng, algo, limit):
    from nxsearch_b200 import engine

    qs = bool_queries(corpus, 240)
    counts, ids, scores = eng.search(engine.Batch.from_lists(algo, limit, qs))
    nonempty = 0
    for i, (toks, prog) in enumerate(qs):
        all_ids, all_sc = oracle.search_all(algo, toks, prog)
        nonempty += len(all_ids) > 0
        check_topk(ids[i, :counts[i]], scores[i, :counts[i]], all_ids, all_sc, limit,
                   exact_scores=(algo == TFIDF))
    assert nonempty > 120


def test_stream_kernel_equals_tile_kernel(corpus, eng):
    """Same batch through both scoring kernels: identical ids and score bits
    (same arithmetic, same total order), for pure-OR and boolean queries."""
    from nxsearch_b200 import engine

    old = os.environ.get("NXSB_KERNEL")
    os.environ["NXSB_KERNEL"] = "v2"
    try:
        e2 = engine.Engine(0)
    finally:
        if old is None:
            del os.environ["NXSB_KERNEL"]
        else:
            os.environ["NXSB_KERNEL"] = old
    e2.load_corpus(corpus)
    try:
        for algo, limit, qs in ((BM25, 10, or_queries(corpus, 1024, 1)),
                                (TFIDF, 100, bool_queries(corpus, 512))):
            b = engine.Batch.from_lists(algo, limit, qs)
            c1, i1, s1 = eng.search(b)
            c2, i2, s2 = e2.search(b)
            assert np.array_equal(c1, c2)
            assert np.array_equal(i1, i2)
            assert np.array_equal(s1.view(np.uint32), s2.view(np.uint32))
    finally:
        e2.close()


def test_limit_above_the_stream_kernel_falls_back(corpus, oracle, eng):
    """limit 129 is served by the tile kernel; its first 128 results are the
    limit-128 answer of the stream kernel."""
    from nxsearch_b200 import engine

    qs = or_queries(corpus, 64, 2)
    c_a, i_a, s_a = eng.search(engine.Batch.from_lists(BM25, 128, qs))
    c_b, i_b, s_b = eng.search(engine.Batch.from_lists(BM25, 129, qs))
    for q in range(len(qs)):
        n = min(int(c_a[q]), 128)
        assert np.array_equal(i_a[q, :n], i_b[q, :n])
        assert np.array_equal(s_a[q, :n], s_b[q, :n])


def test_shared_dense_prefixes_every_token_order(corpus, oracle, eng):
    """Shared dense prefixes (engine.cu fill_batch): queries whose dense terms
    lead the token list take their sum from a virtual query; the other orders
    stream their own columns.  Every shape must equal the tile kernel bit for
    bit (which adds token by token) and the oracle."""
    from nxsearch_b200 import engine

    df = np.asarray(corpus.term_df)
    dense = [int(t) + 1 for t in np.flatnonzero(df >= 0.6 * corpus.n_docs)][:5]
    assert len(dense) >= 3
    d1, d2, d3 = dense[:3]
    sp = [int(t) for t in corpus.query_terms(64, seed=4242) if int(t) not in dense][:12]
    s1, s2, s3, s4 = sp[:4]
    OR = OP_OR

    def q(*toks):
        prog = [0]
        for i in range(1, len(toks)):
            prog += [i, OR]
        return (list(toks), prog)

    shapes = [q(d1), q(d2), q(d1, d2), q(d2, d1), q(d1, d2, d3), q(*dense[:4]),
              q(d1, s1), q(s1, d1), q(s1, d1, s2), q(s1, d1, s2, s3), q(d1, d2, s1, s2), q(d1, s1, s2, s3),
              # not eligible: a dense term after two others, dense terms apart, one named twice
              q(s1, s2, d1), q(s1, s2, s3, d1), q(d1, s1, d2), q(s1, d1, d2), q(s1, d1, s2, d2), q(d1, d1, s1),
              q(s1), q(s1, s2, s3, s4)]
    if len(dense) == 5:
        shapes.append(q(*dense))                 # five dense terms: more than a prefix holds
    qs = shapes * 3 + or_queries(corpus, 64, 5)
    os.environ["NXSB_KERNEL"] = "v2"
    try:
        e2 = engine.Engine(0)
    finally:
        del os.environ["NXSB_KERNEL"]
    e2.load_corpus(corpus)
    try:
        for algo in (BM25, TFIDF):
            for limit in (1, 10, 128):
                b = engine.Batch.from_lists(algo, limit, qs)
                c1, i1, s1_ = eng.search(b)
                c2, i2, s2_ = e2.search(b)
                assert np.array_equal(c1, c2), (algo, limit)
                assert np.array_equal(i1, i2), (algo, limit)
                assert np.array_equal(s1_.view(np.uint32), s2_.view(np.uint32)), (algo, limit)
            # one-query batches: the prefix serves a single user
            for sh in shapes[:12]:
                b = engine.Batch.from_lists(algo, 10, [sh])
                r1, r2 = eng.search(b), e2.search(b)
                assert np.array_equal(r1[1], r2[1]) and np.array_equal(r1[2], r2[2])
        counts, ids, scores = eng.search(engine.Batch.from_lists(TFIDF, 10, shapes))
        for i, (toks, prog) in enumerate(shapes):
            if len(set(toks)) != len(toks):
                continue                          # the oracle merges a repeated term
            all_ids, all_sc = oracle.search_all(TFIDF, toks, prog)
            check_topk(ids[i, :counts[i]], scores[i, :counts[i]], all_ids, all_sc, 10, exact_scores=True)
    finally:
        e2.close()


def test_boolean_queries_with_base_columns(corpus, oracle, eng):
    """Boolean queries whose dense terms lead and cannot match on their own take
    those terms' scores and membership from the score columns; the others
    stream them.  All shapes against the tile kernel (bit-equal) and the oracle."""
    from nxsearch_b200 import engine

    df = np.asarray(corpus.term_df)
    dense = [int(t) + 1 for t in np.flatnonzero(df >= 0.6 * corpus.n_docs)][:5]
    d1, d2, d3 = dense[:3]
    sp = [int(t) for t in corpus.query_terms(64, seed=777) if int(t) not in dense][:8]
    s1, s2, s3, s4 = sp[:4]
    A, O, N = OP_AND, OP_OR, OP_ANDNOT
    shapes = [
        ([d1, s1], [0, 1, A]),                      # d AND s
        ([d1, s1], [1, 0, N]),                      # s AND NOT d
        ([d1, s1], [0, 1, N]),                      # d AND NOT s: dense alone matches -> streams
        ([d1, d2, s1], [0, 1, A, 2, A]),            # d AND d AND s
        ([d1, d2, s1], [0, 1, O, 2, A]),            # (d OR d) AND s
        ([d1, s1, s2], [0, 1, O, 2, A]),            # (d OR s) AND s
        ([s1, d1, s2], [0, 1, A, 2, O]),            # [s, d, s]: (s AND d) OR s
        ([s1, d1, s2], [0, 2, O, 1, A]),            # (s OR s) AND d  -- sparse required
        ([s1, s2, d1], [0, 1, A, 2, A]),            # dense last -> streams
        ([d1, s1, d2], [0, 1, A, 2, A]),            # dense apart -> streams
        ([d1, d2, d3, s1, s2], [0, 1, O, 2, O, 3, 4, O, A]),   # (d|d|d) AND (s|s)
        ([d1, s1, s2, s3], [1, 2, O, 3, O, 0, N]),  # (s|s|s) AND NOT d
        ([d1, s1], [0, 1, O]),                      # pure OR goes the other way (prefix)
        ([s1, s2], [0, 1, A]),
    ]
    qs = shapes * 4 + bool_queries(corpus, 60)
    os.environ["NXSB_KERNEL"] = "v2"
    try:
        e2 = engine.Engine(0)
    finally:
        del os.environ["NXSB_KERNEL"]
    e2.load_corpus(corpus)
    try:
        for algo in (TFIDF, BM25):
            for limit in (1, 10, 100):
                b = engine.Batch.from_lists(algo, limit, qs)
                c1, i1, sc1 = eng.search(b)
                c2, i2, sc2 = e2.search(b)
                assert np.array_equal(c1, c2), (algo, limit)
                assert np.array_equal(i1, i2), (algo, limit)
                assert np.array_equal(sc1.view(np.uint32), sc2.view(np.uint32)), (algo, limit)
        counts, ids, scores = eng.search(engine.Batch.from_lists(TFIDF, 100, shapes))
        matched = 0
        for i, (toks, prog) in enumerate(shapes):
            all_ids, all_sc = oracle.search_all(TFIDF, toks, prog)
            matched += len(all_ids) > 0
            check_topk(ids[i, :counts[i]], scores[i, :counts[i]], all_ids, all_sc, 100, exact_scores=True)
        assert matched >= 10
    finally:
        e2.close()
