"""GPU parity of the block-max scorer (bmw.cuh).

Pruning must not change a single bit: the same batch with pruning on and off
(off = every posting streamed, stream.cuh) gives identical ids and score bits,
both agree with the CPU oracle, and the kernel really skipped most blocks.
"""
import os

import numpy as np
import pytest

import _oracle
from _oracle import BM25, TFIDF, OP_OR, check_topk

pytestmark = pytest.mark.gpu

N_DOCS = 300_000      # 19 tiles, 9375 blocks of 32 documents (two chunks)
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


def or_queries(corpus, n, seed_off=0, max_terms=4):
    from nxsearch_b200 import tools

    qt = corpus.query_terms(max_terms * n, seed=tools.SEED + 101 + seed_off)
    out, pos = [], 0
    for i in range(n):
        nt = 1 + (i % max_terms)
        leaves = [int(t) for t in qt[pos:pos + nt]]
        pos += nt
        toks = []
        for t in reversed(leaves):
            if t not in toks:
                toks.append(t)
        out.append((toks, None))
    return out


def both_ways(eng, batch):
    """(pruned, exhaustive) results of one batch on the same engine."""
    was = eng.set_pruning(True)
    eng.pruning_stats(reset=True)
    pruned = eng.search(batch)
    stats = eng.pruning_stats()
    eng.set_pruning(False)
    full = eng.search(batch)
    eng.set_pruning(was)
    return pruned, full, stats


def assert_identical(a, b):
    (ca, ia, sa), (cb, ib, sb) = a, b
    assert np.array_equal(ca, cb)
    for q in range(len(ca)):
        n = ca[q]
        assert np.array_equal(ia[q, :n], ib[q, :n]), f"query {q}: ids differ"
        assert np.array_equal(sa[q, :n].view(np.uint32), sb[q, :n].view(np.uint32)), f"query {q}: score bits differ"


@pytest.mark.parametrize("algo,limit", [(BM25, 10), (TFIDF, 10), (BM25, 1), (TFIDF, 100), (BM25, 128)])
def test_pruned_equals_exhaustive_and_oracle(corpus, oracle, eng, algo, limit):
    from nxsearch_b200 import engine

    qs = or_queries(corpus, 512)
    pruned, full, stats = both_ways(eng, engine.Batch.from_lists(algo, limit, qs))
    assert_identical(pruned, full)
    counts, ids, scores = pruned
    for i, (toks, _) in enumerate(qs[:192]):
        all_ids, all_sc = oracle.search_all(algo, toks)
        check_topk(ids[i, :counts[i]], scores[i, :counts[i]], all_ids, all_sc, limit,
                   exact_scores=(algo == TFIDF))
    assert stats["items"] > 0 and stats["blocks_scored"] > 0, "the block-max kernel did not run"
    if limit <= 10:
        # The point of the exercise: far fewer postings scored than the batch names.
        named = sum(int(corpus.term_df[t - 1]) for toks, _ in qs for t in toks)
        assert stats["postings_scored"] < named / 4, (stats, named)


def test_many_terms_rare_terms_and_odd_ids(corpus, oracle, eng):
    """Up to 32 tokens, lists too short for block arrays, the same term twice
    (two strings that fuzzy-resolve to one term score twice),
    ids that are not terms, and limits beyond the match count."""
    from nxsearch_b200 import engine

    df = np.asarray(corpus.term_df)
    order = np.argsort(df, kind="stable")
    order = order[df[order] >= 1]
    rare = [int(t) + 1 for t in order[:64]]                       # the shortest lists
    mid = [int(t) + 1 for t in np.nonzero((df >= 200) & (df <= 2000))[0][:64]]
    assert len(rare) == 64 and len(mid) == 64
    big = or_queries(corpus, 8, seed_off=7, max_terms=32)
    qs = [
        ([rare[0]], None), ([rare[1], rare[2], rare[3]], None), ([mid[0], rare[4]], None),
        ([1, mid[1], rare[5]], None), ([2, 2], None), ([mid[2], mid[2], 3], None),
        ([0, 5], None), ([corpus.n_terms + 7, mid[3]], None), ([corpus.n_terms + 1], None),
        ([1], None), ([1, 2, 3, 4, 5], None),
    ] + big
    for algo in (BM25, TFIDF):
        for limit in (10, 128):
            pruned, full, _ = both_ways(eng, engine.Batch.from_lists(algo, limit, qs))
            assert_identical(pruned, full)
            counts, ids, scores = pruned
            for i, (toks, _) in enumerate(qs):
                if any(t == 0 or t > corpus.n_terms for t in toks) or len(set(toks)) != len(toks):
                    continue        # the oracle takes real, distinct terms
                all_ids, all_sc = oracle.search_all(algo, toks)
                check_topk(ids[i, :counts[i]], scores[i, :counts[i]], all_ids, all_sc, limit,
                           exact_scores=(algo == TFIDF))


@pytest.mark.parametrize("shift", [6, 7, 8])
def test_other_block_sizes(corpus, eng, shift):
    from nxsearch_b200 import engine

    old = os.environ.get("NXSB_BMW_SHIFT")
    os.environ["NXSB_BMW_SHIFT"] = str(shift)
    try:
        e2 = engine.Engine(0)
    finally:
        if old is None:
            del os.environ["NXSB_BMW_SHIFT"]
        else:
            os.environ["NXSB_BMW_SHIFT"] = old
    try:
        e2.load_corpus(corpus)
        qs = or_queries(corpus, 256, seed_off=shift)
        for algo, limit in ((BM25, 10), (TFIDF, 37)):
            batch = engine.Batch.from_lists(algo, limit, qs)
            e2.pruning_stats(reset=True)
            got = e2.search(batch)
            assert e2.pruning_stats()["blocks_scored"] > 0
            eng.set_pruning(False)
            want = eng.search(batch)
            eng.set_pruning(True)
            assert_identical(got, want)
    finally:
        e2.close()


@pytest.mark.parametrize("n_docs", [1, 50, 64, 65, 5000])
def test_tiny_shards(n_docs):
    """Shards smaller than a block, a tile, a chunk."""
    from nxsearch_b200 import engine, tools

    corpus = tools.Corpus.generate(n_docs, 3000)
    ora = _oracle.OracleIndex(corpus)
    e = engine.Engine(0)
    try:
        e.load_corpus(corpus)
        qt = [int(t) for t in corpus.query_terms(120)]
        qs = [([qt[3 * i], qt[3 * i + 1], qt[3 * i + 2]] if len({qt[3 * i], qt[3 * i + 1], qt[3 * i + 2]}) == 3
               else [qt[3 * i]], None) for i in range(40)]
        for algo in (BM25, TFIDF):
            pruned, full, _ = both_ways(e, engine.Batch.from_lists(algo, 10, qs))
            assert_identical(pruned, full)
            counts, ids, scores = pruned
            for i, (toks, _) in enumerate(qs):
                all_ids, all_sc = ora.search_all(algo, toks)
                check_topk(ids[i, :counts[i]], scores[i, :counts[i]], all_ids, all_sc, 10,
                           exact_scores=(algo == TFIDF))
    finally:
        e.close()
        ora.close()


def test_statistics_refresh_moves_the_block_maxima(corpus, eng):
    """set_global_stats changes K1 (BM25) and every idf: the bounds must follow."""
    from nxsearch_b200 import engine

    qs = or_queries(corpus, 128, seed_off=3)
    batch = engine.Batch.from_lists(BM25, 10, qs)
    df = np.asarray(corpus.term_df).astype(np.uint32)
    try:
        # pretend the index is 3x larger with much longer documents elsewhere
        eng.set_global_stats(df * 2, corpus.token_count * 9, corpus.doc_count * 3)
        pruned, full, _ = both_ways(eng, batch)
        assert_identical(pruned, full)
    finally:
        eng.set_global_stats(df, corpus.token_count, corpus.doc_count)
    pruned, full, _ = both_ways(eng, batch)
    assert_identical(pruned, full)


def test_no_allocation_in_steady_state(corpus, eng):
    """cudaFree on the search path stalls the device.  After a few
    batches of one shape the allocation counter must not move, whatever the
    batches contain (pruned, exhaustive, boolean)."""
    from nxsearch_b200 import engine
    from test_gpu_stream import bool_queries

    def batches(seed0):
        for s in range(12):
            qs = or_queries(corpus, 256, seed_off=seed0 + s)
            yield engine.Batch.from_lists(BM25, 10, qs)
        yield engine.Batch.from_lists(TFIDF, 10, bool_queries(corpus, 256))

    for pruning in (True, False):
        eng.set_pruning(pruning)
        for b in batches(40):                       # warm-up: arenas reach their size
            eng.search_end(eng.search_begin(b), len(b.queries), b.limit)
        before = engine.alloc_events()
        for b in batches(60):
            eng.search_end(eng.search_begin(b), len(b.queries), b.limit)
        assert engine.alloc_events() == before, "device memory was (re)allocated on the search path"
    eng.set_pruning(True)


def test_kth_weight_ladder(corpus, eng):
    """Threshold priming: the image's k-th largest weight per term (one warp
    per list, long lists in parts) against a sort of the list on the host --
    TF-IDF weights bit for bit, BM25 weights within the kernel's 1e-6."""
    df = np.asarray(corpus.term_df)
    terms = np.asarray(corpus.pairs[0::2])
    counts = np.asarray(corpus.pairs[1::2]).astype(np.int64)
    doc_of = np.repeat(np.arange(corpus.n_docs), np.diff(np.asarray(corpus.doc_off)).astype(np.int64))
    dl = np.asarray(corpus.doc_len).astype(np.float64)
    by_df = np.argsort(-df.astype(np.int64), kind="stable")
    picks = [int(by_df[0]) + 1, int(by_df[1]) + 1, int(by_df[7]) + 1]               # several parts each
    picks += [int(t) + 1 for t in np.nonzero((df > 300) & (df < 5000))[0][:4]]       # one buffer cut or more
    picks += [int(t) + 1 for t in np.nonzero((df >= 1) & (df < 128))[0][:4]]         # shorter than the ladder
    picks += [int(np.nonzero(df == 0)[0][0]) + 1] if (df == 0).any() else []
    assert df[picks[0] - 1] > 32768 * 2
    kth_t = eng.term_kth(TFIDF, picks)
    kth_b = eng.term_kth(BM25, picks)
    k = np.float64(np.float32(1.2))
    adl = float(corpus.token_count // corpus.doc_count)
    K0, K1 = np.float64(np.float32(k * 0.25)), np.float64(np.float32(k * 0.75 / adl))
    for row, t in enumerate(picks):
        sel = terms == t
        tf = counts[sel]
        T = np.log(tf.astype(np.float64) + 1).astype(np.float32)
        wt = np.sort(T)[::-1]
        Tb = T.astype(np.float64)
        wb = np.sort(Tb / (Tb + K0 + K1 * dl[doc_of[sel]]))[::-1]
        assert len(tf) == df[t - 1]
        for j, step in enumerate(eng.KTH_STEPS):
            if step > len(tf):
                assert kth_t[row, j] == 0 and kth_b[row, j] == 0, (t, step)
                continue
            assert kth_t[row, j].view(np.uint32) == wt[step - 1].view(np.uint32), (t, step)
            assert abs(float(kth_b[row, j]) - wb[step - 1]) <= 2e-6 * wb[step - 1], (t, step, kth_b[row, j], wb[step - 1])


def test_priming_only_raises_the_start(corpus, eng, monkeypatch):
    """An engine without priming (NXSB_PRIME=0) gives the same answers and
    scores more blocks."""
    from nxsearch_b200 import engine

    qs = or_queries(corpus, 512, seed_off=3)
    batch = engine.Batch.from_lists(BM25, 10, qs)
    eng.pruning_stats(reset=True)
    primed = eng.search(batch)
    s1 = eng.pruning_stats()
    monkeypatch.setenv("NXSB_PRIME", "0")
    plain = engine.Engine(0)
    try:
        plain.load_corpus(corpus)
        unprimed = plain.search(batch)
        s0 = plain.pruning_stats()
    finally:
        plain.close()
    assert_identical(primed, unprimed)
    assert s1["blocks_scored"] < s0["blocks_scored"], (s1, s0)


@pytest.mark.parametrize("algo,limit", [(TFIDF, 100), (BM25, 10), (TFIDF, 1), (BM25, 128)])
def test_boolean_pruned_equals_exhaustive_and_oracle(corpus, oracle, eng, algo, limit):
    """AND / OR / NOT queries through the pruned scorer (block bounds +
    present-token masks + a membership byte per document) against the
    streaming scorer -- bit for bit -- and the oracle."""
    from nxsearch_b200 import engine
    from test_gpu_stream import bool_queries

    qs = bool_queries(corpus, 384)
    pruned, full, stats = both_ways(eng, engine.Batch.from_lists(algo, limit, qs))
    assert_identical(pruned, full)
    assert stats["items"] > 0, "boolean queries did not reach the block-max kernel"
    counts, ids, scores = pruned
    nonempty = 0
    for i, (toks, prog) in enumerate(qs[:160]):
        all_ids, all_sc = oracle.search_all(algo, toks, prog)
        nonempty += len(all_ids) > 0
        check_topk(ids[i, :counts[i]], scores[i, :counts[i]], all_ids, all_sc, limit,
                   exact_scores=(algo == TFIDF))
    assert nonempty > 80


def test_boolean_edge_shapes_pruned(corpus, oracle, eng):
    """Unresolved leaves (the empty set), NOT of everything, the same term on
    both sides, rare AND rare, and a mixed batch of OR and boolean queries."""
    from nxsearch_b200 import engine
    from _oracle import OP_AND, OP_ANDNOT, OP_EMPTY

    df = np.asarray(corpus.term_df)
    order = np.argsort(df, kind="stable")
    rare = [int(t) + 1 for t in order[df[order] >= 1][:8]]
    mid = [int(t) + 1 for t in np.nonzero((df >= 200) & (df <= 2000))[0][:8]]
    qs = [
        ([1, 2], [0, 1, OP_AND]), ([1, 2], [0, 1, OP_ANDNOT]), ([2, 1], [1, 0, OP_ANDNOT]),
        ([1], [0, OP_EMPTY, OP_AND]), ([1], [0, OP_EMPTY, OP_OR]), ([1], [OP_EMPTY, 0, OP_ANDNOT]),
        ([rare[0], rare[1]], [0, 1, OP_AND]), ([rare[0], 1], [0, 1, OP_AND]), ([mid[0], mid[1]], [0, 1, OP_AND]),
        ([mid[0], rare[2], 3], [0, 1, OP_OR, 2, OP_AND]), ([1, 2, 3, 4, 5, 6, 7, 8], [0, 1, OP_AND, 2, OP_AND, 3, OP_OR, 4, OP_ANDNOT, 5, OP_OR, 6, OP_AND, 7, OP_OR]),
        ([1, 2], None), ([mid[2]], None), ([3, corpus.n_terms + 5], [0, 1, OP_AND]),
    ]
    for algo, limit in ((TFIDF, 100), (BM25, 10)):
        pruned, full, _ = both_ways(eng, engine.Batch.from_lists(algo, limit, qs))
        assert_identical(pruned, full)
        counts, ids, scores = pruned
        for i, (toks, prog) in enumerate(qs):
            all_ids, all_sc = oracle.search_all(algo, toks, prog)
            check_topk(ids[i, :counts[i]], scores[i, :counts[i]], all_ids, all_sc, limit,
                       exact_scores=(algo == TFIDF))


def test_boolean_queries_of_any_shape_can_be_pruned(corpus, eng, monkeypatch):
    """The routing (queries with more than three positive tokens stream) is a
    cost estimate, not a limit: with it lifted every template goes through the
    pruned scorer and the answers stay bit-identical."""
    from nxsearch_b200 import engine
    from test_gpu_stream import bool_queries

    monkeypatch.setenv("NXSB_BMW_LOGIC_POS", "8")
    allp = engine.Engine(0)
    try:
        allp.load_corpus(corpus)
        qs = bool_queries(corpus, 192)
        for algo, limit in ((TFIDF, 100), (BM25, 10)):
            batch = engine.Batch.from_lists(algo, limit, qs)
            allp.pruning_stats(reset=True)
            got = allp.search(batch)
            assert allp.pruning_stats()["items"] >= len(qs), "not every query was pruned"
            eng.set_pruning(False)
            full = eng.search(batch)
            eng.set_pruning(True)
            assert_identical(got, full)
    finally:
        allp.close()
