"""BASELINE.json's full size -- 10M synthetic documents, 1M-term vocabulary
(configs C2 / C3) -- checked through properties that do not need a ground
truth for every query: shard invariance (G shards + merge == one image, bit
for bit), agreement of the two scoring kernels, order and count invariants,
incremental refresh == rebuild; plus a sample of queries against the oracle."""
import os

import numpy as np
import pytest

import _oracle
from _oracle import BM25, TFIDF, check_topk

pytestmark = pytest.mark.gpu

N_DOCS = int(os.environ.get("NXSB_FULLSIZE_DOCS", 10_000_000))
N_TERMS = 1_000_000


@pytest.fixture(scope="module")
def corpus():
    from nxsearch_b200 import tools

    return tools.Corpus.generate(N_DOCS, N_TERMS)


@pytest.fixture(scope="module")
def whole(corpus):
    from nxsearch_b200 import engine

    e = engine.Engine(0)
    e.load_corpus(corpus)
    yield e
    e.close()


def queries(corpus):
    from test_gpu_stream import or_queries, bool_queries
    return or_queries(corpus, 256), bool_queries(corpus, 96)


def same(a, b):
    (ca, ia, sa), (cb, ib, sb) = a, b
    assert np.array_equal(ca, cb)
    for q in range(len(ca)):
        n = int(ca[q])
        assert np.array_equal(ia[q, :n], ib[q, :n]), q
        assert np.array_equal(sa[q, :n], sb[q, :n]), q


def test_order_and_count_invariants(corpus, whole):
    from nxsearch_b200 import engine

    ors, bools = queries(corpus)
    for algo, k, qs in ((BM25, 10, ors), (TFIDF, 100, bools), (BM25, 100, ors[:64])):
        counts, ids, scores = whole.search(engine.Batch.from_lists(algo, k, qs))
        for i, (toks, prog) in enumerate(qs):
            n = int(counts[i])
            s, d = scores[i, :n], ids[i, :n].astype(np.int64)
            assert np.all(s > 0) and len(set(d.tolist())) == n
            # descending score, ties by descending id
            assert np.all((s[:-1] > s[1:]) | ((s[:-1] == s[1:]) & (d[:-1] > d[1:])))
            if len(toks) == 1:
                assert n == min(k, int(corpus.term_df[toks[0] - 1]))
            elif qs is ors:
                # an OR matches at least the documents of its most frequent term
                assert n >= min(k, max(int(corpus.term_df[t - 1]) for t in toks))


def test_shards_plus_merge_equal_the_whole_index(corpus, whole):
    import torch
    from nxsearch_b200 import dist as nxdist, engine

    ors, bools = queries(corpus)
    G = 4
    shards = []
    for g in range(G):
        lo, hi = nxdist.shard_range(corpus.n_docs, g, G)
        e = engine.Engine(0)
        e.load_corpus(corpus, lo=lo, hi=hi, df=corpus.term_df,
                      token_count=corpus.token_count, doc_count=corpus.doc_count)
        shards.append(e)
    for algo, k, qs in ((BM25, 10, ors), (TFIDF, 100, bools)):
        batch = engine.Batch.from_lists(algo, k, qs)
        want = whole.search(batch)
        per = len(qs) * k * nxdist.REC_BYTES
        gathered = torch.zeros(G * per, dtype=torch.uint8, device="cuda")
        for g, e in enumerate(shards):
            h = e.upload(batch)
            e.run(h, gathered.data_ptr() + g * per)
            e.sync()
            e.release(h)
        merged = torch.zeros(per, dtype=torch.uint8, device="cuda")
        shards[0].merge_topk(gathered.data_ptr(), G, len(qs), k, merged.data_ptr())
        shards[0].sync()
        recs = merged.cpu().numpy().view(nxdist.REC_DTYPE).reshape(len(qs), k)
        for i in range(len(qs)):
            n = int(recs[i]["valid"].sum())
            assert n == want[0][i]
            assert np.array_equal(recs[i]["doc_id"][:n], want[1][i, :n])
            assert np.array_equal(recs[i]["score"][:n], want[2][i, :n])
    for e in shards:
        e.close()


def test_one_sharded_engine_equals_the_whole_index(corpus, whole):
    """nxsb_engine_create_sharded at full size: three posting-balanced document
    ranges (every visible device, wrapping on a one-GPU box), peer copies and
    the on-device merge -- bit for bit the single-engine answer, with four
    searches in flight."""
    from nxsearch_b200 import engine

    ors, bools = queries(corpus)
    ndev = engine.device_count()
    sh = engine.Engine(devices=[d % ndev for d in range(max(3, ndev))], layout="shards")
    sh.load_corpus(corpus, df=corpus.term_df)
    for algo, k, qs in ((BM25, 10, ors), (TFIDF, 100, bools), (BM25, 200, ors[:48])):
        batch = engine.Batch.from_lists(algo, k, qs)
        want = whole.search(batch)
        same(sh.search(batch), want)
        hs = [sh.search_begin(batch) for _ in range(4)]
        for h in hs:
            same(sh.search_end(h, len(qs), k), want)
    sh.close()


def test_both_kernels_agree(corpus, whole):
    from nxsearch_b200 import engine

    ors, bools = queries(corpus)
    os.environ["NXSB_KERNEL"] = "v2"
    try:
        old = engine.Engine(0)
    finally:
        del os.environ["NXSB_KERNEL"]
    old.load_corpus(corpus)
    for algo, k, qs in ((BM25, 10, ors[:128]), (TFIDF, 100, bools[:48])):
        batch = engine.Batch.from_lists(algo, k, qs)
        same(whole.search(batch), old.search(batch))
    old.close()


def test_delta_segment_and_removals_equal_the_whole_index(corpus, whole):
    """The last 5000 documents as a delta segment, 30 documents removed."""
    from nxsearch_b200 import engine

    ors, bools = queries(corpus)
    cut = corpus.n_docs - 5000
    rng = np.random.default_rng(11)
    dead_base = np.sort(rng.choice(cut, 24, replace=False))
    dead_delta = np.sort(cut + rng.choice(5000, 6, replace=False))
    dead = np.concatenate([dead_base, dead_delta])
    # statistics without the removed documents
    df = np.array(corpus.term_df, dtype=np.int64)
    tokens = int(corpus.token_count)
    for i in dead:
        lo, hi = int(corpus.doc_off[i]), int(corpus.doc_off[i + 1])
        df[corpus.pairs[2 * lo:2 * hi:2].astype(np.int64) - 1] -= 1
        tokens -= int(corpus.doc_len[i])
    ndocs = corpus.n_docs - len(dead)

    seg = engine.Engine(0)
    seg.load_corpus(corpus, lo=0, hi=cut, df=corpus.term_df)
    seg.load_corpus(corpus, lo=cut, hi=corpus.n_docs, df=corpus.term_df, segment=True)
    seg.set_dead(0, corpus.doc_ids[dead_base])
    seg.set_dead(1, corpus.doc_ids[dead_delta])
    seg.set_global_stats(df.astype(np.uint32), tokens, ndocs)
    # the whole image with the same statistics, over-fetched and filtered on the host
    whole.set_global_stats(df.astype(np.uint32), tokens, ndocs)
    try:
        dead_ids = set(int(x) for x in corpus.doc_ids[dead])
        for algo, k, qs in ((BM25, 10, ors), (TFIDF, 100, bools)):
            got = seg.search(engine.Batch.from_lists(algo, k, qs))
            wc, wi, ws = whole.search(engine.Batch.from_lists(algo, k + len(dead), qs))
            for q in range(len(qs)):
                keep = [j for j in range(int(wc[q])) if int(wi[q, j]) not in dead_ids][:k]
                assert int(got[0][q]) == len(keep), q
                assert np.array_equal(got[1][q, :len(keep)], wi[q, keep]), q
                assert np.array_equal(got[2][q, :len(keep)], ws[q, keep]), q
    finally:
        whole.set_global_stats(np.asarray(corpus.term_df), corpus.token_count, corpus.doc_count)
        seg.close()


def test_sample_against_the_oracle(corpus, whole):
    """128 OR + 64 boolean queries of the full-size index against the oracle
    port (every match scored on the CPU, a host thread per query)."""
    from concurrent.futures import ThreadPoolExecutor
    from nxsearch_b200 import engine

    ors, bools = queries(corpus)
    ora = _oracle.OracleIndex(corpus)
    workers = max(1, min(16, len(os.sched_getaffinity(0))))
    try:
        for algo, k, qs in ((BM25, 10, ors[:128]), (TFIDF, 100, bools[:64]), (TFIDF, 10, ors[128:160])):
            counts, ids, scores = whole.search(engine.Batch.from_lists(algo, k, qs))
            with ThreadPoolExecutor(max_workers=workers) as pool:
                truth = list(pool.map(lambda q: ora.search_all(algo, q[0], q[1]), qs))
            for i, (all_ids, all_sc) in enumerate(truth):
                check_topk(ids[i, :counts[i]], scores[i, :counts[i]], all_ids, all_sc, k,
                           exact_scores=(algo == TFIDF))
    finally:
        ora.close()


def test_c4_fuzzy_at_a_million_terms():
    """BASELINE config 4 at size: 100 000 query terms against a 1M-term
    vocabulary in one call; 2 000 of them against the reference's BK-tree
    search (oracle port, host threads) -- chosen term and distance -- and, for
    a sample, the candidate sets: the reached list in the reference's order
    and the brute-force set of terms within distance 2."""
    from concurrent.futures import ThreadPoolExecutor
    from nxsearch_b200 import engine, tools

    corpus = tools.Corpus.generate(2000, N_TERMS)           # only the vocabulary matters
    e = engine.Engine(0)
    ora = _oracle.OracleIndex(corpus)
    workers = max(1, min(16, len(os.sched_getaffinity(0))))
    try:
        e.load_corpus(corpus)
        parent, edge, rank = corpus.bk_mirror()
        e.load_vocab(corpus.term_blob, corpus.term_off, corpus.term_total, parent, edge, rank)
        qs = corpus.fuzzy_terms(100_000)
        term, dist, true = e.fuzzy(qs, want_true=True)
        assert int((term != 0).sum()) > 20_000
        step = len(qs) // 2000
        sample = list(range(0, len(qs), step))[:2000]
        ora.fuzzy(qs[0])                                    # builds the BK-tree once, before the threads
        with ThreadPoolExecutor(max_workers=workers) as pool:
            ref = list(pool.map(lambda i: ora.fuzzy(qs[i], cap=1 << 16), sample))
        lev = _oracle.port().ora_levdist
        for i, (t, cands, dists, _) in zip(sample, ref):
            assert term[i] == t, (qs[i], term[i], t)
            if t:
                s = corpus.term(int(t)).encode()
                assert dist[i] == lev(qs[i], len(qs[i]), s, len(s))
            assert true[i] >= len(cands)
        # candidate sets
        sub = sample[:96]
        _, _, cnt, ct, cd, cf = e.fuzzy_candidates([qs[i] for i in sub], cap=1 << 14)
        with ThreadPoolExecutor(max_workers=workers) as pool:
            brute = list(pool.map(lambda i: ora.fuzzy_true(qs[i], cap=1 << 16), sub))
        for j, i in enumerate(sub):
            n = int(cnt[j])
            assert n == true[i] and n <= 1 << 14
            _, cands, dists, _ = ref[sample.index(i)]
            reached = [(int(t), int(d)) for t, d, f in zip(ct[j, :n], cd[j, :n], cf[j, :n]) if f & 1]
            assert reached == list(zip(cands.tolist(), dists.tolist())), qs[i]
            assert sorted(zip(ct[j, :n].tolist(), cd[j, :n].tolist())) == sorted(zip(brute[j][0].tolist(), brute[j][1].tolist())), qs[i]
    finally:
        ora.close()
        e.close()
        corpus.close()
