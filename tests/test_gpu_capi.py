"""GPU parity through the drop-in boundary: the public nxs_* C API of this
library against (a) the reference's golden vectors, (b) the compiled reference
(oracle/_ref) on the same index files, (c) the CPU oracle."""
import json
import shutil
import tempfile
from pathlib import Path

import numpy as np
import pytest

import _mk
import _oracle
from _oracle import BM25, TFIDF, OP_OR, check_topk
from nxsearch_b200 import capi, tools

pytestmark = pytest.mark.gpu


@pytest.fixture()
def nxs():
    base = tempfile.mkdtemp(prefix="nxsb_g_")
    n = capi.Nxs(base)
    n.base = base
    yield n
    n.close()
    shutil.rmtree(base, ignore_errors=True)


def ref_nxs():
    lib = _oracle.ref()
    if lib is None:
        pytest.skip("oracle/_ref not built")
    base = tempfile.mkdtemp(prefix="nxsb_gr_")
    n = capi.Nxs(base, lib=lib)
    n.base = base
    return n


def same_results(got, ref, rtol=1e-5):
    """Equal up to (near-)ties: same score sequence; a document present on one
    side only must sit in the score group that the limit cut through."""
    assert len(got) == len(ref)
    if not got:
        return
    gs, rs = [s for _, s in got], [s for _, s in ref]
    assert all(abs(a - b) <= rtol * abs(b) for a, b in zip(gs, rs)), (gs, rs)
    assert all(a >= b - rtol * abs(a) for a, b in zip(gs, gs[1:]))
    cut = min(rs)
    gd, rd = dict(got), dict(ref)
    for d in set(gd) ^ set(rd):
        s = gd.get(d, rd.get(d))
        assert abs(s - cut) <= rtol * abs(cut), (d, s, cut)
    for d in set(gd) & set(rd):
        assert abs(gd[d] - rd[d]) <= rtol * abs(rd[d]), (d, gd[d], rd[d])


@pytest.mark.parametrize("case", range(len(_mk.SCORING_CASES)))
def test_reference_scoring_goldens(nxs, case):
    """ref tests/t_scoring.c through nxs_index_add / nxs_index_search, both algorithms."""
    docs, query, expected = _mk.SCORING_CASES[case]
    idx = nxs.create_index("t")
    for d, text in docs:
        idx.add(d, text)
    for algo, col in (("TF-IDF", 0), ("BM25", 1)):
        got = dict(idx.search(query, algo=algo))
        assert set(got) == set(expected)
        for d, vals in expected.items():
            assert abs(got[d] - vals[col]) < 1e-4, (algo, d, got[d], vals[col])
    idx.close()


@pytest.mark.parametrize("case", range(len(_mk.STEMMED_SCORING_CASES)))
def test_reference_scoring_goldens_that_need_the_stemmer(nxs, case, monkeypatch):
    """ref tests/t_scoring.c cases 2 and 3 with the reference's default filters:
    "fox" finds "foxes" (the restated Snowball english stemmer, stem_en.c)."""
    monkeypatch.delenv("NXSB_STEMMER_PASSTHROUGH", raising=False)
    docs, query, expected = _mk.STEMMED_SCORING_CASES[case]
    idx = nxs.create_index("t", filters=["normalizer", "stopwords", "stemmer"])
    for d, text in docs:
        idx.add(d, text)
    for algo, col in (("TF-IDF", 0), ("BM25", 1)):
        got = dict(idx.search(query, algo=algo))
        assert set(got) == set(expected)
        for d, vals in expected.items():
            assert abs(got[d] - vals[col]) < 1e-4, (algo, d, got[d], vals[col])
    idx.close()


def test_reference_querylogic_goldens(nxs):
    idx = nxs.create_index("t")
    for d, text in _mk.LOGIC_DOCS:
        idx.add(d, text)
    for query, expected in _mk.LOGIC_CASES:
        for algo in ("BM25", "TF-IDF"):
            assert sorted(d for d, _ in idx.search(query, algo=algo)) == expected, query
    # SURVEY 8a F5 [probed]: tokens under NOT still add their score
    plain = dict(idx.search("erlang AND NOT windows"))
    both = dict(idx.search("erlang AND NOT (windows AND linux)"))
    assert set(both) == set(plain) and both[1] > plain[1] and both[3] == plain[3]
    idx.close()


def test_response_json_and_iteration(nxs):
    idx = nxs.create_index("t")
    idx.add(1, "alpha beta")
    idx.add(2, "alpha alpha gamma")
    r = idx.search_resp("alpha", algo="TF-IDF")
    doc = json.loads(r.tojson())
    assert r.count == 2 and doc["count"] == 2 and [x["doc_id"] for x in doc["results"]] == [2, 1]
    assert r.tojson().startswith('{"results":[{"doc_id":2,"score":') and r.tojson().endswith('],"count":2}')
    res = r.results()
    assert res == r.results()                       # iter_reset restarts
    assert [np.float32(x["score"]) for x in doc["results"]] == [np.float32(s) for _, s in res]
    r.release()
    idx.close()


@pytest.fixture(scope="module")
def c1_both(c1_corpus):
    """The C1 index opened by this library and by the compiled reference."""
    base = tempfile.mkdtemp(prefix="nxsb_c1_")
    ours = capi.Nxs(base)
    ours.create_index("c1").close()
    c1_corpus.write(f"{base}/data/c1/nxsterms", f"{base}/data/c1/nxsdtmap")
    oidx = ours.open_index("c1")
    ridx = rn = None
    if _oracle.ref() is not None:
        rn = capi.Nxs(base, lib=_oracle.ref())
        ridx = rn.open_index("c1")
    yield oidx, ridx
    oidx.close()
    ours.close()
    if ridx:
        ridx.close()
        rn.close()
    shutil.rmtree(base, ignore_errors=True)


@pytest.mark.parametrize("algo,name", [(BM25, "BM25"), (TFIDF, "TF-IDF")])
def test_c1_batch_search_matches_reference(c1_corpus, c1_oracle, c1_both, algo, name):
    """BASELINE config 1 through nxs_index_search_batch; every query compared
    with the reference's nxs_index_search on the same files."""
    ours, ref = c1_both
    qt = c1_corpus.query_terms(2500)
    queries, meta, pos = [], [], 0
    for qi in range(1000):
        nt = 1 + qi // 250
        leaves = [int(t) for t in qt[pos:pos + nt]]
        pos += nt
        queries.append(" OR ".join(c1_corpus.term(t) for t in leaves))
        toks = []
        for t in reversed(leaves):
            if t not in toks:
                toks.append(t)
        meta.append(toks)
    got = ours.search_batch(queries, limit=10, algo=name, fuzzymatch=False)
    identical = 0
    for q, toks, res in zip(queries, meta, got):
        all_ids, all_sc = c1_oracle.search_all(algo, toks)
        check_topk(np.array([d for d, _ in res], dtype=np.uint64), np.array([s for _, s in res], dtype=np.float32),
                   all_ids, all_sc, 10, exact_scores=(algo == TFIDF))
        if ref is not None:
            r = ref.search(q, limit=10, algo=name, fuzzymatch=False)
            assert len(r) == len(res)
            identical += [d for d, _ in r] == [d for d, _ in res]
            # same multiset of scores (ties may pick different documents)
            assert np.allclose(sorted(s for _, s in r), sorted(s for _, s in res), rtol=1e-5, atol=0)
    if ref is not None:
        assert identical >= 100
    # single-query entry point agrees with the batch
    for q, res in list(zip(queries, got))[::97]:
        assert ours.search(q, limit=10, algo=name, fuzzymatch=False) == res


def test_fuzzy_queries_match_reference(c1_corpus, c1_oracle, c1_both):
    """fuzzymatch on: misspelled terms resolve to the term the reference's
    BK-tree search picks, so the result lists coincide (modulo ties)."""
    ours, ref = c1_both
    if ref is None:
        pytest.skip("oracle/_ref not built")
    # The reference reads freed memory when a query mixes resolvable and
    # unresolvable terms, so only probes that DO resolve are used.
    probes = [q.decode() for q in c1_corpus.fuzzy_terms(900, seed=21) if c1_oracle.fuzzy(q)[0]][:300]
    assert len(probes) == 300
    exact = [c1_corpus.term(int(t)) for t in c1_corpus.query_terms(300, seed=22)]
    hits = 0
    for i in range(0, 300, 2):
        q = f"{probes[i]} OR {probes[i + 1]}"
        r = ref.search(q, limit=20, algo="BM25")
        g = ours.search(q, limit=20, algo="BM25")
        same_results(g, r)
        hits += bool(r)
    assert hits > 30
    # batched path with a mix of exact and fuzzy queries
    qs = [f"{probes[i]}" if i % 2 else exact[i] for i in range(200)]
    got = ours.search_batch(qs, limit=10, algo="TF-IDF")
    for q, g in zip(qs, got):
        same_results(g, ref.search(q, limit=10, algo="TF-IDF"))


def test_unresolved_leaf_is_the_empty_set(c1_corpus, c1_oracle, c1_both):
    """Mixing resolvable and unresolvable terms crashes the reference (SURVEY 8a
    F6); here an unresolved leaf is the empty set, checked against the oracle."""
    ours, _ = c1_both
    t = [int(x) for x in c1_corpus.query_terms(4, seed=3)]
    a, b = c1_corpus.term(t[0]), c1_corpus.term(t[1])
    cases = {
        f"{a} OR zzzzzzzzzzzzzzzzzzzzzz": ([t[0]], [-1, 0, OP_OR][1:2] + [-1, OP_OR]),
        f"{a} AND zzzzzzzzzzzzzzzzzzzzzz": ([t[0]], [0, -1, -2]),
        f"({a} AND NOT zzzzzzzzzzzzzzzzzzzzzz) OR {b}": None,
    }
    for q, spec in cases.items():
        got = ours.search(q, limit=50, algo="BM25", fuzzymatch=False)
        if spec is None:
            toks = [t[1], t[0]] if t[0] != t[1] else [t[0]]
            prog = [toks.index(t[0]), -1, -4, toks.index(t[1]), OP_OR]
        else:
            toks, prog = spec
        all_ids, all_sc = c1_oracle.search_all(BM25, toks, prog)
        check_topk(np.array([d for d, _ in got], dtype=np.uint64), np.array([s for _, s in got], dtype=np.float32),
                   all_ids, all_sc, 50)


def test_incremental_add_remove_and_second_handle(nxs):
    """nxs_index_search re-syncs the files first (search.c:309-310): documents
    added or removed -- by this handle or another instance -- show up."""
    idx = nxs.create_index("live")
    for d in range(1, 41):
        idx.add(d, f"common word{d % 5} filler{d}")
    assert len(idx.search("common", limit=100)) == 40
    idx.add(100, "common common rare")
    top = idx.search("common OR rare", limit=3)
    assert top[0][0] == 100
    idx.remove(100)
    assert 100 not in [d for d, _ in idx.search("common OR rare", limit=100)]
    assert idx.search("rare") == []
    other = capi.Nxs(nxs.base)                       # a second "process"
    oidx = other.open_index("live")
    oidx.add(200, "common brandnewterm")
    assert [d for d, _ in idx.search("brandnewterm")] == [200]
    oidx.remove(3)
    assert 3 not in [d for d, _ in idx.search("common", limit=100)]
    oidx.close()
    other.close()
    # and the reference agrees on the final state of the very same files
    ref = _oracle.ref()
    if ref is not None:
        rn = capi.Nxs(nxs.base, lib=ref)
        ridx = rn.open_index("live")
        for q in ("common", "word1 OR word2", "common AND NOT word3"):
            for algo in ("BM25", "TF-IDF"):
                same_results(idx.search(q, limit=100, algo=algo), ridx.search(q, limit=100, algo=algo))
        ridx.close()
        rn.close()
    idx.close()


def test_ids_limits_and_wide_documents(nxs):
    """Arbitrary 64-bit ids in any order, the default limit of 1000, limits
    beyond the matches, and a document too long for packed postings."""
    idx = nxs.create_index("w")
    ids = [2**63 + 5, 7, 2**40, 3, 2**64 - 1, 11]
    for i, d in enumerate(ids):
        idx.add(d, "shared " + "extra " * i)
    res = idx.search("shared")
    assert sorted(d for d, _ in res) == sorted(ids)
    assert [d for d, _ in idx.search("shared", algo="TF-IDF")] == sorted(ids, reverse=True)   # all tied: id desc
    idx.add(99, "giant " * 70_000 + "shared")            # tf and doc length >= 65536
    res = dict(idx.search("giant OR shared", algo="BM25", limit=100))
    ref = _oracle.ref()
    if ref is not None:
        rn = capi.Nxs(nxs.base, lib=ref)
        ridx = rn.open_index("w")
        same_results(sorted(res.items(), key=lambda x: -x[1]), ridx.search("giant OR shared", algo="BM25", limit=100))
        ridx.close()
        rn.close()
    for d in range(1000, 2200):
        idx.add(d, "bulk")
    assert len(idx.search("bulk")) == 1000                # NXS_DEFAULT_RESULTS_LIMIT
    assert len(idx.search("bulk", limit=5000)) == 1200
    assert len(idx.search("bulk", limit=2**32 - 1)) == 1200
    idx.close()


def test_batch_begin_end_pipeline_equals_synchronous_calls(c1_corpus, c1_both):
    """nxs_index_search_batch_begin/_end: several batches in flight give the
    arrays the synchronous call gives, in any end order; a fifth begin fails
    with NXS_ERR_SYSTEM; resps = NULL abandons a batch."""
    import ctypes as C

    ours, _ = c1_both
    qt = c1_corpus.query_terms(4 * 3 * 200, seed=77)
    batches = []
    for b in range(4):
        qs = []
        for i in range(200):
            leaves = qt[(b * 200 + i) * 3:(b * 200 + i) * 3 + 1 + i % 3]
            qs.append(" OR ".join(c1_corpus.term(int(t)) for t in leaves))
        if b == 2:
            qs[5] = "(("           # a syntax error fails this query only
        batches.append(qs)
    params = dict(algo="BM25", fuzzymatch=False)
    want = []
    for qs in batches:
        ok = [q for q in qs if q != "(("]
        want.append(ours.search_batch_arrays(ok, 10, **params))
    tickets = [ours.search_batch_begin(qs, 10, **params) for qs in batches]
    with pytest.raises(capi.NxsError):
        ours.search_batch_begin(batches[0], 10, **params)
    for b in (1, 0, 3, 2):
        counts, ids, scores = ours.search_batch_end_arrays(tickets[b])
        if b == 2:
            keep = [i for i in range(200) if i != 5]
            assert counts[5] == 0
            counts, ids, scores = counts[keep], ids[keep], scores[keep]
        assert np.array_equal(counts, want[b][0])
        assert np.array_equal(ids, want[b][1])
        assert np.array_equal(scores, want[b][2])
    # abandoned batch: _end with resps = NULL frees the slot
    for _ in range(6):
        bt, n, _, _ = ours.search_batch_begin(batches[0], 10, **params)
        assert ours._lib.nxs_index_search_batch_end(bt, None) == 0
    c, i, s = ours.search_batch_arrays(batches[0], 10, **params)
    assert np.array_equal(i, want[0][1])


def test_committed_reference_goldens_through_the_c_api(c1_corpus, c1_both):
    """tests/golden/c1_reference.npz -- outputs of the compiled reference, made by
    tests/golden/make_golden.py -- against nxs_index_search_batch /
    nxs_index_search on the same C1 files.  Needs neither oracle/_ref nor
    /root/reference.  TF-IDF scores are bit-equal; BM25 within 1e-5 relative;
    documents may differ only inside the score group the limit cuts through."""
    from _golden import Golden

    ours, _ = c1_both
    g = Golden()
    g.check_corpus(c1_corpus)
    for family, queries, limit in (("or", g.or_queries, 10), ("bool", g.bool_queries, 100)):
        for name, key in (("BM25", "bm25"), ("TF-IDF", "tfidf")):
            got = ours.search_batch(queries, limit=limit, algo=name, fuzzymatch=False)
            same_ids = 0
            for i, (q, res) in enumerate(zip(queries, got)):
                ref = g.results(family, key, i)
                same_results(res, ref)
                if key == "tfidf":
                    assert [np.float32(s) for _, s in res] == [np.float32(s) for _, s in ref], q
                same_ids += sorted(d for d, _ in res) == sorted(d for d, _ in ref)
            assert same_ids >= len(queries) // 10, (family, key, same_ids)   # the rest differ in boundary ties only
    # fuzzy: a misspelt term alone must return exactly what the picked term returns
    probes = [(q, p) for q, p in zip(g.fuzzy_queries, g.fuzzy_pick.tolist()) if p][:120]
    got = ours.search_batch([q for q, _ in probes], limit=10, algo="TF-IDF", fuzzymatch=True)
    want = ours.search_batch([c1_corpus.term(p) for _, p in probes], limit=10, algo="TF-IDF", fuzzymatch=False)
    assert got == want
    # ... and one the reference found nothing for stays empty
    misses = [q for q, p in zip(g.fuzzy_queries, g.fuzzy_pick.tolist()) if not p][:20]
    assert all(r == [] for r in ours.search_batch(misses, limit=10, algo="TF-IDF", fuzzymatch=True))


def test_fuzzy_after_remove_and_readd_agrees_with_the_reference(nxs):
    """Fuzzy picks depend on the per-term totals in nxsterms (ref idxterm.c:239
    reads the counter at search time), which move with every add / remove, also
    when no new term appears.  The reference's counters start at
    twice the first count (terms.c + dtmap.c:236 both add it), so a remove
    leaves them above zero and the misspelling keeps resolving to the emptied
    term -- an empty result, on both sides.  Checked step by step on the same
    files; the flags themselves are exercised in
    test_gpu_engine.py::test_fuzzy_live_flags_follow_the_totals."""
    idx = nxs.create_index("fz")
    idx.add(1, "alpha betas gamma")
    idx.add(2, "alpha betaz delta")
    idx.add(3, "alpha omega")
    ref = _oracle.ref()

    def agree(q):
        got = idx.search(q, limit=10, algo="BM25")
        if ref is not None:
            rn = capi.Nxs(nxs.base, lib=ref)
            ridx = rn.open_index("fz")
            same_results(got, ridx.search(q, limit=10, algo="BM25"))
            ridx.close()
            rn.close()
        return [d for d, _ in got]

    assert agree("betax") == [1]            # BFS-first candidate: betas
    idx.remove(1)                           # no term is added: only totals and df move
    assert agree("betax") == []             # still betas (total 2 - 1 > 0), which has no documents left
    assert agree("betaz") == [2]
    idx.add(7, "betas again")               # an EXISTING term gets a document back
    assert agree("betax") == [7]
    idx.close()


def test_one_oversized_query_does_not_take_the_batch_down(nxs):
    """a right-nested query whose postfix program needs a deeper
    operand stack than the engine has (one repeated term: 1 token, ~80 nodes)
    passes the token and node limits; it must fail alone with NXS_ERR_LIMIT."""
    idx = nxs.create_index("deep")
    for d in range(1, 30):
        idx.add(d, f"aa bb{d % 3} cc")
    deep = "aa" + "".join(" AND (aa" for _ in range(40)) + ")" * 40
    res = idx.search_batch(["aa", deep, "bb1 OR cc"], limit=50)
    assert res[1] is None and nxs.error()[0] == capi.ERR_LIMIT
    assert len(res[0]) == 29 and len(res[2]) == 29
    with pytest.raises(capi.NxsError) as e:
        idx.search(deep)
    assert e.value.code == capi.ERR_LIMIT
    # nesting the engine does hold still works
    ok = "aa" + "".join(" AND (aa" for _ in range(20)) + ")" * 20
    assert len(idx.search(ok, limit=50)) == 29
    idx.close()


def test_a_c_program_links_the_library_and_gets_the_same_answers(nxs):
    """The boundary from C: tests/c/nxs_caller.c, compiled against
    include/nxs.h and linked with libnxsearch.so (as the reference's own CLI
    is, ref src/utils/benchmark.c:199-215), opens the index files another
    process wrote and returns what the ctypes binding returns."""
    import _ccaller

    corpus = tools.Corpus.generate(20_000, 5_000)
    nxs.create_index("c").close()
    corpus.write(f"{nxs.base}/data/c/nxsterms", f"{nxs.base}/data/c/nxsdtmap")
    idx = nxs.open_index("c")
    qt = corpus.query_terms(400)
    queries = [" OR ".join(corpus.term(int(t)) for t in qt[i:i + 1 + i % 4]) for i in range(0, 300, 4)]
    queries += [f"{corpus.term(int(qt[300 + i]))} AND NOT {corpus.term(int(qt[350 + i]))}" for i in range(20)]
    for algo, limit in (("BM25", 10), ("TF-IDF", 100)):
        got = _ccaller.results(nxs.base, "c", algo, limit, queries)
        assert len(got) == len(queries)
        for q, g in zip(queries, got):
            ref = idx.search(q, limit=limit, algo=algo, fuzzymatch=False)
            assert [d for d, _ in g] == [d for d, _ in ref], q
            assert all(np.float32(a) == np.float32(b) for (_, a), (_, b) in zip(g, ref)), q
    lat = _ccaller.latency(nxs.base, "c", "BM25", 10, queries)
    assert lat["queries"] == len(queries) and lat["p50_us"] > 0
    idx.close()


@pytest.mark.parametrize("layout", ["replicas", "shards"])
def test_replicated_engine_behind_the_c_api(nxs, monkeypatch, layout):
    """NXS_GPU_DEVICES: one nxs_t over every listed device -- a replica of the
    image on each and a batch's queries split between them, or
    (NXS_GPU_LAYOUT=shards) a range of the documents on each, every query
    scored everywhere and the lists merged on the first device.  Three of them
    (all on device 0 here, every visible device on a multi-GPU box) must answer
    exactly as one engine does: batches, single searches, fuzzy terms, and
    after adds and removes with searches in between (delta segments, removal
    notes)."""
    from nxsearch_b200 import engine

    corpus = tools.Corpus.generate(30_000, 8_000)
    nxs.create_index("r").close()
    corpus.write(f"{nxs.base}/data/r/nxsterms", f"{nxs.base}/data/r/nxsdtmap")
    qt = corpus.query_terms(600)
    queries = [" OR ".join(corpus.term(int(t)) for t in qt[i:i + 1 + i % 4]) for i in range(0, 400, 4)]
    queries += [f"{corpus.term(int(qt[400 + i]))} AND NOT {corpus.term(int(qt[450 + i]))}" for i in range(30)]
    queries += [q.decode() for q in corpus.fuzzy_terms(12)]            # resolved by the vocabulary scan
    queries += ["(", "zzzzzzzzzzzzzz"]                                # one syntax error, one no-match
    ndev = engine.device_count()
    devs = ",".join(str(d % ndev) for d in range(max(3, ndev)))

    def run(base_nxs):
        idx = base_nxs.open_index("r")
        out = {}
        for algo, limit in (("BM25", 10), ("TF-IDF", 100)):
            out[algo, "batch"] = idx.search_batch(queries, limit=limit, algo=algo)
            out[algo, "two"] = idx.search_batch(queries[:2], limit=limit, algo=algo)   # fewer queries than replicas
            out[algo, "single"] = [idx.search(q, limit=limit, algo=algo) for q in queries[:5]]
        top = out["BM25", "batch"][0][0][0]
        idx.add(10_000_000 + len(out), f"{corpus.term(int(qt[0]))} {corpus.term(int(qt[1]))} brandnewterm")
        idx.remove(top)
        out["after"] = idx.search_batch(queries[:64] + ["brandnewterm"], limit=10, algo="BM25")
        # more edits, a search after each: several delta segments and removal notes
        for j in range(1, 6):
            idx.add(20_000_000 + j, f"{corpus.term(int(qt[j]))} {corpus.term(int(qt[j + 1]))} brandnewterm")
            res = idx.search_batch(queries[4 * j: 4 * j + 24] + ["brandnewterm"], limit=10, algo="BM25")
            out["edit", j] = res
            if res[0]:
                idx.remove(res[0][0][0])
            if j == 3:
                idx.remove(20_000_001)
        out["after edits"] = idx.search_batch(queries + ["brandnewterm"], limit=100, algo="TF-IDF")
        idx.close()
        return out

    one = run(nxs)
    # undo the edits: the second run must start from the same files
    shutil.rmtree(f"{nxs.base}/data/r")
    nxs.create_index("r").close()
    corpus.write(f"{nxs.base}/data/r/nxsterms", f"{nxs.base}/data/r/nxsdtmap")
    monkeypatch.setenv("NXS_GPU_DEVICES", devs)
    monkeypatch.setenv("NXS_GPU_LAYOUT", layout)
    multi_nxs = capi.Nxs(nxs.base)
    try:
        many = run(multi_nxs)
    finally:
        multi_nxs.close()
    assert one.keys() == many.keys()
    for k in one:
        assert one[k] == many[k], k
    assert one["after"][-1] and one["BM25", "batch"][-2] is None and one["BM25", "batch"][-1] == []


def test_long_fuzzy_patterns_are_exact_or_loud(nxs):
    """The vocabulary scan takes patterns of up to 64 bytes.  A longer query
    term can only match a vocabulary term within 2 bytes of its length: with
    none that long "no match" is the exact answer; with one, the query fails
    alone with NXS_ERR_LIMIT instead of missing silently (the reference would
    find it, ref src/index/idxterm.c:210-249)."""
    idx = nxs.create_index("f", filters=["normalizer"])
    idx.add(1, "alpha beta gamma")
    idx.add(2, "beta delta")
    long_q = "x" * 70
    res = idx.search_batch([long_q, "beta", f"{long_q} OR delta"], limit=10, algo="BM25")
    assert res[0] == [] and [d for d, _ in res[1]] == [2, 1] and [d for d, _ in res[2]] == [2]
    idx.add(3, "y" * 69 + " beta")                       # now the vocabulary has a term that long
    res = idx.search_batch([long_q, "beta", "y" * 69], limit=10, algo="BM25")
    assert res[0] is None and len(res[1]) == 3 and [d for d, _ in res[2]] == [3]
    code, msg = nxs.error()
    assert code == capi.ERR_LIMIT and "fuzzy match" in msg
    with pytest.raises(capi.NxsError) as e:
        idx.search(long_q, limit=10)
    assert e.value.code == capi.ERR_LIMIT
    assert [d for d, _ in idx.search(long_q, limit=10, fuzzymatch=False)] == []
    idx.close()
