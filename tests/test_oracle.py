"""The oracle is pinned before it is trusted (CPU only).

1. the CPU restatement (oracle/nxs_oracle.c) against every golden vector the
   reference's own tests hold for this path;
2. the restatement against the reference itself -- its own C files compiled
   with the shims under oracle/shims (oracle/_ref) -- on the synthetic C1
   index: identical ids, scores and order, ties included.
"""
import ctypes as C
import shutil
import tempfile

import numpy as np
import pytest

import _mk
import _oracle
from _oracle import BM25, TFIDF, OP_AND, OP_OR, OP_ANDNOT


def test_levdist_goldens():
    lib = _oracle.port()
    for a, b, d in _mk.LEVDIST_CASES:
        assert lib.ora_levdist(a.encode(), len(a), b.encode(), len(b)) == d, (a, b)
        assert lib.ora_levdist(b.encode(), len(b), a.encode(), len(a)) == d, (b, a)


def test_bktree_goldens():
    """ref tests/t_bktree.c: the LAST candidate pushed for each misspelling is the word."""
    corpus = _mk.make_corpus([(1, " ".join(_mk.BK_WORDS))])
    ora = _oracle.OracleIndex(corpus)
    for word, probe in zip(_mk.BK_WORDS, _mk.BK_SEARCH):
        _, cands, dists, _ = ora.fuzzy(probe.encode())
        assert len(cands) and corpus.term(int(cands[-1])) == word, (probe, cands)
        assert all(d <= 2 for d in dists)


def test_heap_goldens():
    """ref tests/t_heap.c:115-129: uncapped -> plain descending sort; capped -> top-N."""
    lib = _oracle.port()
    rng = np.random.default_rng(1)
    for n in range(1, 101):
        vals = rng.integers(0, 1000, n).astype(np.float32)
        pay = np.arange(n, dtype=np.uint64)
        out_p = np.zeros(n, dtype=np.uint64)
        out_s = np.zeros(n, dtype=np.float32)
        k = lib.ora_heap_topn(100, n, vals.ctypes.data, pay.ctypes.data, out_p.ctypes.data, out_s.ctypes.data)
        assert k == n and np.array_equal(out_s, np.sort(vals)[::-1])
        cap = max(1, n // 3)
        k = lib.ora_heap_topn(cap, n, vals.ctypes.data, pay.ctypes.data, out_p.ctypes.data, out_s.ctypes.data)
        assert k == cap and np.array_equal(out_s[:cap], np.sort(vals)[::-1][:cap])
    # SURVEY 8a F5 [probed]: limit 2 over four tied docs 5..8, fed in descending
    # id order as nxs_resp_build does, keeps 8 and 7.
    vals = np.ones(4, dtype=np.float32)
    pay = np.array([8, 7, 6, 5], dtype=np.uint64)
    out_p, out_s = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.float32)
    assert lib.ora_heap_topn(2, 4, vals.ctypes.data, pay.ctypes.data, out_p.ctypes.data, out_s.ctypes.data) == 2
    assert sorted(out_p[:2].tolist()) == [7, 8]


@pytest.mark.parametrize("case", range(len(_mk.SCORING_CASES)))
def test_scoring_goldens(case):
    """ref tests/t_scoring.c cases 1,4,5,6,7 at its own tolerance (helpers.c:215)."""
    docs, query, expected = _mk.SCORING_CASES[case]
    corpus = _mk.make_corpus(docs)
    ora = _oracle.OracleIndex(corpus)
    leaves = [corpus.tid(w) for w in query.split()]
    toks = []
    for t in reversed(leaves):
        if t not in toks:
            toks.append(t)
    slot = {t: s for s, t in enumerate(toks)}
    prog = [slot[leaves[0]]]
    for t in leaves[1:]:
        prog += [slot[t], OP_OR]
    for algo, col in ((TFIDF, 0), (BM25, 1)):
        ids, sc = ora.search(algo, 1000, toks, prog)
        got = dict(zip(ids.tolist(), sc.tolist()))
        assert set(got) == set(expected)
        for d, vals in expected.items():
            assert abs(got[d] - vals[col]) < 1e-4, (d, got[d], vals[col])


@pytest.mark.parametrize("case", range(len(_mk.STEMMED_SCORING_CASES)))
def test_scoring_goldens_that_need_the_stemmer(case):
    """ref tests/t_scoring.c cases 2 and 3: the documents go through this tree's
    front end (UAX #29 words, normalizer, the restated Snowball english stemmer,
    nxsearch_b200/csrc/host/stem_en.c) and the port scores them -- "fox" finds
    "foxes" with the reference's own scores."""
    from nxsearch_b200 import tools

    docs, query, expected = _mk.STEMMED_SCORING_CASES[case]
    stem = lambda text: [w for w, n in tools.tokenize(text, stem=True) for _ in range(n)]
    corpus = _mk.make_corpus([(d, stem(t)) for d, t in docs])
    assert corpus.tid("fox") and not corpus.tid("foxes") and corpus.tid("jump") and corpus.tid("lazi")
    ora = _oracle.OracleIndex(corpus)
    leaves = [corpus.tid(w) for w in stem(query)]
    toks = []
    for t in reversed(leaves):
        if t not in toks:
            toks.append(t)
    slot = {t: s for s, t in enumerate(toks)}
    prog = [slot[leaves[0]]]
    for t in leaves[1:]:
        prog += [slot[t], OP_OR]
    for algo, col in ((TFIDF, 0), (BM25, 1)):
        ids, sc = ora.search(algo, 1000, toks, prog)
        got = dict(zip(ids.tolist(), sc.tolist()))
        assert set(got) == set(expected)
        for d, vals in expected.items():
            assert abs(got[d] - vals[col]) < 1e-4, (d, got[d], vals[col])


def test_querylogic_goldens():
    """ref tests/t_querylogic.c:16-52."""
    from nxsearch_b200 import tools

    corpus = _mk.make_corpus(_mk.LOGIC_DOCS)
    ora = _oracle.OracleIndex(corpus)
    for query, expected in _mk.LOGIC_CASES:
        leaves, prog = tools.query_compile(query)
        ids = [corpus.tid(w.lower()) for w in leaves]
        keep = [i for i, t in enumerate(ids) if t]
        remap = {old: new for new, old in enumerate(keep)}
        prog = [(remap.get(op, -1) if op >= 0 else op) for op in prog]
        toks = [ids[i] for i in keep]
        got, _ = ora.search(BM25, 1000, toks, prog) if toks else (np.array([]), None)
        assert sorted(got.tolist()) == expected, query


# ---------------------------------------------------------------------------
# against the compiled reference


needs_ref = pytest.mark.skipif(not _oracle.REF_SO.exists(), reason="oracle/_ref not built (no /root/reference)")


@pytest.fixture(scope="module")
def ref_c1(c1_corpus):
    from nxsearch_b200 import capi

    base = tempfile.mkdtemp(prefix="nxsb_ref_")
    nxs = capi.Nxs(base, lib=_oracle.ref())
    nxs.create_index("c1").close()
    c1_corpus.write(f"{base}/data/c1/nxsterms", f"{base}/data/c1/nxsdtmap")
    idx = nxs.open_index("c1")
    yield idx
    idx.close()
    nxs.close()
    shutil.rmtree(base, ignore_errors=True)


@needs_ref
@pytest.mark.parametrize("algo,name", [(BM25, "BM25"), (TFIDF, "TF-IDF")])
def test_port_equals_reference_on_c1(c1_corpus, c1_oracle, ref_c1, algo, name):
    """BASELINE config 1: 1000 OR queries, 250 each of 1..4 terms, top-10."""
    qt = c1_corpus.query_terms(2500)
    pos = 0
    for qi in range(1000):
        nt = 1 + qi // 250
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
        ref = ref_c1.search(" OR ".join(c1_corpus.term(t) for t in leaves), limit=10, algo=name, fuzzymatch=False)
        ids, sc = c1_oracle.search(algo, 10, toks, prog)
        assert ref == list(zip(ids.tolist(), sc.tolist())), leaves


@needs_ref
def test_port_equals_reference_on_boolean_queries(c1_corpus, c1_oracle, ref_c1):
    """SURVEY 8d C3 templates through the reference's parser shim and search path."""
    from nxsearch_b200 import tools

    qt = [int(t) for t in c1_corpus.query_terms(6 * 120, seed=11)]
    w = c1_corpus.term
    for i in range(120):
        a, b, c, d, e, f = qt[6 * i: 6 * i + 6]
        q = [f"{w(a)} AND {w(b)}", f"({w(a)} OR {w(b)}) AND {w(c)}", f"{w(a)} AND NOT {w(b)}",
             f"({w(a)} OR {w(b)}) AND ({w(c)} OR {w(d)}) AND NOT ({w(e)} OR {w(f)})"][i % 4]
        leaves, prog = tools.query_compile(q)
        toks = [c1_corpus_tid(c1_corpus, s) for s in leaves]
        for algo, name in ((TFIDF, "TF-IDF"), (BM25, "BM25")):
            ref = ref_c1.search(q, limit=100, algo=name, fuzzymatch=False)
            ids, sc = c1_oracle.search(algo, 100, toks, prog)
            assert ref == list(zip(ids.tolist(), sc.tolist())), q


def c1_corpus_tid(corpus, s):
    if not hasattr(corpus, "_tid"):
        corpus._tid = {corpus.term(i + 1): i + 1 for i in range(corpus.n_terms)}
    return corpus._tid[s]


@needs_ref
def test_port_fuzzy_equals_reference(c1_corpus, c1_oracle, ref_c1):
    """idxterm_fuzzysearch of the compiled reference == the restated BK search."""
    lib = _oracle.ref()
    lib.idxterm_fuzzysearch.restype = C.c_void_p
    lib.idxterm_fuzzysearch.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    hits = 0
    for q in c1_corpus.fuzzy_terms(1500):
        term = lib.idxterm_fuzzysearch(ref_c1.h, q, len(q))
        ref_id = C.cast(term, C.POINTER(C.c_uint32))[0] if term else 0   # idxterm_t.id, index.h:43
        got, _, _, _ = c1_oracle.fuzzy(q)
        assert got == ref_id, q
        hits += ref_id != 0
    assert hits > 700


def test_port_fuzzy_candidate_list_equals_reference(c1_corpus, c1_oracle, ref_c1):
    """The whole candidate list, in order: what the reference's own
    bktree_search() pushes (ref src/algo/bktree.c:252-254, read through
    oracle/shims/probe_shim.c) == the restated BFS; and it is a subset of the
    brute-force <= 2 set (the half-open child range loses matches)."""
    lib = _oracle.ref()
    if not hasattr(lib, "nxsb_ref_fuzzy_list"):
        pytest.skip("oracle/_ref was built without the probe shim")
    lib.nxsb_ref_fuzzy_list.restype = C.c_size_t
    lib.nxsb_ref_fuzzy_list.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
    buf = np.zeros(8192, dtype=np.uint32)
    nonempty = lost = 0
    for q in c1_corpus.fuzzy_terms(600):
        n = lib.nxsb_ref_fuzzy_list(ref_c1.h, q, len(q), buf.ctypes.data, len(buf))
        _, cands, dists, _ = c1_oracle.fuzzy(q, cap=8192)
        assert n <= len(buf) and buf[:n].tolist() == cands.tolist(), q
        true_t, true_d = c1_oracle.fuzzy_true(q)
        truth = dict(zip(true_t.tolist(), true_d.tolist()))
        assert all(truth.get(int(t)) == int(d) for t, d in zip(cands, dists)), q
        nonempty += n > 0
        lost += len(truth) - n
    assert nonempty > 300 and lost > 0


# ---------------------------------------------------------------------------
# against the committed reference-made fixtures (no oracle/_ref needed)


def test_port_equals_committed_reference_goldens(c1_corpus, c1_oracle):
    """tests/golden/c1_reference.npz (made by tests/golden/make_golden.py from the
    compiled reference): ids, float scores and order -- ties included -- equal."""
    from _golden import Golden
    from nxsearch_b200 import tools

    g = Golden()
    g.check_corpus(c1_corpus)
    for family, queries, limit in (("or", g.or_queries, 10), ("bool", g.bool_queries, 100)):
        for i, q in enumerate(queries):
            leaves, prog = tools.query_compile(q)
            toks = [c1_corpus_tid(c1_corpus, s) for s in leaves]
            for algo, key in ((BM25, "bm25"), (TFIDF, "tfidf")):
                ids, sc = c1_oracle.search(algo, limit, toks, prog)
                assert list(zip(ids.tolist(), sc.tolist())) == g.results(family, key, i), (q, key)
    for q, pick in zip(g.fuzzy_queries, g.fuzzy_pick.tolist()):
        got, _, _, _ = c1_oracle.fuzzy(q)
        assert got == pick, q
    assert int((g.fuzzy_pick != 0).sum()) > 300
