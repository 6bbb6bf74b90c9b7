"""Host-side logic of the product library, CPU only: parser, JSON, params,
responses, on-disk format, API error behaviour, exported symbols."""
import ctypes as C
import json
import re
import shutil
import tempfile
from pathlib import Path

import numpy as np
import pytest

import _mk
import _oracle
from nxsearch_b200 import capi, tools, library_path

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture()
def nxs():
    base = tempfile.mkdtemp(prefix="nxsb_t_")
    n = capi.Nxs(base)
    n.base = base
    yield n
    n.close()
    shutil.rmtree(base, ignore_errors=True)


# ---- query language --------------------------------------------------------

@pytest.mark.parametrize("query,repr_,kinds", _mk.PARSER_CASES)
def test_parser_goldens(query, repr_, kinds):
    """ref tests/t_queryparser.c:27-115: token streams and IR s-expressions."""
    assert tools.query_lex(query) == kinds
    dump, err = tools.query_dump(query)
    assert dump == repr_
    if repr_ is None:
        assert err and err.startswith("syntax error near ")


def test_lexer_longest_match_rules():
    """re2c semantics (scan.re:64-119): keywords only when followed by a separator."""
    assert tools.query_lex("ANDROID ORacle NOTary") == ["FF", "FF", "FF"]
    assert tools.query_lex("a&b a & b") == ["FF", "FF", "AND", "FF"]
    assert tools.query_lex("'a b' 'ab'c 'unterminated") == ["QS", "FF", "FF"]
    assert tools.query_lex("a|b | Or oR") == ["FF", "OR", "OR", "OR"]
    assert tools.query_dump("a AND NOT b AND c")[0] == "(AND (NOT `a` `b`) `c`)"
    assert tools.query_dump("a b OR c")[0] == "(OR `a` (OR `b` `c`))"
    assert tools.query_dump("(a b)")[0] is None          # juxtaposition only at top level
    assert tools.query_dump("NOT a")[0] is None
    assert tools.query_dump("a OR NOT b")[0] is None
    assert tools.query_dump("")[0] is None
    assert tools.query_dump("a AND")[1] == 'syntax error near 1:5: " ..."'
    assert tools.query_dump("a\n\n) b")[1].startswith("syntax error near ")


def test_token_order_is_right_to_left_and_deduplicated():
    """ref query.c:89-103 (LIFO walk) + tokenizer.c:94-117 (dedup by string)."""
    assert tools.query_compile("a OR b OR c") == (["c", "b", "a"], [2, 1, -3, 0, -3])
    toks, prog = tools.query_compile("aa AND NOT (bb AND cc) OR aa")
    assert toks == ["aa", "cc", "bb"] and prog == [0, 2, 1, -2, -4, 0, -3]
    deep = " OR ".join(f"t{i}" for i in range(101))      # left-deep chain, depth 100: allowed
    assert tools.query_compile(deep) is not None
    assert tools.query_compile(deep + " OR x") is None    # depth 101 > NXS_QUERY_RLIMIT


# ---- JSON / params / response ----------------------------------------------

def test_params_roundtrip_and_first_match_semantics(nxs):
    lib = nxs._lib
    p = capi.Params(lib, algo="BM25", limit=7, fuzzymatch=False, filters=["normalizer", "stemmer"])
    doc = json.loads(p.tojson())
    assert doc == {"algo": "BM25", "limit": 7, "fuzzymatch": False, "filters": ["normalizer", "stemmer"]}
    p.set("algo", "TF-IDF")                               # yyjson_mut_obj_add appends; get returns the first
    assert '"algo": "BM25"' in p.tojson() and p.tojson().count('"algo"') == 2
    p.release()
    src = b'{"a": [1, -2, 3.5, true, null, "x\\u00e9\\n"], "b": {"c": 18446744073709551615}}'
    h = lib.nxs_params_fromjson(nxs.h, src, len(src))
    assert h
    assert json.loads(capi._take_string(lib.nxs_params_tojson(h, None))) == json.loads(src)
    lib.nxs_params_release(h)
    assert not lib.nxs_params_fromjson(nxs.h, b"{bad", 4)
    assert nxs.error()[0] == capi.ERR_SYSTEM


def test_real_number_formatting_matches_yyjson():
    lib = capi.bind()
    lib_fmt = C.CDLL(str(library_path()))
    cases = {1.5: "1.5", 3.0: "3.0", 0.1: "0.1", 1e21: "1e21", 1.25e-7: "1.25e-7", 123456789.0: "123456789.0",
             float(np.float32(1.1736002)): "1.1736001968383789", 0.000001: "0.000001", 1e-7: "1e-7"}
    # exercised through a response object: scores are floats widened to double
    for val in (1.5, 3.0, 0.25378534197807312, 1.1736001968383789, 100.0, 0.0009765625):
        src = json.dumps({"x": val}).encode()
        h = lib.nxs_params_fromjson(None, src, len(src))
        out = capi._take_string(lib.nxs_params_tojson(h, None))
        assert json.loads(out)["x"] == val
        lib.nxs_params_release(h)
    del lib_fmt, cases


# ---- on-disk format --------------------------------------------------------

TERMS_GOLDEN = bytes([
    0x4e, 0x58, 0x53, 0x5f, 0x54, 0x01, 0x00, 0x00, 0x00, 0x00, 0x00, 0x38, 0x00, 0x00, 0x00, 0x00,
    0x00, 0x0b, 0x73, 0x6f, 0x6d, 0x65, 0x2d, 0x74, 0x65, 0x72, 0x6d, 0x2d, 0x31, 0x00, 0x00, 0x00,
    0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x01, 0x00, 0x0e, 0x61, 0x6e, 0x6f, 0x74, 0x68, 0x65,
    0x72, 0x2d, 0x74, 0x65, 0x72, 0x6d, 0x2d, 0x32, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00,
    0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x02])   # ref tests/t_index_terms.c:23-37
DTMAP_GOLDEN = bytes([
    0x4e, 0x58, 0x53, 0x5f, 0x44, 0x01, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x38,
    0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x04, 0x00, 0x00, 0x00, 0x02, 0x00, 0x00, 0x00, 0x00,
    0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x03, 0xe9, 0x00, 0x00, 0x00, 0x03, 0x00, 0x00, 0x00, 0x02,
    0x00, 0x00, 0x00, 0x01, 0x00, 0x00, 0x00, 0x01, 0x00, 0x00, 0x00, 0x02, 0x00, 0x00, 0x00, 0x02,
    0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x03, 0xea, 0x00, 0x00, 0x00, 0x01, 0x00, 0x00, 0x00, 0x01,
    0x00, 0x00, 0x00, 0x03, 0x00, 0x00, 0x00, 0x01])   # ref tests/t_index_dtmap.c:25-41


def _golden_corpus():
    c = _mk.make_corpus([(1001, ["some-term-1", "another-term-2", "another-term-2"]), (1002, ["term-3"])])
    return c


def test_bulk_writer_reproduces_reference_golden_bytes(tmp_path):
    c = _golden_corpus()
    lib = tools._bind()
    cs = tools.CorpusStruct(
        c.n_docs, c.n_terms, c.n_pairs, c.token_count, c.doc_count,
        c.doc_ids.ctypes.data_as(C.POINTER(C.c_uint64)), c.doc_len.ctypes.data_as(C.POINTER(C.c_uint32)),
        c.doc_off.ctypes.data_as(C.POINTER(C.c_uint64)), c.pairs.ctypes.data_as(C.POINTER(C.c_uint32)),
        C.cast(C.create_string_buffer(c.term_blob), C.c_void_p), c.term_off.ctypes.data_as(C.POINTER(C.c_uint32)),
        c.term_total.ctypes.data_as(C.POINTER(C.c_uint64)), c.term_df.ctypes.data_as(C.POINTER(C.c_uint32)))
    # the terms golden holds only the first two terms (t_index_terms adds one token set)
    two = tools.CorpusStruct.from_buffer_copy(cs)
    two.n_terms = 2
    assert lib.nxsb_write_terms_file(str(tmp_path / "t").encode(), C.byref(two)) == 0
    assert lib.nxsb_write_dtmap_file(str(tmp_path / "d").encode(), C.byref(cs)) == 0
    assert (tmp_path / "t").read_bytes()[:len(TERMS_GOLDEN)] == TERMS_GOLDEN
    assert (tmp_path / "d").read_bytes()[:len(DTMAP_GOLDEN)] == DTMAP_GOLDEN
    assert (tmp_path / "t").stat().st_size == 32768 and (tmp_path / "d").stat().st_size == 32768


def test_reader_roundtrip_and_deletion_markers(tmp_path, c1_corpus):
    t, d = tmp_path / "nxsterms", tmp_path / "nxsdtmap"
    c1_corpus.write(t, d)
    back = tools.Corpus.read(t, d)
    assert back.n_docs == c1_corpus.n_docs and back.n_terms == c1_corpus.n_terms
    assert np.array_equal(back.pairs, c1_corpus.pairs) and np.array_equal(back.doc_off, c1_corpus.doc_off)
    assert np.array_equal(back.doc_len, c1_corpus.doc_len) and back.term_blob == c1_corpus.term_blob
    assert back.token_count == c1_corpus.token_count and back.doc_count == c1_corpus.doc_count
    assert np.array_equal(back.term_df, c1_corpus.term_df) and np.array_equal(back.term_total, c1_corpus.term_total)
    # delete document 3 the way idx_dtmap_remove does: zero the id in place, append {id, 0}
    raw = bytearray(d.read_bytes())
    data_len = int.from_bytes(raw[8:16], "big")
    off = 32
    for _ in range(2):
        off += 16 + 8 * int.from_bytes(raw[off + 12: off + 16], "big")
    assert int.from_bytes(raw[off: off + 8], "big") == 3
    raw[off: off + 8] = bytes(8)
    raw[32 + data_len: 32 + data_len + 16] = (3).to_bytes(8, "big") + bytes(8)
    raw[8:16] = (data_len + 16).to_bytes(8, "big")
    d.write_bytes(bytes(raw))
    gone = tools.Corpus.read(t, d)
    assert gone.n_docs == c1_corpus.n_docs - 1 and 3 not in set(gone.doc_ids[:10].tolist())


# ---- API behaviour ---------------------------------------------------------

def test_index_lifecycle_and_error_codes(nxs):
    idx = nxs.create_index("main")
    # the reference's defaults minus the stemmer, which this build does not have
    assert idx.params_json() == {"filters": ["normalizer", "stopwords"], "algo": "BM25", "lang": "en"}
    with pytest.raises(capi.NxsError) as e:
        nxs.create_index("main")
    assert e.value.code == capi.ERR_EXISTS
    with pytest.raises(capi.NxsError) as e:
        nxs.open_index("main")                             # one handle per name (nxs.c:391-395)
    assert e.value.code == capi.ERR_EXISTS
    with pytest.raises(capi.NxsError) as e:
        nxs.open_index("nope")
    assert e.value.code == capi.ERR_MISSING
    with pytest.raises(capi.NxsError) as e:
        nxs.create_index("bad name!")
    assert e.value.code == capi.ERR_INVALID

    idx.add(7, "Hello hello WORLD")
    with pytest.raises(capi.NxsError) as e:
        idx.add(7, "again")
    assert e.value.code == capi.ERR_EXISTS
    with pytest.raises(capi.NxsError) as e:
        idx.add(0, "zero id")
    assert e.value.code == capi.ERR_INVALID
    with pytest.raises(capi.NxsError) as e:
        idx.add(8, " ... ")
    assert e.value.code == capi.ERR_MISSING
    with pytest.raises(capi.NxsError) as e:
        idx.remove(99)
    assert e.value.code == capi.ERR_MISSING
    assert nxs._lib.nxs_luafilter_load(nxs.h, b"f", b"code") == -1 and nxs.error()[0] == capi.ERR_INVALID

    for q, kw, code in (("", {}, capi.ERR_INVALID), ("a AND", {}, capi.ERR_INVALID),
                        ("hello", {"limit": 0}, capi.ERR_INVALID), ("hello", {"algo": "PageRank"}, capi.ERR_INVALID)):
        with pytest.raises(capi.NxsError) as e:
            idx.search(q, **kw)
        assert e.value.code == code, (q, kw)
    # nothing resolves => empty result without touching the GPU (search.c:224-226)
    assert idx.search("unknownterm", fuzzymatch=False) == []
    idx.close()
    nxs.destroy_index("main")
    assert not (Path(nxs.base) / "data" / "main").exists()


def test_no_gpu_means_a_loud_error_not_a_fallback(nxs):
    from nxsearch_b200 import engine

    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    idx = nxs.create_index("x")
    idx.add(1, "alpha beta")
    with pytest.raises(capi.NxsError) as e:
        idx.search("alpha")
    assert e.value.code == capi.ERR_SYSTEM and "no CPU fallback" in e.value.msg
    with pytest.raises(RuntimeError):
        engine.Engine(0)


def test_a_c_program_compiles_against_the_header_and_links_the_library(nxs):
    """tests/c/nxs_caller.c builds with -Wall -Wextra -Werror against
    include/nxs.h and libnxsearch.so and runs; without a GPU every search is the
    loud NXS_ERR_SYSTEM, with one the answers are checked in test_gpu_capi.py."""
    import _ccaller
    from nxsearch_b200 import engine

    idx = nxs.create_index("c")
    idx.add(1, "alpha beta")
    idx.add(2, "beta gamma")
    idx.close()
    out = _ccaller.run(nxs.base, "c", "BM25", 10, ["alpha", "beta OR gamma"]).splitlines()
    assert len(out) == 2
    if engine.device_count() == 0:
        assert all(line.startswith("error 2 ") for line in out), out
    else:
        assert out[0].split()[0] == "1" and out[1].split()[0] == "2", out


@pytest.mark.parametrize("spec", ["0-3,5", "3-1", "0,,x"])
def test_gpu_device_list_is_parsed_or_ignored(spec, monkeypatch):
    """NXS_GPU_DEVICES takes ranges and lists; anything malformed leaves the
    single device of NXS_GPU_DEVICE.  Without a GPU the search then fails as
    loudly as ever (and with one, answers are checked in test_gpu_capi.py)."""
    from nxsearch_b200 import engine

    monkeypatch.setenv("NXS_GPU_DEVICES", spec)
    base = tempfile.mkdtemp(prefix="nxsb_t_")
    n = capi.Nxs(base)
    try:
        idx = n.create_index("d")
        idx.add(1, "alpha beta")
        if engine.device_count() == 0:
            with pytest.raises(capi.NxsError) as e:
                idx.search("alpha")
            assert e.value.code == capi.ERR_SYSTEM
        idx.close()
    finally:
        n.close()
        shutil.rmtree(base, ignore_errors=True)


def test_tokenizer_matches_the_reference_goldens():
    """Word segmentation and the normalizer against the reference's own cases
    (ref src/tests/t_tokenize.c:17-62, src/tests/t_utf8.c:70-74), plus the
    UAX #29 rules they imply for digits and what stays out of reach here."""
    for text, want in _mk.TOKENIZE_CASES:
        assert [w for w, _ in tools.tokenize(text)] == want, text
    assert tools.tokenize("The quick brown fox jumped over the lazy dog.")[0] == ("the", 2)
    for text, want in _mk.NORMALIZER_CASES:
        assert tools.tokenize(text) == [(want, 1)]
    # NFKD + "[:Nonspacing Mark:] Remove; Latin-ASCII" (ref utf8.c:30-31) over the two-byte blocks
    assert [w for w, _ in tools.tokenize("Київ ДНІПР ЁЖИК Ελλάδα ΆΝΘΡΩΠΟΣ ÅNGSTRÖM straße Þór œuvre Łódź İstanbul")] == \
        ["киів", "дніпр", "ежик", "ελλαδα", "ανθρωποσ", "angstrom", "strasse", "thor", "oeuvre", "lodz", "istanbul"]
    assert tools.tokenize("cafe\u0301 café") == [("cafe", 2)]       # decomposed and precomposed meet
    assert tools.tokenize("Henry Ⅷ") == [("henry", 1), ("Ⅷ", 1)]    # no general compatibility decomposition (t_utf8.c:93)
    # ... but the forms that fold in place do: full-width letters and digits, Latin ligatures
    assert [w for w, _ in tools.tokenize("ＡＢＣ１２３ｘｙｚ ﬁnance oﬃce ﬆop")] == ["abc123xyz", "finance", "office", "stop"]
    # WB11/12: digits join over . , ; ' -- letters only over . and '
    assert [w for w, _ in tools.tokenize("pi is 3.14, or 1,000;5 a,b a;b x.y 1.a a.1 v2.0")] == \
        ["pi", "is", "3.14", "or", "1,000;5", "a", "b", "x.y", "1", "v2.0"]
    # a segment without a letter or a digit is not a word; '_' joins (WB13a/b); ':' and '-' never do
    assert [w for w, _ in tools.tokenize("_ __ _x_ year-end 10:30 a:b 'quoted' ..")] == \
        ["_x_", "year", "end", "10", "30", "a", "b", "quoted"]
    # without the normalizer the case stays
    assert tools.tokenize("Some.Text", normalize=False) == [("Some.Text", 1)]
    # the case the reference keeps out of its own run (t_tokenize.c:64-79): ICU keeps the
    # underscores and "some.text" in one piece, and an emoji is not a word
    assert [w for w, _ in tools.tokenize("_underscore_, year-end, join--double Some.Text, 🥎")] == \
        ["_underscore_", "year", "end", "join", "double", "some.text"]
    # punctuation outside ASCII breaks words like ASCII punctuation does; U+2019 and U+00B7
    # join letters (WB6/7) and the apostrophe comes out in ASCII (Latin-ASCII)
    assert [w for w, _ in tools.tokenize("“Quoted” text — with dashes…and ellipsis")] == \
        ["quoted", "text", "with", "dashes", "and", "ellipsis"]
    assert [w for w, _ in tools.tokenize("doesn’t l’été rock’n’roll ‘quoted’ a·b")] == \
        ["doesn't", "l'ete", "rock'n'roll", "quoted", "a·b"]
    assert [w for w, _ in tools.tokenize("10\u00a0000 €5 café\u00a0au lait x→y 3×4 ½ ① ＡＢＣ！ｄ")] == \
        ["10", "000", "5", "cafe", "au", "lait", "x", "y", "3", "4", "abc", "d"]
    # no dictionary here: a run of ideographs or kana is one word, CJK punctuation still breaks it
    assert [w for w, _ in tools.tokenize("日本語のテキスト。次の文")] == ["日本語のテキスト", "次の文"]
    # malformed UTF-8 neither crashes nor disappears: stray bytes are letters of one byte
    assert len(tools.tokenize(b"bad \xff\xfe bytes \xc3 \xe2\x80 end\xe2")) == 6


def test_index_limits_of_the_reference(nxs):
    """ref src/tests/t_index_limits.c: one document with 65 535 distinct terms
    (ids in order of first appearance, every count 1, the header counters
    follow), a term of 65 535 bytes is stored, one byte more is NXS_ERR_LIMIT
    with the reference's message."""
    idx = nxs.create_index("l", filters=["normalizer"])
    n = 0xffff
    words = [f"w{i:07d}" for i in range(n)]
    idx.add(1001, " ".join(words))
    idx.add(1002, "a" * 0xffff)
    with pytest.raises(capi.NxsError) as e:
        idx.add(1003, "b" * 0x10000)
    assert e.value.code == capi.ERR_LIMIT and e.value.msg == "term too long (65536)"
    idx.close()
    c = tools.Corpus.read(f"{nxs.base}/data/l/nxsterms", f"{nxs.base}/data/l/nxsdtmap")
    assert c.n_terms == n + 1 and c.n_docs == 2 and c.token_count == n + 1
    assert c.term(1) == words[0] and c.term(n) == words[-1] and c.term(n + 1) == "a" * 0xffff
    first = c.pairs[: 2 * n].reshape(n, 2)
    assert int(c.doc_len[0]) == n and np.array_equal(first[:, 0], np.arange(1, n + 1)) and np.all(first[:, 1] == 1)
    # a new term starts at the adding document's count (ref terms.c:260) and the document
    # is then counted again (ref dtmap.c:236): the files say 2, as the reference's do
    assert np.all(c.term_df[:n] == 1) and np.all(c.term_total == 2)


def test_gpu_layout_is_replicas_or_shards_nothing_else(monkeypatch):
    """NXS_GPU_LAYOUT picks what several devices hold (replicas of the image, or
    a range of the documents each).  A misspelt value fails nxs_open instead of
    quietly choosing the other layout."""
    base = tempfile.mkdtemp(prefix="nxsb_t_")
    try:
        for ok in ("replicas", "shards"):
            monkeypatch.setenv("NXS_GPU_LAYOUT", ok)
            capi.Nxs(base).close()
        monkeypatch.setenv("NXS_GPU_LAYOUT", "shard")
        with pytest.raises(OSError):
            capi.Nxs(base)
    finally:
        shutil.rmtree(base, ignore_errors=True)


def test_index_files_are_interchangeable_with_the_reference(nxs, c1_corpus):
    """Files written through nxs_index_add here open in the compiled reference
    and vice versa, byte-identical for the same sequence of adds."""
    ref = _oracle.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    texts = [(5, "the cat sat on the mat"), (9, "a cat and a dog"), (2, "dog eat dog world"), (11, "mat")]
    # the reference's own tokenizer cases: punctuation, acronyms, snake case go the same way on both sides
    texts += [(20 + i, t) for i, (t, _) in enumerate(_mk.TOKENIZE_CASES) if t]
    texts.append((40, "pi is 3.14, or 1,000;5 -- v2.0 at 10:30, year-end _x_ 'quoted'"))
    ours = nxs.create_index("a")
    for d, t in texts:
        ours.add(d, t)
    ours.remove(9)
    ours.close()
    rbase = tempfile.mkdtemp(prefix="nxsb_r_")
    try:
        rn = capi.Nxs(rbase, lib=ref)
        theirs = rn.create_index("a")
        for d, t in texts:
            theirs.add(d, t)
        theirs.remove(9)
        theirs.close()
        for f in ("nxsterms", "nxsdtmap"):
            assert (Path(nxs.base) / "data/a" / f).read_bytes() == (Path(rbase) / "data/a" / f).read_bytes(), f
        # and the reference searches an index this library wrote
        shutil.copy(Path(nxs.base) / "data/a/nxsterms", Path(rbase) / "data/a/nxsterms")
        shutil.copy(Path(nxs.base) / "data/a/nxsdtmap", Path(rbase) / "data/a/nxsdtmap")
        again = rn.open_index("a")
        assert sorted(d for d, _ in again.search("dog OR mat")) == [2, 5, 11, 21]
        assert [d for d, _ in again.search("'i.b.m'")] == [22]
        again.close()
        rn.close()
    finally:
        shutil.rmtree(rbase, ignore_errors=True)


def test_stemmer_is_english_or_refused_never_passed_through_silently(nxs, monkeypatch):
    """a pipeline naming `stemmer' must not silently index unstemmed
    terms.  English is the restated Snowball algorithm (stem_en.c); any other
    language fails with NXS_ERR_INVALID unless the caller opts in to the
    identity stemmer."""
    monkeypatch.delenv("NXSB_STEMMER_PASSTHROUGH", raising=False)
    idx = nxs.create_index("s", filters=["normalizer", "stemmer"])
    idx.add(1, "The quick brown foxes jumped over the lazy dogs, generously")
    idx.close()
    terms = tools.Corpus.read(f"{nxs.base}/data/s/nxsterms", f"{nxs.base}/data/s/nxsdtmap")
    assert [terms.term(t) for t in range(1, terms.n_terms + 1)] == \
        ["the", "quick", "brown", "fox", "jump", "over", "lazi", "dog", "generous"]
    # an index the reference created with ITS defaults names the stemmer too: it opens
    (Path(nxs.base) / "data/s/params.db").write_text(
        json.dumps({"filters": ["normalizer", "stopwords", "stemmer"], "algo": "BM25", "lang": "en"}))
    idx = nxs.open_index("s")
    assert idx.params_json()["filters"] == ["normalizer", "stopwords", "stemmer"]
    idx.close()
    # ... but not for a language whose algorithm is not here
    with pytest.raises(capi.NxsError) as e:
        nxs.create_index("d", filters=["normalizer", "stemmer"], lang="de")
    assert e.value.code == capi.ERR_INVALID and "stemmer" in str(e.value)
    (Path(nxs.base) / "data/s/params.db").write_text(
        json.dumps({"filters": ["normalizer", "stopwords", "stemmer"], "algo": "BM25", "lang": "de"}))
    with pytest.raises(capi.NxsError) as e:
        nxs.open_index("s")
    assert e.value.code == capi.ERR_INVALID
    monkeypatch.setenv("NXSB_STEMMER_PASSTHROUGH", "1")
    idx = nxs.open_index("s")
    assert idx.params_json()["filters"] == ["normalizer", "stopwords", "stemmer"]
    idx.close()


def test_english_stemmer_on_the_published_vocabulary():
    """The Snowball english algorithm as published (snowballstem.org): the
    worked examples of its description and the sample vocabulary shown beside
    it, through the same filter the index uses."""
    pairs = """
    caresses caress ponies poni ties tie caress caress cats cat feed feed agreed agre plastered plaster
    bled bled motoring motor sing sing conflated conflat troubled troubl sized size hopping hop tanned tan
    falling fall hissing hiss fizzed fizz failing fail filing file happy happi sky sky
    relational relat conditional condit rational ration valenci valenc digitizer digit conformabli conform
    radicalli radic differentli differ vileli vile analogousli analog vietnamization vietnam
    predication predic operator oper feudalism feudal decisiveness decis hopefulness hope callousness callous
    formaliti formal sensitiviti sensit sensibiliti sensibl triplicate triplic formative format formalize formal
    electriciti electr electrical electr hopeful hope goodness good revival reviv allowance allow
    inference infer airliner airlin gyroscopic gyroscop adjustable adjust defensible defens irritant irrit
    replacement replac adjustment adjust dependent depend adoption adopt homologous homolog activate activ
    angulariti angular effective effect bowdlerize bowdler probate probat rate rate cease ceas
    controll control roll roll
    generate generat generates generat generated generat generating generat general general
    generally general generic generic generically generic generous generous generously generous
    consign consign consigned consign consigning consign consignment consign consist consist
    consisted consist consistency consist consistent consist consistently consist consisting consist
    consists consist consolation consol consolations consol consolatory consolatori console consol
    consoled consol consoles consol consolidate consolid consolidated consolid consolidating consolid
    consoling consol consolingly consol consols consol consonant conson consort consort consorted consort
    consorting consort conspicuous conspicu conspicuously conspicu conspiracy conspiraci
    conspirator conspir conspirators conspir conspire conspir conspired conspir conspiring conspir
    constable constabl constables constabl constance constanc constancy constanc constant constant
    knack knack knackeries knackeri knacks knack knag knag knave knave knaves knave knavish knavish
    kneaded knead kneading knead knee knee kneel kneel kneeled kneel kneeling kneel kneels kneel
    knees knee knell knell knelt knelt knew knew knick knick knif knif knife knife knight knight
    knightly knight knights knight knit knit knits knit knitted knit knitting knit knives knive
    knob knob knobs knob knock knock knocked knock knocker knocker knockers knocker knocking knock
    knocks knock knopp knopp knot knot knots knot
    skis ski skies sky dying die lying lie tying tie idly idl gently gentl ugly ugli early earli
    only onli singly singl news news howe howe atlas atlas cosmos cosmos bias bias andes andes
    inning inning outing outing canning canning herring herring earring earring proceed proceed
    exceed exceed succeed succeed
    cries cri tied tie gas gas this this gaps gap kiwis kiwi luxuriated luxuri hoped hope hoping hope
    by by say say cry cri boy's boy owed owe foxes fox jumped jump lazy lazi
    argument argument arguing argu agreement agreement university univers national nation
    running run happiness happi organization organ easily easili fairly fair something someth
    beautiful beauti flies fli died die yields yield playing play
    """.split()
    assert len(pairs) % 2 == 0 and len(pairs) > 400
    for word, want in zip(pairs[0::2], pairs[1::2]):
        assert tools.tokenize(word, stem=True) == [(want, 1)], word
    # two letters or fewer stay; the normalizer runs first; other bytes are consonants
    assert [w for w, _ in tools.tokenize("As IS Foxes ąžuolas 12345s", stem=True)] == \
        ["as", "is", "fox", "azuola", "12345s"]


def test_every_declared_symbol_is_exported():
    """include/nxs.h + nxsb200_gpu.h <-> libnxsearch.so (the product),
    include/nxsb200_tools.h <-> libnxsb_tools.so (test and benchmark tooling);
    the product library carries none of the tooling."""
    from nxsearch_b200._lib import load_tools_library

    lib = C.CDLL(str(library_path()))
    tools_lib = load_tools_library()

    def declared_in(name):
        text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / name).read_text(), flags=re.S)
        return set(re.findall(r"\b(nxsb?_[a-z0-9_]+)\s*\(", text))

    declared = declared_in("nxs.h") | declared_in("nxsb200_gpu.h")
    assert len(declared) >= 55
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    on_handles = {"nxsb_index_term_df", "nxsb_index_image_stats", "nxsb_resp_collect", "nxsb_bkmirror_build"}
    tooling = declared_in("nxsb200_tools.h")
    assert len(tooling) >= 10
    for s in sorted(tooling):
        # introspection of live handles stays with the library that owns them
        assert hasattr(lib if s in on_handles else tools_lib, s), s
    for s in sorted(tooling - on_handles):
        assert not hasattr(lib, s), f"{s} leaked into the product library"
    # the reference's 25 public symbols
    for s in ("nxs_open nxs_close nxs_luafilter_load nxs_get_error nxs_params_create nxs_params_fromjson "
              "nxs_params_set_strlist nxs_params_set_str nxs_params_set_uint nxs_params_set_bool nxs_params_tojson "
              "nxs_params_release nxs_index_create nxs_index_destroy nxs_index_get_params nxs_index_open "
              "nxs_index_close nxs_index_add nxs_index_remove nxs_index_search nxs_resp_iter_reset "
              "nxs_resp_iter_result nxs_resp_resultcount nxs_resp_tojson nxs_resp_release").split():
        assert hasattr(lib, s), s


def test_bk_mirror_reproduces_the_reference_search(c1_corpus, c1_oracle):
    """parent/edge/rank arrays + exact distances == the oracle's BK-tree BFS
    (host-side emulation of what the fuzzy kernel computes)."""
    parent, edge, rank = c1_corpus.bk_mirror()
    lev = _oracle.port().ora_levdist
    terms = [c1_corpus.term(i + 1).encode() for i in range(c1_corpus.n_terms)]
    by_len = {}
    for i, t in enumerate(terms):
        by_len.setdefault(len(t), []).append(i)
    for q in c1_corpus.fuzzy_terms(60, seed=5):
        want, cands, _, _ = c1_oracle.fuzzy(q)
        best = None
        for L in range(len(q) - 2, len(q) + 3):
            for t in by_len.get(L, []):
                # idxterm.c:239: only terms with a non-zero total are picked
                if c1_corpus.term_total[t] == 0 or lev(q, len(q), terms[t], len(terms[t])) > 2:
                    continue
                c, ok = t, True
                while parent[c] != 0xFFFFFFFF:
                    p = int(parent[c])
                    d = lev(q, len(q), terms[p], len(terms[p]))
                    if not (max(d - 2, 0) <= edge[c] < min(d + 2, 63)):
                        ok = False
                        break
                    c = p
                if ok and (best is None or rank[t] < rank[best]):
                    best = t
        assert (best + 1 if best is not None else 0) == want, q


def _doc_terms(corpus, i):
    lo, hi = int(corpus.doc_off[i]), int(corpus.doc_off[i + 1])
    return set(int(t) for t in corpus.pairs[2 * lo: 2 * hi: 2])


def test_bulk_threaded_sync_equals_block_by_block(tmp_path, c1_corpus, monkeypatch):
    """Opening a large index checks term ids and counts df[] on host threads
    (index.c dtmap_sync_bulk); the state must equal the block-by-block sync,
    with deleted blocks, deletion markers and an unknown term id in the way."""
    base = tmp_path / "base"
    base.mkdir()
    nxs = capi.Nxs(str(base))
    nxs.create_index("b", filters=["normalizer"]).close()
    c1_corpus.write(base / "data/b/nxsterms", base / "data/b/nxsdtmap")
    w = nxs.open_index("b")
    for d in (3, 77, 4000, 9999):
        w.remove(int(c1_corpus.doc_ids[d]))
    w.add(2**40, " ".join(c1_corpus.term(t) for t in (5, 5, 9, 1234)))
    w.add(int(c1_corpus.doc_ids[77]), c1_corpus.term(42))          # an id comes back
    w.close()

    def state(min_bytes):
        monkeypatch.setenv("NXSB_BULK_MIN_BYTES", str(min_bytes))
        n = capi.Nxs(str(base))
        i = n.open_index("b")
        st = i.image_stats()
        df = np.array([i.term_df(t) for t in range(1, c1_corpus.n_terms + 1)], dtype=np.int64)
        i.close()
        n.close()
        return st, df

    serial, df_serial = state(1 << 60)
    bulk, df_bulk = state(1)
    assert serial == bulk and serial["live"] == c1_corpus.n_docs - 4 + 2
    assert np.array_equal(df_serial, df_bulk) and df_bulk.sum() > 0
    # against the file reader's own recount
    back = tools.Corpus.read(base / "data/b/nxsterms", base / "data/b/nxsdtmap")
    assert np.array_equal(df_bulk, np.asarray(back.term_df, dtype=np.int64))

    # a bulk open leaves the id -> slot map to the first writer: it must still know every document
    monkeypatch.setenv("NXSB_BULK_MIN_BYTES", "1")
    n = capi.Nxs(str(base))
    i = n.open_index("b")
    with pytest.raises(capi.NxsError) as ei:
        i.add(int(c1_corpus.doc_ids[10]), "anything")
    assert ei.value.code == 4                      # NXS_ERR_EXISTS
    with pytest.raises(capi.NxsError) as ei:
        i.remove(int(c1_corpus.doc_ids[3]))         # removed before
    assert ei.value.code == 5                      # NXS_ERR_MISSING
    i.remove(int(c1_corpus.doc_ids[10]))
    i.add(int(c1_corpus.doc_ids[10]), c1_corpus.term(7))
    assert i.image_stats()["live"] == serial["live"]
    assert i.term_df(7) == df_serial[6] + 1 - (1 if 7 in _doc_terms(c1_corpus, 10) else 0)
    i.close()
    n.close()

    # a block naming a term the vocabulary does not have yet: both paths refuse it the same way
    raw = bytearray((base / "data/b/nxsdtmap").read_bytes())
    raw[32 + 16: 32 + 20] = (c1_corpus.n_terms + 7).to_bytes(4, "big")
    (base / "data/b/nxsdtmap").write_bytes(bytes(raw))
    for mb in (1 << 60, 1):
        monkeypatch.setenv("NXSB_BULK_MIN_BYTES", str(mb))
        n = capi.Nxs(str(base))
        i = n.open_index("b")          # partial sync at open: stops at the block, no error
        assert i.image_stats()["live"] == 0
        i.close()
        n.close()
    nxs.close()


def test_df_stays_current_through_adds_removes_and_a_second_writer(tmp_path):
    """The whole-index df[] the GPU segments score with is kept on the
    host as blocks and deletion markers are consumed; after any mix of
    nxs_index_add / nxs_index_remove from two handles it must equal a recount of
    the files (what the reference's per-term bitmap cardinality would be)."""
    rng = np.random.default_rng(17)
    words = [f"w{i}" for i in range(60)]
    base = tmp_path / "b"
    base.mkdir()
    a, b = capi.Nxs(str(base)), capi.Nxs(str(base))
    ia = a.create_index("x", filters=["normalizer"])
    ib = b.open_index("x")
    live, next_id = set(), 1

    def check(idx):
        back = tools.Corpus.read(base / "data/x/nxsterms", base / "data/x/nxsdtmap")
        # a handle learns of foreign changes when it syncs: adding/removing does that
        got = np.array([idx.term_df(t) for t in range(1, back.n_terms + 1)], dtype=np.int64)
        assert np.array_equal(got, np.asarray(back.term_df, dtype=np.int64))
        assert idx.image_stats()["live"] == len(live) == back.n_docs

    for step in range(120):
        idx = ia if rng.random() < 0.5 else ib
        if live and rng.random() < 0.35:
            d = int(rng.choice(sorted(live)))
            idx.remove(d)
            live.discard(d)
        else:
            n = int(rng.integers(1, 12))
            idx.add(next_id, " ".join(rng.choice(words, n)))
            live.add(next_id)
            next_id += 1
        if step % 10 == 9:
            check(idx)              # the handle that wrote last is in sync
    # the other handle catches up on its next write
    ia.add(next_id, "w1 w2")
    live.add(next_id)
    check(ia)
    ib.add(next_id + 1, "w3")
    live.add(next_id + 1)
    check(ib)
    for h in (ia, ib):
        h.close()
    a.close()
    b.close()
