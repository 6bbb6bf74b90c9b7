"""Image build from the bytes of an nxsdtmap file: the blocks
are copied to the device as they lie in the file (big-endian, ref
src/index/storage.h:64-133) and decoded there; the image must equal the one
built from host-converted pairs."""
import struct

import numpy as np
import pytest

from _oracle import BM25, TFIDF

pytestmark = pytest.mark.gpu


def block_table(raw: bytes):
    """(ids, lens, raw_off, raw_n) of the live blocks of a dtmap file image."""
    data_len = struct.unpack_from(">Q", raw, 8)[0]
    off, end = 32, 32 + data_len
    ids, lens, roff, rn = [], [], [], []
    while off < end:
        doc_id, dl, n = struct.unpack_from(">QII", raw, off)
        if doc_id and dl:
            ids.append(doc_id)
            lens.append(dl)
            roff.append(off + 16)
            rn.append(n)
        off += 16 + 8 * n
    return ids, lens, roff, rn


def same(a, b):
    (ca, ia, sa), (cb, ib, sb) = a, b
    assert np.array_equal(ca, cb)
    for q in range(len(ca)):
        n = int(ca[q])
        assert np.array_equal(ia[q, :n], ib[q, :n]) and np.array_equal(sa[q, :n], sb[q, :n]), q


def test_image_from_raw_dtmap_bytes_equals_image_from_pairs(c1_corpus, tmp_path):
    from nxsearch_b200 import engine
    from test_gpu_engine import c1_queries
    from test_gpu_segments import boolean_queries

    c = c1_corpus
    c.write(tmp_path / "nxsterms", tmp_path / "nxsdtmap")
    raw = (tmp_path / "nxsdtmap").read_bytes()
    ids, lens, roff, rn = block_table(raw)
    assert ids == [int(x) for x in c.doc_ids]
    a, b = engine.Engine(0), engine.Engine(0)
    a.load_corpus(c)
    b.load_dtmap(np.frombuffer(raw, dtype=np.uint8), ids, lens, roff, rn, c.n_terms, c.token_count, c.doc_count)
    assert np.array_equal(a.get_df(c.n_terms), b.get_df(c.n_terms))
    qs = c1_queries(c, 400) + boolean_queries(c, 60)
    for algo in (BM25, TFIDF):
        for k in (10, 100):
            batch = engine.Batch.from_lists(algo, k, qs)
            same(a.search(batch), b.search(batch))

    # a slice of the file as a delta segment: only that byte range is copied
    cut = 9000
    b.load_dtmap(np.frombuffer(raw, dtype=np.uint8), ids[:cut], lens[:cut], roff[:cut], rn[:cut],
                 c.n_terms, c.token_count, c.doc_count, c.term_df)
    b.load_dtmap(np.frombuffer(raw, dtype=np.uint8), ids[cut:], lens[cut:], roff[cut:], rn[cut:],
                 c.n_terms, c.token_count, c.doc_count, c.term_df, segment=True)
    batch = engine.Batch.from_lists(BM25, 10, qs)
    same(a.search(batch), b.search(batch))
    a.close()
    b.close()


def test_raw_blocks_with_a_wide_count_and_an_empty_shard():
    """A count >= 65536 is found on the device and switches the image to wide postings."""
    from nxsearch_b200 import engine

    docs = [(7, 70_003, [(1, 70_000), (2, 3)]), (9, 5, [(2, 5)]), (12, 4, [(1, 1), (3, 3)])]
    raw = bytearray(32)
    ids, lens, roff, rn, pairs, off = [], [], [], [], [], [0]
    for doc_id, dl, pl in docs:
        raw += struct.pack(">QII", doc_id, dl, len(pl))
        roff.append(len(raw))
        for t, cnt in pl:
            raw += struct.pack(">II", t, cnt)
            pairs += [t, cnt]
        ids.append(doc_id)
        lens.append(dl)
        rn.append(len(pl))
        off.append(off[-1] + len(pl))
    tokens = sum(lens)
    a, b = engine.Engine(0), engine.Engine(0)
    a.load_docs(ids, lens, off, pairs, 3, tokens, 3)
    b.load_dtmap(np.frombuffer(bytes(raw), dtype=np.uint8), ids, lens, roff, rn, 3, tokens, 3)
    qs = [([1], [0]), ([2, 1], [1, 0, -3]), ([3], [0]), ([2, 1], [1, 0, -2])]
    for algo in (BM25, TFIDF):
        batch = engine.Batch.from_lists(algo, 10, qs)
        ra, rb = a.search(batch), b.search(batch)
        same(ra, rb)
        assert list(ra[0]) == [2, 3, 1, 1]
    b.load_dtmap(np.zeros(8, dtype=np.uint8), [], [], [], [], 3, 0, 0)
    assert list(b.search(engine.Batch.from_lists(BM25, 10, qs))[0]) == [0, 0, 0, 0]
    a.close()
    b.close()
