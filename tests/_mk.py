"""Test helper: build a Corpus-like object (numpy arrays) from small documents,
assigning term ids in order of first appearance as nxs_index_add does."""
from __future__ import annotations

import re
from types import SimpleNamespace

import numpy as np

WORD = re.compile(rb"[A-Za-z0-9\x80-\xff]+")


def words(text: str | bytes) -> list[bytes]:
    b = text.encode() if isinstance(text, str) else text
    return [w.lower() for w in WORD.findall(b)]


def make_corpus(docs, vocabulary=None):
    """docs: iterable of (doc_id, text or list of words).  Returns a namespace with
    the fields of nxsearch_b200.tools.Corpus that the oracle/engine bindings read."""
    vocab: dict[bytes, int] = {}
    for w in vocabulary or []:
        vocab.setdefault(w if isinstance(w, bytes) else w.encode(), len(vocab) + 1)
    doc_ids, doc_len, doc_off, pairs = [], [], [0], []
    totals: dict[int, int] = {}
    for did, text in docs:
        ws = words(text) if isinstance(text, (str, bytes)) else [w if isinstance(w, bytes) else w.encode() for w in text]
        counts: dict[int, int] = {}
        for w in ws:
            t = vocab.setdefault(w, len(vocab) + 1)
            counts[t] = counts.get(t, 0) + 1
        for t in sorted(counts):
            pairs += [t, counts[t]]
            totals[t] = totals.get(t, 0) + counts[t]
        doc_ids.append(did)
        doc_len.append(len(ws))
        doc_off.append(len(pairs) // 2)
    terms = sorted(vocab, key=vocab.get)
    term_off = np.zeros(len(terms) + 1, dtype=np.uint32)
    term_off[1:] = np.cumsum([len(t) for t in terms])
    df = np.zeros(len(terms), dtype=np.uint32)
    for t in pairs[0::2]:
        df[t - 1] += 1
    c = SimpleNamespace(
        n_docs=len(doc_ids), n_terms=len(terms), n_pairs=len(pairs) // 2,
        token_count=int(sum(doc_len)), doc_count=len(doc_ids),
        doc_ids=np.array(doc_ids, dtype=np.uint64), doc_len=np.array(doc_len, dtype=np.uint32),
        doc_off=np.array(doc_off, dtype=np.uint64), pairs=np.array(pairs, dtype=np.uint32),
        term_off=term_off, term_blob=b"".join(terms),
        term_total=np.array([totals.get(i + 1, 0) for i in range(len(terms))], dtype=np.uint64),
        term_df=df, vocab=vocab,
    )
    c.term = lambda tid: terms[tid - 1].decode()
    c.tid = lambda w: vocab.get(w.encode() if isinstance(w, str) else w, 0)
    return c


# The reference's own scoring fixtures (src/tests/t_scoring.c:16-158).  Cases 2
# and 3 need the stemmer ("foxes" -> "fox"): STEMMED_SCORING_CASES below.
SCORING_CASES = [
    # (docs, query, {doc_id: (tfidf, bm25)})
    ([(1, "The quick brown fox jumped over the lazy dog"),
      (2, "Once upon a time there were three little foxes")],
     "dog", {1: (1.1736, 0.253785)}),
    ([(1, "cat dog rat"), (2, "cat cat dog")],
     "cat", {1: (0.693147, 0.066754), 2: (1.098612, 0.087140)}),
    ([(1, "cat cat dog dog"), (2, "dog dog cat cat"), (3, "cat dog rat cow"), (4, "cat dog rat bat")],
     "cat dog rat cow",
     {1: (2.197225, 0.100713), 2: (2.197225, 0.100713), 3: (4.213948, 0.771754), 4: (2.559895, 0.330938)}),
    ([(1, "aa " * 20), (2, "aa " * 10 + "bb " * 10), (3, "aa " + "bb " * 19)],
     "aa", {1: (3.044523, 0.095780), 2: (2.397895, 0.088995), 3: (0.693147, 0.048890)}),
    ([(1, "This is a very long document about the cats "
          "All kind of cats including the tabby and other cats"),
      (2, "cats cats cats"), (3, "cats cats dogs")],
     "cats", {1: (1.386294, 0.048411), 2: (1.386294, 0.091469), 3: (1.098612, 0.084499)}),
]

# src/tests/t_scoring.c:16-68, the two cases that need the stemmer ("fox" finds "foxes"):
# run with filters normalizer + stopwords + stemmer, the reference's defaults.
STEMMED_SCORING_CASES = [
    ([(1, "The quick brown fox jumped over the lazy dog"),
      (2, "Once upon a time there were three little foxes")],
     "fox", {1: (0.693147, 0.066754), 2: (0.693147, 0.066754)}),
    ([(1, "The quick brown fox jumped over the lazy dog"),
      (2, "Once upon a time there were three little foxes")],
     "fox dog", {1: (1.1736 + 0.693147, 0.253785 + 0.066754), 2: (0.693147, 0.066754)}),
]

# src/tests/t_querylogic.c:16-52
LOGIC_DOCS = [
    (1, "Textbook about Erlang in Linux environment"),
    (2, "Unix Shell scripting textbook"),
    (3, "Erlang and Python examples"),
    (4, "Textbook about Python using Linux and Windows"),
    (5, "All but NOT: Textbook Erlang Python Shell Linux Unix Java"),
    (6, "All keywords: Textbook Erlang Python Shell Linux Unix"),
]
LOGIC_CASES = [
    ("non-existant-term", []),
    ("unix", [2, 5, 6]),
    ("textbook AND (Erlang OR Python OR Shell) AND (Linux OR Unix) AND NOT (Windows OR Java)", [1, 2, 6]),
]

# src/tests/t_levdist.c:32-66
LEVDIST_CASES = [
    ("kitten", "kitten", 0), ("kitten", "sitten", 1), ("sitting", "kitten", 3), ("cat", "chat", 1),
    ("cat", "cactus", 3), ("cat", "gato", 2), ("", "", 0), ("", "a", 1), ("a", "", 1), ("a", "b", 1),
    ("aba", "a", 2), ("aabcc", "bccdd", 4), ("ab", "ac", 1), ("ac", "bc", 1), ("abc", "axc", 1),
    ("abc", "def", 3), ("aabbcd", "aabcd", 1), ("aabcd", "aabbcd", 1), ("aaabccc", "", 7),
    ("ABCDEF", "abcdef", 6), ("ABCDEF", "AbCdEf", 3), ("hello", "hallo", 1), ("variable", "valuable", 2),
    ("leaf", "leaves", 3), ("ab?cd?ef?", "!ab!cd!ef!", 4), ("john smith", "johnathan smith", 5),
    ("levenshtein", "frankenstein", 6), ("123456789", "101010101", 8), ("something", "different", 8),
]

# src/tests/t_bktree.c:24-57
BK_WORDS = ["the", "quick", "brown", "fox", "jumped", "over", "lazy", "dog"]
BK_SEARCH = ["teh", "qvick", "brawn", "fox", "jumps", "ovr", "llazy", "dog"]

# src/tests/t_queryparser.c:27-115 (query, s-expression or None for a syntax error, token kinds)
FF, QS, OR_, AND_, NOT_, BO, BC = "FF", "QS", "OR", "AND", "NOT", "(", ")"
PARSER_CASES = [
    ("A", "`A`", [FF]),
    ("(A OR B) AND C", "(AND (OR `A` `B`) `C`)", [BO, FF, OR_, FF, BC, AND_, FF]),
    ("A OR (B AND C)", "(OR `A` (AND `B` `C`))", [FF, OR_, BO, FF, AND_, FF, BC]),
    ("A OR B AND C", "(OR `A` (AND `B` `C`))", [FF, OR_, FF, AND_, FF]),
    ("A and not B", "(NOT `A` `B`)", [FF, AND_, NOT_, FF]),
    (" \"sp ace\" OR 'quo\\'te' OR ąžuolas OR 🇬🇧🇺🇸 AND Київ OR (1 AND NOT (  2   OR   3 ))",
     "(OR (OR (OR (OR `sp ace` `quo\\'te`) `ąžuolas`) (AND `🇬🇧🇺🇸` `Київ`)) (NOT `1` (OR `2` `3`)))",
     [QS, OR_, QS, OR_, FF, OR_, FF, AND_, FF, OR_, BO, FF, AND_, NOT_, BO, FF, OR_, FF, BC, BC]),
    ("a AND", None, [FF, AND_]),
    ("a b OR (c OR d) AND (e", None, [FF, FF, OR_, BO, FF, OR_, FF, BC, AND_, BO, FF]),
    ("A\nand\nB", "(AND `A` `B`)", [FF, AND_, FF]),
]

# src/tests/t_tokenize.c:17-62 (text, tokens in first-seen order; "normalizer" only).
# The reference keeps a seventh case out of its own run (t_tokenize.c:64-79, "TODO").
TOKENIZE_CASES = [
    ("a", ["a"]),
    ("The quick brown fox jumped over the lazy dog.",
     ["the", "quick", "brown", "fox", "jumped", "over", "lazy", "dog"]),
    ("We will play with I.B.M.", ["we", "will", "play", "with", "i.b.m"]),
    ("Hello_I_m_arbitrary_concatenated, foo and bar", ["hello_i_m_arbitrary_concatenated", "foo", "and", "bar"]),
    ("the [client] is <foo>, some *bold* marks.", ["the", "client", "is", "foo", "some", "bold", "marks"]),
    ("Text,which doesn't  have spaces right;one;two;three..",
     ["text", "which", "doesn't", "have", "spaces", "right", "one", "two", "three"]),
    ("", []),
]

# The "normalizer" filter = utf8_normalize (NFKC case folding) then utf8_subs_diacritics
# (src/core/filters_builtin.c:55-74).  The reference's cases for the two halves
# (src/tests/t_utf8.c:70-74 lower-casing, :124-127 diacritics), composed in that order:
NORMALIZER_CASES = [
    ("TEST", "test"),
    ("ĄČĘĖĮŠŲŪŽ", "aceeisuuz"),         # lower: "ąčęėįšųūž", then without the marks
    ("azúl", "azul"),
    ("ĄŽUOLĖLIS", "azuolelis"),          # diacritics: "AZUOLELIS", lower-cased first
    ("Fuglafjørður", "fuglafjordur"),
    ("Árbæ", "arbae"),
]
