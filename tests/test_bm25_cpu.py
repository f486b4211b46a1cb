"""BM25 oracle known-answer checks, host-side token plumbing and C-ABI argument handling - no GPU needed.

bm25s (the reference's third-party BM25, pinned 0.1.10) is not available here: the oracle restates its published
algorithm (parity unpinned).  These tests pin the restatement against (a) hand-computed values of the published
formulas and (b) an independent vectorised scipy construction of the same CSC score matrix."""
import ctypes
import math

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import bm25_oracle as B


def _scipy_index(corpus, n_vocab, k1=1.2, b=0.75):
    """Independent construction: (doc, token, tf) triples -> scipy CSC, the way bm25s assembles its matrix."""
    n = len(corpus)
    lens = np.array([len(d) for d in corpus], dtype=np.float64)
    l_avg = lens.mean()
    rows, cols, tfs = [], [], []
    for d, toks in enumerate(corpus):
        u, c = np.unique(np.asarray(toks, dtype=np.int64), return_counts=True)
        rows += [d] * len(u); cols += u.tolist(); tfs += c.tolist()
    rows, cols, tfs = np.array(rows), np.array(cols), np.array(tfs, dtype=np.float64)
    df = np.bincount(cols, minlength=n_vocab)
    idf = np.log(1.0 + (n - df + 0.5) / (df + 0.5)).astype(np.float32)
    tfc = tfs / (k1 * ((1.0 - b) + b * lens[rows] / l_avg) + tfs)
    data = (idf[cols].astype(np.float64) * tfc).astype(np.float32)
    return sp.csc_matrix((data, (rows, cols)), shape=(n, n_vocab), dtype=np.float32)


def _random_corpus(seed, n_docs, n_vocab, mean_len):
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, n_vocab + 1); p /= p.sum()                 # Zipf-like token distribution
    return [rng.choice(n_vocab, size=max(1, rng.poisson(mean_len)), p=p).tolist() for _ in range(n_docs)]


def test_oracle_known_answers_of_the_lucene_formulas():
    corpus = [[0, 1, 1], [1, 2], [0]]                                  # N=3, lens 3,2,1, l_avg=2
    idx = B.build_index(corpus, 3)
    assert idx["indptr"].tolist() == [0, 2, 4, 5]
    assert idx["indices"].tolist() == [0, 2, 0, 1, 1]
    idf0 = np.float32(math.log(1 + (3 - 2 + 0.5) / (2 + 0.5)))         # df=2
    idf2 = np.float32(math.log(1 + (3 - 1 + 0.5) / (1 + 0.5)))         # df=1
    def tfc(tf, ld): return tf / (1.2 * (0.25 + 0.75 * ld / 2.0) + tf)
    want = [np.float32(float(idf0) * tfc(1, 3)), np.float32(float(idf0) * tfc(1, 1)),
            np.float32(float(idf0) * tfc(2, 3)), np.float32(float(idf0) * tfc(1, 2)),
            np.float32(float(idf2) * tfc(1, 2))]
    assert idx["data"].tolist() == [float(x) for x in want]
    # get_scores: query order, repeats counted, unknown ids skipped, fp32 accumulation
    s = B.get_scores(idx, [1, 1, 7, 2])
    assert s.dtype == np.float32
    assert s[0] == np.float32(want[2] + want[2])
    assert s[1] == np.float32(np.float32(want[3] + want[3]) + want[4])
    assert s[2] == 0.0


@pytest.mark.parametrize("seed,n_docs,n_vocab,mean_len", [(0, 50, 30, 6), (1, 400, 200, 12), (2, 64, 500, 3)])
def test_oracle_matrix_equals_independent_scipy_construction(seed, n_docs, n_vocab, mean_len):
    corpus = _random_corpus(seed, n_docs, n_vocab, mean_len)
    idx = B.build_index(corpus, n_vocab)
    m = _scipy_index(corpus, n_vocab)
    m.sort_indices()
    assert np.array_equal(idx["indptr"], m.indptr.astype(np.int64))
    assert np.array_equal(idx["indices"], m.indices.astype(np.int32))
    assert np.array_equal(idx["data"], m.data)                          # bit-exact fp32
    # a query's score vector == sum of the matrix columns of its tokens
    rng = np.random.default_rng(seed + 100)
    for _ in range(5):
        q = rng.integers(0, n_vocab, size=6).tolist()
        want = np.zeros(n_docs, np.float32)
        for t in q:
            want += np.asarray(m[:, t].todense()).ravel()
        np.testing.assert_allclose(B.get_scores(idx, q), want, rtol=1e-6, atol=1e-7)


def test_oracle_retrieve_and_shards():
    corpus = _random_corpus(3, 300, 120, 10)
    idx = B.build_index(corpus, 120)
    queries = [[0, 5, 9], [3, 3, 40], [119]]
    rows, vals = B.retrieve(idx, queries, 20)
    assert rows.shape == (3, 20) and np.all(np.diff(vals, axis=1) <= 0)
    for q, toks in enumerate(queries):
        s = B.get_scores(idx, toks)
        assert np.array_equal(vals[q], s[rows[q]])
        assert vals[q][-1] >= np.sort(s)[-21]
    with pytest.raises(ValueError):
        B.retrieve(idx, queries, 301)
    # doc-range shards keep whole-corpus statistics: their score vectors concatenate to the full one
    parts = [B.shard(idx, lo, hi) for lo, hi in [(0, 100), (100, 101), (101, 300)]]
    for toks in queries:
        cat = np.concatenate([B.get_scores(p, toks) for p in parts])
        assert np.array_equal(cat, B.get_scores(idx, toks))


def test_tokenizer_defaults():
    from mfar_b200.data.bm25 import STOPWORDS_EN, tokenize
    assert tuple(STOPWORDS_EN) == tuple(B.STOPWORDS_EN) and len(STOPWORDS_EN) == 33
    text = "The QUICK brown-fox is at a U.S. lab; it was not 3D, x y zz"
    assert tokenize(text)[0] == B.tokenize(text) == ["quick", "brown", "fox", "lab", "3d", "zz"]
    assert tokenize([text, ""], stopwords=None)[1] == []
    # bm25s calls a callable stemmer with the WHOLE token list; a PyStemmer-like object through .stemWords
    assert tokenize("running runs", stemmer=lambda toks: [t[:3] for t in toks])[0] == ["run", "run"]

    class Stem:
        def stemWords(self, toks):
            return [t.rstrip("s") for t in toks]
    assert tokenize("runs cats", stemmer=Stem())[0] == ["run", "cat"]


def test_token_entries_layout_and_vocabulary_lookup():
    from mfar_b200.data.bm25 import token_entries, tokens_to_ids
    vocabs = [{"alpha": 0, "beta": 1, "gamma": 2}, None]
    tokens = [[["beta", "zzz", "beta", "alpha"], [], ["gamma"]],       # field 0: str tokens, OOV dropped
              [[4, 4], [7], []]]                                       # field 1: vocabulary ids
    ent = token_entries(vocabs, tokens)
    assert ent.dtype == np.int32 and ent.shape == (7, 3)
    assert ent.tolist() == [[0, 0, 1], [0, 0, 1], [0, 0, 0], [0, 1, 4], [0, 1, 4], [1, 1, 7], [2, 0, 2]]
    assert tokens_to_ids({"a": 3}, ["a", "b", "a"]) == [3, 3]
    assert token_entries([None], [[[], []]]).shape == (0, 3)
    with pytest.raises(ValueError):
        token_entries(vocabs, tokens[:1])


def test_bm25_index_classes_refuse_cpu():
    from mfar_b200.data.bm25 import DeviceBM25
    with pytest.raises(RuntimeError):
        DeviceBM25(device="cpu")
    with pytest.raises(ValueError):
        DeviceBM25(method="atire", device="cuda")


def test_bm25_c_abi_argument_checks_without_gpu():
    """Argument validation happens before any device work, so it is observable here; well-formed calls then stop at
    the architecture check (no GPU: a CUDA / arch status, never 'ok')."""
    from mfar_b200 import _native as nv
    lib = nv.lib()
    assert lib.mfar_bm25_plan_bytes(0) >= 8 and lib.mfar_bm25_plan_bytes(1000) >= 2001 * 8
    assert lib.mfar_score_topk_bm25_workspace_bytes(8, 100, 5000, 2, 64) == \
        lib.mfar_score_topk_workspace_bytes(8, 100, 5000, 2) + lib.mfar_bm25_plan_bytes(64)
    assert lib.mfar_search_host_bm25_scratch_bytes(8, 768, 768, 4, 2, 5000, 64, 100) > 8 * 5000 * 4
    fake = 0x1000                                                       # never dereferenced on these paths
    ptrs = (ctypes.c_void_p * 2)(fake, fake)
    vocab = (ctypes.c_int32 * 2)(10, 10)
    ARG, SHAPE, WS = 1, 2, 4
    # null pointer tables / negative sizes
    assert lib.mfar_bm25_scores(None, ptrs, ptrs, vocab, 2, fake, 4, 1, None, 0, 0, 100, fake, 100, 0, fake, 1 << 20,
                                None) == ARG
    assert lib.mfar_bm25_scores(ptrs, ptrs, ptrs, vocab, 2, fake, -1, 1, None, 0, 0, 100, fake, 100, 0, fake, 1 << 20,
                                None) == ARG
    assert lib.mfar_bm25_scores(ptrs, ptrs, ptrs, vocab, 2, fake, 4, 1, None, 0, 0, 100, fake, 64, 0, fake, 1 << 20,
                                None) == ARG                            # ld < n_docs
    assert lib.mfar_bm25_scores(ptrs, ptrs, ptrs, vocab, 2, fake, 4, 1, None, 0, 0, 100, fake, 100, 0, fake, 8,
                                None) == WS                             # plan scratch too small
    assert lib.mfar_bm25_scores(ptrs, ptrs, ptrs, vocab, 0, fake, 4, 1, None, 0, 0, 100, fake, 100, 0, fake, 1 << 20,
                                None) == SHAPE
    if not torch.cuda.is_available():
        rc = lib.mfar_bm25_scores(ptrs, ptrs, ptrs, vocab, 2, fake, 4, 1, None, 0, 0, 100, fake, 100, 0, fake, 1 << 20,
                                  None)
        assert rc not in (0, ARG, SHAPE, WS)
        rc = lib.mfar_bm25_build_scores(fake, fake, fake, 5, fake, fake, 10, 2.0, 1.2, 0.75, fake, None)
        assert rc not in (0, ARG, SHAPE, WS)
    assert lib.mfar_bm25_build_scores(fake, fake, fake, 5, fake, fake, 10, 0.0, 1.2, 0.75, fake, None) == ARG
    assert lib.mfar_bm25_build_scores(None, fake, fake, 5, fake, fake, 10, 2.0, 1.2, 0.75, fake, None) == ARG
