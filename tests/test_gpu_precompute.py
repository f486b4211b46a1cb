"""Device producer of the precomputed-BM25 score files (csrc/sparse_coo.cu, mfar_sparse_coo_count/write) against
the goldens the reference's own ``precompute_score_for_field`` produced and against the oracle; bit-exact (integer
keys, fp16 values).  Every call goes through the C ABI."""
import glob
import os

import numpy as np
import pytest
import torch

import bm25_oracle as B
import precompute_oracle as PO

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "precompute", "*.npz")))


def _coo(scores, n_docs, safe=None, qids=None, base=0, total=None, dtype=torch.float16):
    from mfar_b200.data.bm25 import rows_to_coo, safe_docs_bitmap
    bits = None
    if safe is not None:
        bits = torch.from_numpy(safe_docs_bitmap(safe, total if total is not None else base + n_docs).view(np.int32)).to(DEV)
    q = None if qids is None else torch.as_tensor(np.asarray(qids), dtype=torch.int32)
    k, v = rows_to_coo(scores, n_docs, bits, q, base, dtype)
    return k.cpu().numpy(), v.cpu().numpy()


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_rows_to_coo_vs_reference_golden(path):
    z = np.load(path, allow_pickle=False)
    scores = torch.from_numpy(z["scores"]).to(DEV)
    keys, vals = _coo(scores, scores.shape[1], set(z["safe"].tolist()), z["qids"])
    assert keys.dtype == np.int32 and vals.dtype == np.float16
    np.testing.assert_array_equal(keys, z["ref_keys"])
    np.testing.assert_array_equal(vals.view(np.uint16), z["ref_vals"].view(np.uint16))


@pytest.mark.parametrize("N,Q,pad,density", [(1, 1, 0, 1.0), (4096, 3, 0, 0.5), (4097, 2, 7, 0.01), (70001, 9, 3, 0.2),
                                             (300000, 5, 0, 0.001)])
def test_rows_to_coo_vs_oracle_padded_rows_shards_and_fp32(N, Q, pad, density):
    rng = np.random.RandomState(N + Q)
    dense = np.where(rng.rand(Q, N) < density, rng.gamma(2.0, 2.0, (Q, N)), 0.0).astype(np.float32)
    buf = torch.full((Q, N + pad), 7.0, device=DEV)                 # padding columns must never be emitted
    buf[:, :N] = torch.from_numpy(dense).to(DEV)
    view = buf[:, :N]
    safe = set(rng.choice(N, size=max(1, N // 3), replace=False).tolist())
    qids = 5 + 3 * np.arange(Q)
    rows = {int(q): dense[i] for i, q in enumerate(qids)}
    want_k, want_v = PO.precompute_score_for_field(rows, safe)
    got_k, got_v = _coo(view, N, safe, qids)
    np.testing.assert_array_equal(got_k, want_k)
    np.testing.assert_array_equal(got_v.view(np.uint16), want_v.view(np.uint16))
    # no safe set = every doc; no qids = row numbers; fp32 values are the scores themselves
    all_k, all_v = _coo(view, N, None, None, dtype=torch.float32)
    qq, dd = np.nonzero(dense)
    np.testing.assert_array_equal(all_k, np.stack([qq, dd], axis=1).astype(np.int32))
    np.testing.assert_array_equal(all_v, dense[qq, dd])
    # doc-range shards with global ids: per query the shard outputs concatenate to the unsharded output
    if N >= 2:
        cut = N // 2 + 1
        parts = [_coo(buf[:, lo:hi], hi - lo, safe, qids, base=lo, total=N) for lo, hi in ((0, cut), (cut, N))]
        merged_k = np.concatenate([p[0] for p in parts])
        merged_v = np.concatenate([p[1] for p in parts])
        order = np.lexsort((merged_k[:, 1], np.searchsorted(qids, merged_k[:, 0])))
        np.testing.assert_array_equal(merged_k[order], want_k)
        np.testing.assert_array_equal(merged_v[order].view(np.uint16), want_v.view(np.uint16))


def _text_corpus(seed, n_docs, n_words):
    rng = np.random.default_rng(seed)
    vocab = [f"w{i:03d}" for i in range(n_words)]
    docs = {str(i): " ".join(rng.choice(vocab, size=rng.integers(1, 12))) for i in range(n_docs)}
    return vocab, docs


def test_get_scores_sparse_batch_equals_the_per_query_dicts():
    from mfar_b200.data.index import BM25sSparseIndex
    vocab, docs = _text_corpus(3, 5000, 300)
    idx = BM25sSparseIndex.create(docs, device=DEV)
    rng = np.random.default_rng(4)
    queries = [" ".join(rng.choice(vocab, size=4)) for _ in range(7)] + ["zzz unknown words only"]
    safe = set(rng.choice(5000, size=2000, replace=False).tolist())
    idx.set_safe_docs(safe)
    keys, vals = idx.get_scores_sparse_batch(queries, query_ids=range(100, 108))
    off = 0
    for qi, q in enumerate(queries):
        d = idx.get_scores_sparse(q)                                  # index.py:78-84 mirror, one query
        assert all(doc in safe for doc in d)
        n = len(d)
        np.testing.assert_array_equal(keys[off:off + n], np.array([(100 + qi, doc) for doc in d], np.int32).reshape(-1, 2))
        # values: the single-query and the batched scatter may add a doc's token contributions in different orders
        np.testing.assert_allclose(vals[off:off + n].astype(np.float32),
                                   np.array([np.float16(s) for s in d.values()], np.float32), rtol=2e-3, atol=1e-7)
        off += n
    assert off == len(keys) and len(idx.get_scores_sparse(queries[-1])) == 0
    idx.set_safe_docs({1, 2, 3})                                      # re-binding the safe set rebuilds the bitmap
    k2, _ = idx.get_scores_sparse_batch(queries[:2])
    assert set(k2[:, 1].tolist()) <= {1, 2, 3} and set(k2[:, 0].tolist()) <= {0, 1}


def test_precompute_files_feed_the_coo_search_path(tmp_path):
    """precompute_score_for_field (device) -> the reference's files -> PrecomputedSparseScores -> search(sparse_coo=)
    == search over the dense score rows restricted to the safe docs (score_batch_with_cache semantics: absent -> 0);
    and the pairs agree with the BM25 oracle's score vectors."""
    from mfar_b200.commands.precompute_bm25s_scores import precompute_score_for_field
    from mfar_b200.data.bm25 import tokenize
    from mfar_b200.data.index import BM25sSparseIndex
    from mfar_b200.data.typedef import Field, FieldType
    from mfar_b200.modeling.retrieval import MultiFieldRetriever, PackedCorpus
    from mfar_b200.modeling.util import PrecomputedSparseScores
    from mfar_b200.modeling.weighting import LinearWeights
    import mfar_oracle as O
    N, d, Fs, Q, k = 3000, 64, 2, 6, 20
    rng = np.random.default_rng(11)
    finfo, indices, texts = {}, {}, {}
    vocab = None
    for j in range(Fs):
        vocab, docs = _text_corpus(20 + j, N, 200)
        fk = f"f{j}_sparse"
        finfo[fk] = Field(fk, f"f{j}", FieldType.SPARSE)
        indices[fk] = BM25sSparseIndex.create(docs, device=DEV)
        texts[fk] = docs
    train_queries = {900 + 11 * i: " ".join(rng.choice(vocab, size=5)) for i in range(Q)}
    safe = set(rng.choice(N, size=N // 2, replace=False).tolist())
    for fk, idx in indices.items():
        keys, vals = precompute_score_for_field(idx, safe, train_queries, str(tmp_path), fk, batch_size=4)
        assert np.array_equal(np.load(tmp_path / f"{fk}_keys_bm25.npy"), keys) and keys.dtype == np.int32
        assert np.load(tmp_path / f"{fk}_vals_bm25.npy").dtype == np.float16
        # against the BM25 oracle (parity unpinned for BM25 itself; sums reorder under atomics -> 1 f16 ulp)
        toks = tokenize(list(texts[fk].values()))
        vd = idx.index.vocab_dict
        oidx = B.build_index([[vd[t] for t in doc] for doc in toks], len(vd))
        rows = {qid: B.get_scores(oidx, [vd[t] for t in tokenize(q)[0] if t in vd]) for qid, q in train_queries.items()}
        want_k, want_v = PO.precompute_score_for_field(rows, safe)
        np.testing.assert_array_equal(keys, want_k)
        np.testing.assert_allclose(vals.astype(np.float32), want_v.astype(np.float32), rtol=2e-3, atol=1e-6)
    store = PrecomputedSparseScores.load(str(tmp_path), finfo)
    qids = list(train_queries.keys())
    g = torch.Generator().manual_seed(5)
    fields = [O.round_bf16(torch.randn(N, d, generator=g))]
    q = O.round_bf16(torch.randn(Q, d, generator=g))
    layer = LinearWeights(d, 1 + Fs, query_cond=True)
    with torch.no_grad():
        layer.weight.copy_(0.05 * torch.randn(d, 1 + Fs, generator=g))
    r = MultiFieldRetriever(PackedCorpus.from_fields(fields, DEV), layer.to(DEV), n_sparse=Fs, top_k=k)
    s_coo, i_coo = r.search(q.to(DEV), q.to(DEV), sparse_coo=store.batch(qids, DEV))
    dense = torch.zeros(Q, Fs, N)
    for j, fk in enumerate(finfo):
        kk, vv = np.load(tmp_path / f"{fk}_keys_bm25.npy"), np.load(tmp_path / f"{fk}_vals_bm25.npy")
        dense[torch.from_numpy(np.searchsorted(qids, kk[:, 0])).long(), j, torch.from_numpy(kk[:, 1]).long()] = \
            torch.from_numpy(vv.astype(np.float32))
    s_den, i_den = r.search(q.to(DEV), q.to(DEV), dense.to(DEV))
    torch.testing.assert_close(s_coo, s_den, rtol=1e-6, atol=1e-5)
    from parity import assert_same_topk_up_to_ties
    assert_same_topk_up_to_ties(s_coo.cpu(), i_coo.cpu(), s_den.cpu(), i_den.cpu())
