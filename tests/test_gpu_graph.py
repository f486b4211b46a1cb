"""CUDA-graph captured search (GraphedSearch) replays to exactly what the eager search returns."""
import numpy as np
import pytest
import torch

import mfar_oracle as O
from parity import assert_same_topk_up_to_ties

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(N, d, Fd, Fs, Q, bm25=False, seed=0):
    from mfar_b200.data.bm25 import DeviceBM25
    from mfar_b200.modeling.retrieval import MultiFieldRetriever, PackedCorpus
    from mfar_b200.modeling.weighting import LinearWeights
    g = torch.Generator().manual_seed(seed)
    fields = [O.round_bf16(torch.randn(N, d, generator=g)) for _ in range(Fd)]
    layer = LinearWeights(d, Fd + Fs, query_cond=True)
    with torch.no_grad():
        layer.weight.copy_(0.05 * torch.randn(d, Fd + Fs, generator=g))
    bm = None
    if bm25:
        rng = np.random.default_rng(seed)
        bm = [DeviceBM25(device=DEV).index([rng.integers(0, 200, size=rng.integers(1, 12)).tolist() for _ in range(N)],
                                           vocab=200) for _ in range(Fs)]
    r = MultiFieldRetriever(PackedCorpus.from_fields(fields, DEV), layer.to(DEV), n_sparse=Fs, sparse_indices=bm)
    return r, g


@pytest.mark.parametrize("Q", [4, 70])
def test_graphed_dense_search_equals_eager(Q):
    from mfar_b200.modeling.retrieval import GraphedSearch
    r, g = _setup(3000, 128, 3, 0, Q)
    gs = GraphedSearch(r, Q)
    assert gs.launches >= 3
    for _ in range(3):
        q = O.round_bf16(torch.randn(Q, 128, generator=g))
        s0, i0 = r.search(q.to(DEV), q.to(DEV))
        s1, i1 = gs(q, q)                                     # host tensors are copied into the static buffers
        assert torch.equal(i0, i1) and torch.equal(s0, s1)


def test_graphed_hybrid_search_dense_sparse_tensor_and_bm25_entries():
    from mfar_b200.modeling.retrieval import GraphedSearch
    Q, N = 6, 2500
    r, g = _setup(N, 64, 2, 2, Q)
    gs = GraphedSearch(r, Q, sparse="dense")
    for _ in range(2):
        q = O.round_bf16(torch.randn(Q, 64, generator=g))
        sp = torch.where(torch.rand(Q, 2, N, generator=g) < 0.9, torch.zeros(()), torch.rand(Q, 2, N, generator=g)).half()
        s0, i0 = r.search(q.to(DEV), q.to(DEV), sp.to(DEV))
        s1, i1 = gs(q, q, sparse=sp)
        assert torch.equal(i0, i1) and torch.equal(s0, s1)
    rb, g = _setup(N, 64, 2, 2, Q, bm25=True, seed=1)
    gb = GraphedSearch(rb, Q, sparse="bm25", max_entries=256)
    rng = np.random.default_rng(5)
    for n_tok in (7, 3, 0):                                   # the entry count changes between replays
        q = O.round_bf16(torch.randn(Q, 64, generator=g))
        tokens = [[rng.integers(0, 200, size=n_tok).tolist() for _ in range(Q)] for _ in range(2)]
        ent = rb.bm25.entries(tokens)
        s0, i0 = rb.search(q.to(DEV), q.to(DEV), sparse_tokens=ent)
        s1, i1 = gb(q, q, entries=ent)
        assert_same_topk_up_to_ties(s0.cpu(), i0.cpu(), s1.cpu(), i1.cpu())   # fp32 atomics: summation order
        torch.testing.assert_close(s0, s1, rtol=2e-6, atol=1e-6)
    with pytest.raises(ValueError):
        gb(q, q, entries=torch.zeros((300, 3), dtype=torch.int32, device=DEV))
