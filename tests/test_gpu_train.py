"""CUDA training-time scorer (csrc/train.cu through the C ABI) vs the reference-generated golden vectors and the
torch CPU oracle (oracle/train_oracle.py): forward values, loss, and every gradient (q, d_pos, d_neg, W).
fp32 throughout; tolerance 1e-4 of the tensor's max magnitude (summation order differs; dq uses fp32 atomics)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import train_oracle as T

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "train")
DECOMPOSED = sorted(glob.glob(os.path.join(GOLDEN, "train_decomposed_*.npz")))
HYBRID = sorted(glob.glob(os.path.join(GOLDEN, "train_hybrid_*.npz")))
RTOL = 1e-4


def _mods():
    from mfar_b200.modeling import losses as L
    from mfar_b200.modeling.weighting import LinearWeights
    return L, LinearWeights


def load(path):
    z = np.load(path, allow_pickle=False)
    return z, json.loads(str(z["meta"]))


def dleaf(a):
    return torch.from_numpy(np.array(a)).to(DEV).requires_grad_(True)


def close(a, b, what, rtol=RTOL):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    a, b = a.astype(np.float64), b.astype(np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    scale = max(np.abs(b).max(), 1e-30)
    assert np.abs(a - b).max() <= rtol * scale, f"{what}: max abs diff {np.abs(a - b).max()} vs scale {scale}"


def make_layer(W, query_cond):
    _, LinearWeights = _mods()
    W = W.detach().clone() if torch.is_tensor(W) else torch.from_numpy(np.array(W))
    layer = LinearWeights(W.shape[0], W.shape[1], query_cond=query_cond)
    with torch.no_grad():
        layer.weight.copy_(W)
    return layer.to(DEV)


@pytest.mark.parametrize("path", DECOMPOSED, ids=lambda p: os.path.basename(p)[:-4])
def test_decomposed_loss_vs_reference_golden(path):
    L, _ = _mods()
    z, m = load(path)
    q, dp, dn = dleaf(z["q"]), dleaf(z["d_pos"]), dleaf(z["d_neg"])
    layer = make_layer(z["W"], m["query_cond"])
    mod = L.DecomposedContrastiveLoss(temperature=m["T"], in_batch_negative=m["in_batch"], reverse=m["reverse"],
                                      all_gather_multi_gpu=False, mixture_of_fields_layer=layer)
    loss = mod(q, dp, dn)
    loss.backward()
    close(loss, z["loss"], "loss")
    close(q.grad, z["dq"], "dq"); close(dp.grad, z["dd_pos"], "dd_pos")
    close(dn.grad, z["dd_neg"], "dd_neg"); close(layer.weight.grad, z["dW"], "dW")
    if z["pos_components"].size:
        pc, nc = mod.compute_query_doc_field_components(q.detach(), dp.detach(), dn.detach())
        close(pc, z["pos_components"], "pos components", 2e-6); close(nc, z["neg_components"], "neg components", 2e-6)
        close(mod.compute_doc_query_scores(dp.detach(), q.detach()), z["rev_scores"], "reverse scores", 1e-5)


@pytest.mark.parametrize("path", HYBRID, ids=lambda p: os.path.basename(p)[:-4])
def test_hybrid_scores_vs_reference_golden(path):
    L, _ = _mods()
    z, m = load(path)
    B, F, Neg = m["B"], m["F"], m["Neg"]
    q, dp, dn = dleaf(z["q"]), dleaf(z["d_pos"]), dleaf(z["d_neg"])
    layer = make_layer(z["W"], True)
    mod = L.HybridContrastiveLoss(temperature=m["T"], all_gather_multi_gpu=False, mixture_of_fields_layer=layer,
                                  sparse_indices_dict={}, num_fields=F, use_batchnorm=m["use_bn"]).to(DEV)
    sp, sn = mod.compute_query_doc_scores(q, ["q"] * B, dp, ["d"] * B, dn, ["n"] * B * Neg, list(range(B)), {})
    loss = mod.sliced_nll(torch.cat([sp, sn], dim=1), B, 0)
    loss.backward()
    close(sp, z["scores_pos"], "scores_pos", 2e-5); close(sn, z["scores_neg"], "scores_neg", 2e-5)
    close(loss, z["loss"], "loss")
    close(q.grad, z["dq"], "dq", 3e-4); close(dp.grad, z["dd_pos"], "dd_pos", 3e-4)
    close(dn.grad, z["dd_neg"], "dd_neg", 3e-4); close(layer.weight.grad, z["dW"], "dW", 3e-4)


SHAPES = [  # B, P, F, Neg, E, query_cond   (P = docs gathered from all devices; B = this device's queries)
    (7, 7, 3, 2, 64, True),          # Neg > 1: the doc order p*Neg + s the reference's view intends
    (24, 48, 8, 1, 768, True),       # training defaults, 2 devices' worth of docs
    (40, 40, 5, 3, 100, True),       # two query tiles (B > 32), E with a partial 128-chunk
    (33, 70, 22, 1, 768, False),     # PRIME field count, static mixture
    (3, 5, 1, 4, 1024, True),        # single field, max E
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "B%d_P%d_F%d_Neg%d_E%d_%s" % (*s[:5], "qc" if s[5] else "static"))
def test_components_and_mixture_forward_backward_vs_oracle(shape):
    L, _ = _mods()
    B, P, F, Neg, E, qc = shape
    g = torch.Generator().manual_seed(sum(shape[:5]))
    sc = 1.0 / np.sqrt(E)
    q0, dp0 = torch.randn(B, E, generator=g) * sc, torch.randn(P, F, E, generator=g) * sc
    dn0 = torch.randn(P, F, Neg, E, generator=g) * sc
    W0 = torch.randn(E, F, generator=g) * 0.5 if qc else torch.randn(F, 1, generator=g)
    temp = 0.05
    # oracle
    q, dp, dn, W = (t.clone().requires_grad_(True) for t in (q0, dp0, dn0, W0))
    pc, nc = T.field_components(q, dp, dn, temp)
    scores = torch.cat([T.mixture(pc, q, W, qc), T.mixture(nc, q, W, qc)], dim=1)
    go = torch.randn(scores.shape, generator=g)
    (scores * go).sum().backward()
    # CUDA
    cq, cdp, cdn = (t.clone().to(DEV).requires_grad_(True) for t in (q0, dp0, dn0))
    layer = make_layer(W0, qc)
    mod = L.DecomposedContrastiveLoss(temperature=temp, all_gather_multi_gpu=False, mixture_of_fields_layer=layer)
    cpc, cnc = mod.compute_query_doc_field_components(cq, cdp, cdn)
    close(cpc, pc, "pos components", 2e-6); close(cnc, nc, "neg components", 2e-6)
    sp, sn = mod.compute_query_doc_scores(cq, cdp, cdn)
    cscores = torch.cat([sp, sn], dim=1)
    close(cscores, scores, "scores", 1e-5)
    (cscores * go.to(DEV)).sum().backward()
    close(cq.grad, q.grad, "dq"); close(cdp.grad, dp.grad, "dd_pos"); close(cdn.grad, dn.grad, "dd_neg")
    close(layer.weight.grad, W.grad, "dW")


def test_gradients_only_where_requested_and_cpu_tensors_refused():
    L, LinearWeights = _mods()
    g = torch.Generator().manual_seed(0)
    q = torch.randn(4, 64, generator=g).to(DEV)
    d = torch.randn(9, 2, 64, generator=g).to(DEV).requires_grad_(True)
    c = L.field_components(q, d, 0.5)
    assert c.shape == (4, 9, 2) and c.requires_grad
    c.sum().backward()
    ref = torch.einsum("be->e", q.cpu()) / 0.5
    close(d.grad[3, 1], ref, "ddocs row")
    with pytest.raises(RuntimeError):
        L.field_components(q.cpu(), d, 0.5)
    with pytest.raises(RuntimeError):
        L.field_components(q[:, :60].contiguous(), d, 0.5)          # E mismatch
    # inference path of LinearWeights unchanged under no_grad
    layer = LinearWeights(64, 2, query_cond=True).to(DEV)
    with torch.no_grad():
        out = layer(c.detach(), q)
    assert not out.requires_grad and out.shape == (4, 9)
