"""Training-scorer oracle (oracle/train_oracle.py, torch CPU fp32 + autograd) vs golden vectors produced by the
reference's own DecomposedContrastiveLoss / HybridContrastiveLoss / LinearWeights (oracle/make_golden_train.py)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import train_oracle as T

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "train")
DECOMPOSED = sorted(glob.glob(os.path.join(GOLDEN, "train_decomposed_*.npz")))
HYBRID = sorted(glob.glob(os.path.join(GOLDEN, "train_hybrid_*.npz")))


def load(path):
    z = np.load(path, allow_pickle=False)
    return z, json.loads(str(z["meta"]))


def leaf(a):
    return torch.from_numpy(np.array(a)).requires_grad_(True)


def close(a, b, what, rtol=2e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    assert np.abs(a - b).max() <= rtol * scale, f"{what}: max abs diff {np.abs(a - b).max()} vs scale {scale}"


def test_fixtures_present():
    assert len(DECOMPOSED) == 4 and len(HYBRID) == 2


@pytest.mark.parametrize("path", DECOMPOSED, ids=lambda p: os.path.basename(p)[:-4])
def test_decomposed_loss_and_gradients_match_reference(path):
    z, m = load(path)
    q, dp, dn, W = leaf(z["q"]), leaf(z["d_pos"]), leaf(z["d_neg"]), leaf(z["W"])
    loss = T.decomposed_loss(q, dp, dn, W, m["T"], m["query_cond"], m["reverse"], m["in_batch"])
    loss.backward()
    close(loss.detach(), z["loss"], "loss")
    close(q.grad, z["dq"], "dq"); close(dp.grad, z["dd_pos"], "dd_pos")
    close(dn.grad, z["dd_neg"], "dd_neg"); close(W.grad, z["dW"], "dW")
    if z["pos_components"].size:
        pc, nc = T.field_components(q.detach(), dp.detach(), dn.detach(), m["T"])
        close(pc, z["pos_components"], "pos components"); close(nc, z["neg_components"], "neg components")


@pytest.mark.parametrize("path", HYBRID, ids=lambda p: os.path.basename(p)[:-4])
def test_hybrid_scores_match_reference(path):
    z, m = load(path)
    q, dp, dn, W = leaf(z["q"]), leaf(z["d_pos"]), leaf(z["d_neg"]), leaf(z["W"])
    B, F, Neg = m["B"], m["F"], m["Neg"]
    bn = torch.nn.BatchNorm1d(F, track_running_stats=True) if m["use_bn"] else None
    sp, sn = T.hybrid_scores(q, dp, dn, torch.empty(B, B, 0), torch.empty(B, B * Neg, 0), W, m["T"], True, bn)
    loss = T.sliced_nll(torch.cat([sp, sn], dim=1), B)
    loss.backward()
    close(sp.detach(), z["scores_pos"], "scores_pos"); close(sn.detach(), z["scores_neg"], "scores_neg")
    close(loss.detach(), z["loss"], "loss")
    close(q.grad, z["dq"], "dq", 1e-4); close(dp.grad, z["dd_pos"], "dd_pos", 1e-4)
    close(dn.grad, z["dd_neg"], "dd_neg", 1e-4); close(W.grad, z["dW"], "dW", 1e-4)
