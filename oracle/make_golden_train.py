"""Generate tests/golden/train/*.npz by EXECUTING THE REFERENCE'S OWN loss classes (unmodified, imported from
/root/reference) on small seeded inputs, forward and backward, on the CPU.

Run in the build container only:   python oracle/make_golden_train.py
What is executed from the reference:
  * DecomposedContrastiveLoss.forward / compute_query_doc_field_components / compute_doc_query_scores
                                                                        mfar/modeling/losses.py:148-202
  * HybridContrastiveLoss.compute_query_doc_scores (no sparse indices; with and without BatchNorm1d)
                                                                        mfar/modeling/losses.py:327-350
  * LinearWeights                                                       mfar/modeling/weighting.py
In-batch cases use one negative per query: the reference's ``d_neg.permute(0,2,1,3).view(...)`` (losses.py:186) is
only a legal view when Neg == 1 or F == 1 and raises otherwise, so Neg > 1 has no reference behaviour to record.
"""
from __future__ import annotations

import json
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden", "train")


def main():
    ref_import.load()
    from mfar.modeling.losses import DecomposedContrastiveLoss, HybridContrastiveLoss
    from mfar.modeling.weighting import LinearWeights
    cases = [
        dict(name="train_decomposed_qc", B=6, F=3, Neg=1, E=64, T=0.05, query_cond=True, reverse=True, in_batch=True),
        dict(name="train_decomposed_static", B=5, F=4, Neg=1, E=32, T=0.1, query_cond=False, reverse=True, in_batch=True),
        dict(name="train_decomposed_simple", B=4, F=2, Neg=3, E=32, T=0.1, query_cond=True, reverse=False, in_batch=False),
        dict(name="train_decomposed_e768", B=8, F=4, Neg=1, E=768, T=0.01, query_cond=True, reverse=True, in_batch=True),
    ]
    for c in cases:
        g = torch.Generator().manual_seed(zlib.crc32(c["name"].encode()) % (2 ** 31))
        B, F, Neg, E = c["B"], c["F"], c["Neg"], c["E"]
        scale = 1.0 / np.sqrt(E)
        q = (torch.randn(B, E, generator=g) * scale).requires_grad_(True)
        d_pos = (torch.randn(B, F, E, generator=g) * scale).requires_grad_(True)
        d_neg = (torch.randn(B, F, Neg, E, generator=g) * scale).requires_grad_(True)
        layer = LinearWeights(E, F, query_cond=True) if c["query_cond"] else LinearWeights(F, 1)
        with torch.no_grad():
            layer.weight.copy_(torch.randn(layer.weight.shape, generator=g) * (0.5 if c["query_cond"] else 1.0))
        loss_mod = DecomposedContrastiveLoss(temperature=c["T"], in_batch_negative=c["in_batch"], reverse=c["reverse"],
                                             all_gather_multi_gpu=False, mixture_of_fields_layer=layer)
        loss = loss_mod(q, d_pos, d_neg)
        loss.backward()
        with torch.no_grad():
            if Neg == 1:
                pos_c, neg_c = loss_mod.compute_query_doc_field_components(q, d_pos, d_neg)
            else:                                       # the view at losses.py:186 raises; nothing to record
                pos_c = neg_c = torch.zeros(0)
            rev = loss_mod.compute_doc_query_scores(d_pos, q)
        meta = {k: v for k, v in c.items() if k != "name"}
        meta["source"] = "reference DecomposedContrastiveLoss + LinearWeights, torch CPU fp32"
        np.savez_compressed(os.path.join(GOLDEN_DIR, c["name"] + ".npz"), meta=json.dumps(meta),
                            q=q.detach().numpy(), d_pos=d_pos.detach().numpy(), d_neg=d_neg.detach().numpy(),
                            W=layer.weight.detach().numpy(), loss=loss.detach().numpy(), dq=q.grad.numpy(),
                            dd_pos=d_pos.grad.numpy(), dd_neg=d_neg.grad.numpy(), dW=layer.weight.grad.numpy(),
                            pos_components=pos_c.numpy(), neg_components=neg_c.numpy(), rev_scores=rev.numpy())
        print(c["name"], float(loss.detach()))

    # Hybrid scorer with BatchNorm1d over fields (training-mode batch statistics), no sparse indices
    for use_bn in (False, True):
        name = "train_hybrid_bn" if use_bn else "train_hybrid_nobn"
        g = torch.Generator().manual_seed(7 + int(use_bn))
        B, F, Neg, E, T = 6, 3, 1, 64, 0.05
        scale = 1.0 / np.sqrt(E)
        q = (torch.randn(B, E, generator=g) * scale).requires_grad_(True)
        d_pos = (torch.randn(B, F, E, generator=g) * scale).requires_grad_(True)
        d_neg = (torch.randn(B, F, Neg, E, generator=g) * scale).requires_grad_(True)
        layer = LinearWeights(E, F, query_cond=True)
        with torch.no_grad():
            layer.weight.copy_(torch.randn(E, F, generator=g) * 0.5)
        mod = HybridContrastiveLoss(temperature=T, all_gather_multi_gpu=False, mixture_of_fields_layer=layer,
                                    sparse_indices_dict={}, num_fields=F, use_batchnorm=use_bn)
        torch.Tensor.cuda = lambda self, *a, **k: self                 # losses.py:323,325 call .cuda(); CPU run
        sp, sn = mod.compute_query_doc_scores(q, ["q"] * B, d_pos, ["d"] * B, d_neg, ["n"] * B * Neg, list(range(B)), {})
        loss = mod.sliced_nll(torch.cat([sp, sn], dim=1), B, 0)
        loss.backward()
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"),
                            meta=json.dumps(dict(B=B, F=F, Neg=Neg, E=E, T=T, use_bn=use_bn, query_cond=True,
                                                 source="reference HybridContrastiveLoss.compute_query_doc_scores")),
                            q=q.detach().numpy(), d_pos=d_pos.detach().numpy(), d_neg=d_neg.detach().numpy(),
                            W=layer.weight.detach().numpy(), scores_pos=sp.detach().numpy(),
                            scores_neg=sn.detach().numpy(), loss=loss.detach().numpy(), dq=q.grad.numpy(),
                            dd_pos=d_pos.grad.numpy(), dd_neg=d_neg.grad.numpy(), dW=layer.weight.grad.numpy())
        print(name, float(loss.detach()))


if __name__ == "__main__":
    main()
