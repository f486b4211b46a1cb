"""CPU oracle for the training-time scorer (torch CPU fp32 + autograd).

TEST INFRASTRUCTURE ONLY - imported by ``tests/`` (and ``oracle/make_golden_train.py``), never by the product.

Fresh restatement of the reference's multi-field contrastive scoring; gradients come from torch autograd over the
restated forward, which is what the reference itself relies on.  Pinned by ``tests/golden/train/*.npz``: outputs of
the reference's own ``DecomposedContrastiveLoss`` / ``HybridContrastiveLoss`` / ``LinearWeights`` classes (imported
unmodified from /root/reference by ``oracle/make_golden_train.py``) - loss values and gradients on seeded inputs.

  field_components     mfar/modeling/losses.py:176-188
  mixture              mfar/modeling/weighting.py:17-29
  sliced_nll           mfar/modeling/losses.py:59-66
  decomposed_loss      mfar/modeling/losses.py:68-88 (in-batch negatives, optional reverse term 83-87),
                       90-109 (simple loss), 190-202
  hybrid_scores        mfar/modeling/losses.py:327-360 (sparse columns concatenated, optional BatchNorm1d)
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def field_components(q: torch.Tensor, d_pos: torch.Tensor, d_neg: torch.Tensor, temperature: float
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """q [B,E], d_pos [P,F,E], d_neg [P,F,Neg,E] -> ([B,P,F], [B,P*Neg,F]) (losses.py:176-188)."""
    pos = torch.einsum("pfe,be->bpf", d_pos, q) / temperature
    P, F, Neg, E = d_neg.shape
    neg_docs = d_neg.permute(0, 2, 1, 3).reshape(P * Neg, F, E)          # doc order p*Neg + s (losses.py:186)
    neg = torch.einsum("nfe,be->bnf", neg_docs, q) / temperature
    return pos, neg


def mixture(x: torch.Tensor, q: Optional[torch.Tensor], W: torch.Tensor, query_cond: bool) -> torch.Tensor:
    """LinearWeights.forward (weighting.py:17-29): x [B,S,F] -> [B,S]."""
    logits = q @ W if query_cond else W.transpose(1, 0)
    return torch.sum(torch.softmax(logits, dim=1).unsqueeze(1) * x, dim=-1)


def sliced_nll(scores: torch.Tensor, batch_size: int, gpu_id: int = 0) -> torch.Tensor:
    lp = torch.log_softmax(scores, dim=1)[:, batch_size * gpu_id: batch_size * (gpu_id + 1)]
    return -torch.mean(torch.diag(lp))


def decomposed_loss(q, d_pos, d_neg, W, temperature: float, query_cond: bool, reverse: bool = True,
                    in_batch_negative: bool = True) -> torch.Tensor:
    """Single-process DecomposedContrastiveLoss.forward (losses.py:161-174)."""
    if in_batch_negative:
        pos_c, neg_c = field_components(q, d_pos, d_neg, temperature)
        scores = torch.cat([mixture(pos_c, q, W, query_cond), mixture(neg_c, q, W, query_cond)], dim=1)
        nll = sliced_nll(scores, q.size(0))
        if reverse:                                                      # losses.py:199-202
            rev_c = torch.matmul(d_pos, q.t().unsqueeze(0)) / temperature            # [P,F,B]
            rev = mixture(rev_c.permute(2, 0, 1), q, W, query_cond).t()
            nll = nll + sliced_nll(rev, q.size(0))
        return nll
    pos = torch.matmul(d_pos, q.unsqueeze(2)).squeeze(2).unsqueeze(1) / temperature               # [B,1,F]
    neg = torch.sum(q.view(q.size(0), 1, 1, q.size(1)) * d_neg.permute(0, 2, 1, 3), dim=-1) / temperature
    scores = torch.cat([mixture(pos, q, W, query_cond), mixture(neg, q, W, query_cond)], dim=1)
    return -torch.mean(torch.log_softmax(scores, dim=1)[:, 0])


def hybrid_scores(q, d_pos, d_neg, sparse_pos, sparse_neg, W, temperature: float, query_cond: bool,
                  bn: Optional[torch.nn.Module] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """HybridContrastiveLoss.compute_query_doc_scores (losses.py:327-350) with the sparse columns given as
    tensors [B,P,Fs] / [B,P*Neg,Fs]."""
    dp, dn = field_components(q, d_pos, d_neg, temperature)
    allc = torch.cat([torch.cat([dp, sparse_pos], dim=-1), torch.cat([dn, sparse_neg], dim=-1)], dim=1)
    if bn is not None:
        allc = bn(allc.permute(0, 2, 1)).permute(0, 2, 1)
    comb = mixture(allc, q, W, query_cond)
    return comb[:, :dp.size(1)], comb[:, dp.size(1):]
