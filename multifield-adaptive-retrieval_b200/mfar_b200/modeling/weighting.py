"""LinearWeights - the query-conditioned field mixture (mfar/modeling/weighting.py:3-29).

Same constructor, parameter name/shape and forward signature as the reference module:
  ``LinearWeights(emb_size, num_fields, query_cond)``, ``.weight`` [emb_size, num_fields]
  (init ones, weighting.py:14), ``forward(x[B,S,F] or [S,F], q[B,E] or None) -> [B,S]``.
When ``query_cond`` is False the reference constructs it as ``LinearWeights(num_fields, 1)``
(contrastive.py:285), i.e. weight is [F,1] and the softmax runs over its transpose.

Inference only: forward runs the CUDA kernels behind ``mfar_mixture_weights`` /
``mfar_mixture_apply`` (fp32) and is not differentiable.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import _native as nv


class LinearWeights(torch.nn.Module):

    def __init__(self, emb_size: int, num_fields: int, query_cond: bool = False):
        super().__init__()
        self.query_cond = query_cond
        self.weight = torch.nn.Parameter(torch.ones(emb_size, num_fields), requires_grad=False)

    @property
    def num_fields(self) -> int:
        return self.weight.shape[1] if self.query_cond else self.weight.shape[0]

    @torch.no_grad()
    def field_weights(self, q: Optional[torch.Tensor], mask: Optional[torch.Tensor] = None,
                      batch: Optional[int] = None) -> torch.Tensor:
        """softmax field weights [Q,F] (times the 0/1 field mask, un-renormalised)."""
        W = self.weight.detach()
        nv.require_device(W, "LinearWeights.weight")
        W = W.contiguous().float()
        F = self.num_fields
        if self.query_cond:
            if q is None:
                raise ValueError("query_cond=True needs the query embedding")
            nv.require_device(q, "q")
            qe = q.detach().contiguous().float()
            Q, E = qe.shape
            if E != W.shape[0]:
                raise RuntimeError(f"q is [*,{E}] but weight is [{W.shape[0]},{W.shape[1]}]")
        else:
            qe, Q, E = None, (batch or 1), 0
        m = None
        if mask is not None:
            m = mask.detach().to(device=W.device, dtype=torch.float32).reshape(-1).contiguous()
            if m.numel() != F:
                raise RuntimeError(f"mask has {m.numel()} entries for {F} fields")
        out = torch.empty((Q, F), dtype=torch.float32, device=W.device)
        nv.check(nv.lib().mfar_mixture_weights(nv.ptr(qe), nv.ptr(W), nv.ptr(m), Q, E, F, int(self.query_cond),
                                               nv.ptr(out), nv.stream()), "mixture_weights")
        return out

    @torch.no_grad()
    def forward(self, x: torch.Tensor, q: Optional[torch.Tensor]) -> torch.Tensor:
        nv.require_device(x, "x")
        squeeze = x.dim() == 2                       # trec_eval_step passes [S,F] (contrastive.py:694)
        x3 = (x.unsqueeze(0) if squeeze else x).contiguous().float()
        B, S, F = x3.shape
        if F != self.num_fields:
            raise RuntimeError(f"x has {F} fields, layer has {self.num_fields}")
        w = self.field_weights(q)                    # [Q,F] or [1,F]
        if squeeze and w.shape[0] > 1:               # [S,F] against B queries broadcasts to [B,S] (weighting.py:29)
            B = w.shape[0]
            x3 = x3.expand(B, S, F).contiguous()
        if w.shape[0] not in (1, B):
            raise RuntimeError("batch size of x and q must match (weighting.py:21-23)")
        out = torch.empty((B, S), dtype=torch.float32, device=x3.device)
        nv.check(nv.lib().mfar_mixture_apply(nv.ptr(x3), nv.ptr(w), B, S, F, w.shape[0], nv.ptr(out), nv.stream()),
                 "mixture_apply")
        return out
