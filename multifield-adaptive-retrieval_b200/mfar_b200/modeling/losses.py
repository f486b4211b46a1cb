"""Training-time scorer and contrastive losses (mfar/modeling/losses.py).

The scoring inside the losses - per-field query.doc components divided by the temperature and the query-conditioned
field mixture, forward and backward - runs on CUDA kernels behind the C ABI (``mfar_field_components_fwd/bwd``,
``mfar_mixture_weights/apply/bwd``); the loss tail (log-softmax over a [B, N] score matrix, diagonal, mean) and the
cross-GPU gather are the same few torch / torch.distributed calls the reference makes.

``field_components``            losses.py:176-188   [B,E] x [N,F,E] (or [P,F,Neg,E]) -> [B,N,F], differentiable
``DecomposedContrastiveLoss``   losses.py:148-202   same constructor / forward / helper names
``HybridContrastiveLoss``       losses.py:204-360   dense + sparse columns, optional BatchNorm1d(F), mixture
"""
from __future__ import annotations

import pickle
from typing import Dict, Optional, Tuple

import torch

from .. import _native as nv

try:                                                    # losses.py:7-10
    import torch.distributed.nn.functional as dist_F
except Exception:                                       # pragma: no cover
    dist_F = None


def _layout(docs: torch.Tensor) -> Tuple[int, int, int, int, int, int]:
    """(N, F, inner, stride_p, stride_f, stride_s) of a contiguous [N,F,E] or [P,F,Neg,E] tensor."""
    if docs.dim() == 3:
        P, F, E = docs.shape
        return P, F, 1, F * E, E, 0
    if docs.dim() == 4:
        P, F, Neg, E = docs.shape
        return P * Neg, F, Neg, F * Neg * E, Neg * E, E
    raise ValueError(f"docs must be [N,F,E] or [P,F,Neg,E], got {tuple(docs.shape)}")


class FieldComponentsFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, q: torch.Tensor, docs: torch.Tensor, temperature: float):
        nv.require_device(q, "q")
        nv.require_device(docs, "docs")
        qc = q.detach().contiguous().float()
        dc = docs.detach().contiguous().float()
        B, E = qc.shape
        if dc.shape[-1] != E:
            raise RuntimeError(f"q is [*,{E}] but docs are [*,{dc.shape[-1]}]")
        N, F, inner, sp, sf, ss = _layout(dc)
        comp = torch.empty((B, N, F), dtype=torch.float32, device=qc.device)
        nv.check(nv.lib().mfar_field_components_fwd(nv.ptr(qc), B, E, nv.ptr(dc), N, F, inner, sp, sf, ss,
                                                    float(temperature), nv.ptr(comp), nv.stream()),
                 "field_components_fwd")
        ctx.save_for_backward(qc, dc)
        ctx.temperature = float(temperature)
        ctx.docs_dtype, ctx.q_dtype = docs.dtype, q.dtype
        return comp

    @staticmethod
    def backward(ctx, dcomp: torch.Tensor):
        qc, dc = ctx.saved_tensors
        B, E = qc.shape
        N, F, inner, sp, sf, ss = _layout(dc)
        need_q, need_d, _ = ctx.needs_input_grad
        dq = torch.empty_like(qc) if need_q else None
        dd = torch.empty_like(dc) if need_d else None
        if dq is None and dd is None:
            return None, None, None
        g = dcomp.contiguous().float()
        nv.check(nv.lib().mfar_field_components_bwd(nv.ptr(qc), B, E, nv.ptr(dc), N, F, inner, sp, sf, ss,
                                                    ctx.temperature, nv.ptr(g), nv.ptr(dq), nv.ptr(dd), nv.stream()),
                 "field_components_bwd")
        return (None if dq is None else dq.to(ctx.q_dtype)), (None if dd is None else dd.to(ctx.docs_dtype)), None


def field_components(q: torch.Tensor, docs: torch.Tensor, temperature: float = 1.0) -> torch.Tensor:
    """comp[b, n, f] = <q[b], docs[n, f]> / temperature  (losses.py:182-187).

    docs [N,F,E], or the reference's negatives layout [P,F,Neg,E] whose docs are ordered p*Neg + s like
    ``d_neg.permute(0,2,1,3).view(1, P*Neg, F, E)`` (losses.py:186) - no permuted copy is made."""
    return FieldComponentsFunction.apply(q, docs, temperature)


class BaseContrastiveLoss(torch.nn.Module):
    """losses.py:12-109 (the parts the multi-field losses use)."""

    def __init__(self, temperature: float = 0.01, in_batch_negative: bool = True, reverse: bool = True,
                 all_gather_multi_gpu: bool = True, multi_fields: bool = False):
        super().__init__()
        self.temperature = temperature
        self.in_batch_negative = in_batch_negative
        self.reverse = reverse
        self.all_gather_multi_gpu = all_gather_multi_gpu
        self.multi_fields = multi_fields

    def gather_all_embeddings(self, q, d_pos, d_neg, use_multi_gpu):
        if use_multi_gpu:                                               # losses.py:44-52
            q = torch.cat(dist_F.all_gather(q), dim=0)
            d_pos = torch.cat(dist_F.all_gather(d_pos), dim=0)
            d_neg = torch.cat(dist_F.all_gather(d_neg), dim=0)
        return q, d_pos, d_neg

    def distributed_reduce_mean_nll(self, nll, use_multi_gpu):
        if use_multi_gpu:                                               # losses.py:54-57
            nll = dist_F.all_reduce(nll) / torch.distributed.get_world_size()
        return nll

    def sliced_nll(self, scores, batch_size, gpu_id):
        log_probs = torch.log_softmax(scores, dim=1)                    # losses.py:59-66
        sliced = log_probs[:, (batch_size * gpu_id): (batch_size * (gpu_id + 1))]
        return -torch.mean(torch.diag(sliced))

    def in_batch_negative_loss(self, q, d_pos, d_neg, use_multi_gpu):
        all_q, all_d_pos, all_d_neg = self.gather_all_embeddings(q, d_pos, d_neg, use_multi_gpu)   # losses.py:68-88
        scores_pos, scores_neg = self.compute_query_doc_scores(q, all_d_pos, all_d_neg)
        all_scores = torch.cat([scores_pos, scores_neg], dim=1)
        per_device_batch_size = q.size(0)
        gpu_id = torch.distributed.get_rank() if use_multi_gpu else 0
        nll = self.sliced_nll(all_scores, per_device_batch_size, gpu_id)
        if self.reverse:
            rev_scores = self.compute_doc_query_scores(d_pos, all_q)
            nll = nll + self.sliced_nll(rev_scores, per_device_batch_size, gpu_id)
        return nll


class DecomposedContrastiveLoss(BaseContrastiveLoss):
    """losses.py:148-202: multi-field contrastive loss over the mixture of per-field scores."""

    def __init__(self, temperature: float = 0.01, in_batch_negative: bool = True, reverse: bool = True,
                 all_gather_multi_gpu: bool = True, mixture_of_fields_layer: torch.nn.Module = None):
        super().__init__(temperature, in_batch_negative, reverse, all_gather_multi_gpu, multi_fields=True)
        self.mixture_of_fields_layer = mixture_of_fields_layer

    def forward(self, q: torch.Tensor, d_pos: torch.Tensor, d_neg: Optional[torch.Tensor]) -> torch.Tensor:
        use_multi_gpu = self.all_gather_multi_gpu and torch.distributed.is_initialized()
        if self.in_batch_negative:
            nll = self.in_batch_negative_loss(q, d_pos, d_neg, use_multi_gpu)
        else:
            nll = self.simple_loss(q, d_pos, d_neg)
        return self.distributed_reduce_mean_nll(nll, use_multi_gpu)

    def compute_query_doc_field_components(self, q, d_pos, d_neg) -> Tuple[torch.Tensor, torch.Tensor]:
        """[B,E], [P,F,E], [P,F,Neg,E] -> ([B,P,F], [B,P*Neg,F])  (losses.py:176-188)."""
        return field_components(q, d_pos, self.temperature), field_components(q, d_neg, self.temperature)

    def compute_query_doc_scores(self, q, d_pos, d_neg) -> Tuple[torch.Tensor, torch.Tensor]:
        pos_c, neg_c = self.compute_query_doc_field_components(q, d_pos, d_neg)          # losses.py:190-198
        return self.mixture_of_fields_layer(pos_c, q), self.mixture_of_fields_layer(neg_c, q)

    def compute_doc_query_scores(self, d_pos, q):
        """[B_loc,F,E], [Bq,E] -> [B_loc,Bq]: the same components seen from the queries' side (losses.py:199-202)."""
        return self.mixture_of_fields_layer(field_components(q, d_pos, self.temperature), q).t()

    def simple_loss(self, q, d_pos, d_neg):
        """No in-batch negatives (losses.py:90-109, multi_fields branch): every query against its own positive and
        its own negatives; d_pos [B,F,E], d_neg [B,F,Neg,E] (the layout the reference's permute(0,2,1,3) implies).
        All B x B(.Neg) pairs go through the kernels and the own-doc entries are selected - B is a training batch."""
        B, n_neg = q.size(0), d_neg.size(2)
        idx = torch.arange(B, device=q.device)
        pos_all = self.mixture_of_fields_layer(field_components(q, d_pos, self.temperature), q)      # [B,B]
        scores_pos = pos_all[idx, idx].unsqueeze(1)                                                     # [B,1]
        neg_all = self.mixture_of_fields_layer(field_components(q, d_neg, self.temperature), q)      # [B,B*Neg]
        scores_neg = neg_all.view(B, B, n_neg)[idx, idx]                                                # [B,Neg]
        all_scores = torch.cat([scores_pos, scores_neg], dim=1)
        return -torch.mean(torch.log_softmax(all_scores, dim=1)[:, 0])


def _unpickle(x):
    return pickle.loads(x) if isinstance(x, (bytes, bytearray)) else x


class HybridContrastiveLoss(DecomposedContrastiveLoss):
    """losses.py:204-360: dense components + per-field sparse scores (+ BatchNorm1d over fields) -> mixture."""

    def __init__(self, temperature: float = 0.01, in_batch_negative: bool = True, reverse: bool = True,
                 all_gather_multi_gpu: bool = True, mixture_of_fields_layer: torch.nn.Module = None,
                 sparse_indices_dict: Optional[Dict] = None, num_fields: int = 0, use_batchnorm: bool = False):
        super().__init__(temperature, in_batch_negative, reverse, all_gather_multi_gpu, mixture_of_fields_layer)
        self.sparse_indices_dict = sparse_indices_dict or {}
        self.bn = torch.nn.BatchNorm1d(num_fields, track_running_stats=True) if use_batchnorm else torch.nn.Identity()

    def forward(self, q, queries, d_pos, pos_docs, d_neg, neg_docs, query_ids, sparse_scores: Optional[Dict] = None):
        use_multi_gpu = self.all_gather_multi_gpu and torch.distributed.is_initialized()
        if self.in_batch_negative:
            nll = self.in_batch_negative_loss(q, queries, d_pos, pos_docs, d_neg, neg_docs, use_multi_gpu, query_ids,
                                              sparse_scores)
        else:
            nll = self.simple_loss(q, d_pos, d_neg)
        return self.distributed_reduce_mean_nll(nll, use_multi_gpu)

    def gather_all_embeddings(self, q, queries, query_ids, d_pos, pos_docs, d_neg, neg_docs, use_multi_gpu):
        queries, query_ids = _unpickle(queries), _unpickle(query_ids)                 # losses.py:242-273
        pos_docs, neg_docs = _unpickle(pos_docs), _unpickle(neg_docs)
        if use_multi_gpu:
            q = torch.cat(dist_F.all_gather(q), dim=0)
            d_pos = torch.cat(dist_F.all_gather(d_pos), dim=0)
            d_neg = torch.cat(dist_F.all_gather(d_neg), dim=0)
            world = torch.distributed.get_world_size()
            gathered = []
            for obj in (queries, pos_docs, neg_docs, query_ids):
                buf = [None] * world
                torch.distributed.all_gather_object(buf, obj)
                gathered.append([x for part in buf for x in part])
            queries, pos_docs, neg_docs, query_ids = gathered
        return q, list(queries), d_pos, list(pos_docs), d_neg, list(neg_docs), list(query_ids)

    def in_batch_negative_loss(self, q, queries, d_pos, pos_docs, d_neg, neg_docs, use_multi_gpu, query_ids,
                               sparse_scores):
        all_q, all_queries, all_d_pos, all_pos_ids, all_d_neg, all_neg_ids, all_query_ids = \
            self.gather_all_embeddings(q, queries, query_ids, d_pos, pos_docs, d_neg, neg_docs, use_multi_gpu)
        queries, query_ids, pos_docs = _unpickle(queries), _unpickle(query_ids), _unpickle(pos_docs)
        scores_pos, scores_neg = self.compute_query_doc_scores(q, queries, all_d_pos, all_pos_ids, all_d_neg,
                                                               all_neg_ids, query_ids, sparse_scores)
        all_scores = torch.cat([scores_pos, scores_neg], dim=1)                      # losses.py:290
        per_device_batch_size = q.size(0)
        gpu_id = torch.distributed.get_rank() if use_multi_gpu else 0
        nll = self.sliced_nll(all_scores, per_device_batch_size, gpu_id)
        if self.reverse:
            rev = self.compute_doc_query_scores(d_pos, pos_docs, all_q, all_queries, all_query_ids, sparse_scores)
            nll = nll + self.sliced_nll(rev, per_device_batch_size, gpu_id)
        return nll

    def compute_sparse_query_doc_scores(self, queries, doc_ids, query_ids, sparse_scores, device=None):
        """-> [QueryBatch, DocBatch, n_sparse] on the device (losses.py:303-325): cached precomputed scores when every
        query id has them, else ``score_batch`` of every sparse index (here the device BM25 scorer)."""
        device = device or "cuda"
        if len(self.sparse_indices_dict) == 0:
            return torch.empty(len(queries), len(doc_ids), 0, device=device)
        if sparse_scores and any(all(qid in by_field for qid in query_ids) for by_field in sparse_scores.values()):
            per_field = [si.score_batch_with_cache(query_ids, doc_ids, sparse_scores[name])
                         for name, si in self.sparse_indices_dict.items()]
        else:
            per_field = [si.score_batch(queries, doc_ids) for si in self.sparse_indices_dict.values()]
        return torch.stack([p.float() for p in per_field], dim=-1).to(device)

    def compute_query_doc_scores(self, q, queries, d_pos, pos_text, d_neg, neg_text, query_ids, sparse_scores):
        dense_pos, dense_neg = self.compute_query_doc_field_components(q, d_pos, d_neg)     # losses.py:327-350
        sparse_pos = self.compute_sparse_query_doc_scores(queries, pos_text, query_ids, sparse_scores, q.device)
        sparse_neg = self.compute_sparse_query_doc_scores(queries, neg_text, query_ids, sparse_scores, q.device)
        all_scores = torch.cat([torch.cat([dense_pos, sparse_pos], dim=-1),
                                torch.cat([dense_neg, sparse_neg], dim=-1)], dim=1)         # [B, N, F]
        normed = self.bn(all_scores.permute(0, 2, 1)).permute(0, 2, 1).contiguous()
        combined = self.mixture_of_fields_layer(normed, q)
        n_pos = dense_pos.size(1)
        return combined[:, :n_pos], combined[:, n_pos:]

    def compute_doc_query_scores(self, d_pos, pos_docs, q, queries, query_ids, sparse_scores):
        dense_rev = field_components(q, d_pos, self.temperature)                            # [Bq, B_loc, F]  352-360
        sparse_rev = self.compute_sparse_query_doc_scores(queries, pos_docs, query_ids, sparse_scores, q.device)
        all_scores = torch.cat([dense_rev, sparse_rev], dim=2)
        all_scores = self.bn(all_scores.permute(0, 2, 1)).permute(0, 2, 1).contiguous()
        return self.mixture_of_fields_layer(all_scores, q).t()
