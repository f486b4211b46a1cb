"""Doc-id-range sharding across GPUs and the cross-shard top-k merge.

The corpus is split into contiguous doc-id ranges, rank r owning ``[N*r//R, N*(r+1)//R)`` - the
split the reference already uses to shard corpus *encoding* (mfar/modeling/contrastive.py:470).
Queries / W / mask are replicated.  Each rank scores its shard with GLOBAL doc ids, the packed
(score,id) keys [Q,k] are exchanged with ONE all-gather (NCCL over NVLink), and every rank runs
the same merge kernel - replacing the reference's ``{rank}.qres`` files + barrier + rank-0 merge
(contrastive.py:616-631).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _native as nv


def shard_range(n_docs: int, rank: int, world: int) -> Tuple[int, int]:
    return n_docs * rank // world, n_docs * (rank + 1) // world


# ---- host-side mirror of the device key packing (csrc/common.cuh) - used by tests and debugging
def encode_keys(scores: np.ndarray, ids: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(scores, dtype=np.float32).view(np.uint32).astype(np.uint64)
    neg = (u & np.uint64(0x80000000)) != 0
    o = np.where(neg, (~u) & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))
    low = (~np.asarray(ids).astype(np.uint64)) & np.uint64(0xFFFFFFFF)
    return (o << np.uint64(32)) | low


def decode_keys(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    keys = np.asarray(keys).astype(np.uint64)
    o = (keys >> np.uint64(32)).astype(np.uint32)
    u = np.where((o & np.uint32(0x80000000)) != 0, o & np.uint32(0x7FFFFFFF), ~o)
    scores = u.astype(np.uint32).view(np.float32)
    ids = ((~keys) & np.uint64(0xFFFFFFFF)).astype(np.int64)
    empty = keys == 0
    return np.where(empty, -np.inf, scores).astype(np.float32), np.where(empty, -1, ids)


def all_gather_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """keys: int64 view of the packed uint64 keys [Q,k] -> [R,Q,k] (same on every rank)."""
    world = dist.get_world_size(group)
    keys = keys.contiguous()
    out = torch.empty((world * keys.shape[0],) + tuple(keys.shape[1:]), dtype=keys.dtype, device=keys.device)
    dist.all_gather_into_tensor(out, keys, group=group)          # concatenated along dim 0 (gloo and nccl)
    return out.view((world,) + tuple(keys.shape))


def merge_keys(all_keys: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[R,Q,k_in] packed keys (device) -> global (scores [Q,k], ids [Q,k]) via mfar_topk_merge."""
    nv.require_device(all_keys, "all_keys")
    R, Q, k_in = all_keys.shape
    scores = torch.empty((Q, k), dtype=torch.float32, device=all_keys.device)
    ids = torch.empty((Q, k), dtype=torch.int64, device=all_keys.device)
    nv.check(nv.lib().mfar_topk_merge(nv.ptr(all_keys.contiguous()), R, Q, k_in, k, 0, nv.ptr(scores), nv.ptr(ids),
                                      nv.stream()), "topk_merge")
    return scores, ids


class ShardedRetriever:
    """One process per GPU; wraps the rank-local ``MultiFieldRetriever`` (built over this rank's doc
    range with ``doc_id_base = shard_range(...)[0]``)."""

    def __init__(self, local, group=None):
        self.local = local
        self.group = group

    @torch.no_grad()
    def search(self, q_vecs, q_emb=None, sparse_local=None, top_k: Optional[int] = None):
        k = top_k or self.local.top_k
        k_local = min(k, self.local.n_docs)
        _, _, keys = self.local.search(q_vecs, q_emb, sparse_local, top_k=k_local, return_keys=True)
        if k_local < k:                                   # tiny shard: pad with empty keys
            keys = torch.nn.functional.pad(keys, (0, k - k_local))
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return merge_keys(keys.unsqueeze(0), k)
        return merge_keys(all_gather_keys(keys, self.group), k)
