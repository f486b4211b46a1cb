"""world_size-2 gloo test of the multi-GPU host logic: doc-range sharding, key exchange with one
all-gather, merge - checked against the oracle's global top-k.  CPU only: the per-shard scores come
from the oracle here (the CUDA path is exercised by the gpu tests; the merge kernel by
test_gpu_parity::test_virtual_shards_merge)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    for p in (os.path.join(ROOT, "multifield-adaptive-retrieval_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import mfar_oracle as O
    from mfar_b200.dist import all_gather_keys, decode_keys, encode_keys, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)                # same global corpus on every rank
    N, d, F, Q, k = 1000, 32, 3, 4, 20
    fields = [O.round_bf16(torch.randn(N, d, generator=g)) for _ in range(F)]
    q = O.round_bf16(torch.randn(Q, d, generator=g))
    w = torch.softmax(torch.randn(Q, F, generator=g), dim=1)
    lo, hi = shard_range(N, rank, world)
    s_local = O.exhaustive_scores(q, [f[lo:hi] for f in fields], None, w)
    v, i = O.topk_sorted(s_local, k)
    keys = torch.from_numpy(encode_keys(v.numpy(), i.numpy() + lo).view(np.int64))
    allk = all_gather_keys(keys)                        # [R,Q,k]
    assert allk.shape == (world, Q, k)
    flat = allk.numpy().view(np.uint64).transpose(1, 0, 2).reshape(Q, -1)
    top = -np.sort(-flat.astype(np.uint64), axis=1)[:, :k] if False else np.sort(flat, axis=1)[:, ::-1][:, :k]
    ms, mi = decode_keys(top)
    gs, gi = O.topk_sorted(O.exhaustive_scores(q, fields, None, w), k)
    np.testing.assert_array_equal(mi, gi.numpy())
    np.testing.assert_array_equal(ms, gs.numpy())
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.destroy_process_group()


def test_sharded_topk_merge_gloo_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


# ---------------------------------------------------------------------------------------------------------------
# ShardedRetriever.search host logic (k clamp on tiny shards, empty-key padding, all-gather route) with a stand-in
# for the rank-local CUDA retriever and for the device merge kernel - both are exercised by the gpu tests.
class _OracleShard:
    """Rank-local retriever stand-in: oracle scores of this rank's doc range -> packed keys with global ids."""

    def __init__(self, fields, w, lo, hi, top_k):
        self.fields, self.w, self.lo, self.hi = fields, w, lo, hi
        self.n_docs, self.top_k = hi - lo, top_k

    def search(self, q_vecs, q_emb=None, sparse=None, top_k=None, return_keys=False, sparse_tokens=None):
        import mfar_oracle as O
        from mfar_b200.dist import encode_keys
        s = O.exhaustive_scores(q_vecs, [f[self.lo:self.hi] for f in self.fields], None, self.w)
        v, i = O.topk_sorted(s, top_k)
        keys = torch.from_numpy(encode_keys(v.numpy(), i.numpy() + self.lo).view(np.int64))
        return v, i + self.lo, keys


def _merge_keys_numpy(all_keys, k):
    from mfar_b200.dist import decode_keys
    R, Q, k_in = all_keys.shape
    flat = all_keys.numpy().view(np.uint64).transpose(1, 0, 2).reshape(Q, -1)
    top = np.sort(flat, axis=1)[:, ::-1][:, :k]
    s, i = decode_keys(top)
    return torch.from_numpy(s.copy()), torch.from_numpy(i.copy())


def _sharded_worker(rank, world, port, out_dir):
    for p in (os.path.join(ROOT, "multifield-adaptive-retrieval_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import mfar_oracle as O
    import mfar_b200.dist as D
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    D.merge_keys = _merge_keys_numpy                    # the device merge kernel's contract, on the host
    g = torch.Generator().manual_seed(9)
    d, F, Q, k = 16, 2, 3, 20
    w = torch.softmax(torch.randn(Q, F, generator=g), dim=1)
    q = O.round_bf16(torch.randn(Q, d, generator=g))
    for N in (500, 25):                                 # 25 docs over 2 ranks: shards (12, 13) are smaller than k = 20
        fields = [O.round_bf16(torch.randn(N, d, generator=g)) for _ in range(F)]
        lo, hi = D.shard_range(N, rank, world)
        sr = D.ShardedRetriever(_OracleShard(fields, w, lo, hi, k))
        s, i = sr.search(q, top_k=k)
        gs, gi = O.topk_sorted(O.exhaustive_scores(q, fields, None, w), k)
        assert torch.equal(i, gi) and torch.equal(s, gs), (N, rank)
    if rank == 0:
        open(os.path.join(out_dir, "ok2"), "w").write("ok")
    dist.destroy_process_group()


def test_sharded_retriever_host_logic_gloo_world2(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok2").read_text() == "ok"
