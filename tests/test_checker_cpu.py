"""The at-scale fp32 checker (tests/checker.py) against the CPU oracle on shapes the oracle ranks in full.
Runs on the CPU: the checker is plain torch over the packed tensor, so a host copy of the layout
([tiles][F][128][dim_pad] bf16, include/mfar_b200.h) stands in for ``PackedCorpus``."""
import types

import numpy as np
import pytest
import torch

import mfar_oracle as O
from checker import Fp32Checker, assert_topk_parity_at_scale
from parity import assert_topk_parity


def pack_host(fields, dim_pad):
    n, d = fields[0].shape
    tiles = (n + 127) // 128
    data = torch.zeros((tiles, len(fields), 128, dim_pad), dtype=torch.bfloat16)
    for f, x in enumerate(fields):
        buf = torch.zeros((tiles * 128, dim_pad), dtype=torch.bfloat16)
        buf[:n, :d] = x.to(torch.bfloat16)
        data[:, f] = buf.view(tiles, 128, dim_pad)
    return types.SimpleNamespace(data=data.reshape(-1), n_docs=n, n_fields=len(fields), dim_pad=dim_pad, dim=d)


@pytest.mark.parametrize("N,d,Fd,Fs,Q,k,base", [(1000, 64, 3, 2, 5, 100, 0), (517, 96, 2, 0, 3, 50, 4096),
                                               (300, 64, 0, 2, 4, 20, 0)])
def test_checker_matches_oracle(N, d, Fd, Fs, Q, k, base):
    g = torch.Generator().manual_seed(N)
    fields = [O.round_bf16(torch.randn(N, d, generator=g)) for _ in range(Fd)]
    q = O.round_bf16(torch.randn(Q, d, generator=g))
    sp = None
    if Fs:
        sp = torch.where(torch.rand(Q, Fs, N, generator=g) < 0.9, torch.zeros(()), 5 * torch.rand(Q, Fs, N, generator=g)).half()
    W = 0.05 * torch.randn(d, Fd + Fs, generator=g)
    w = O.mixture_weights(q, W, True)
    mask = torch.ones(Fd + Fs, 1)
    mask[0] = 0
    ref_all = O.exhaustive_scores(q, fields, None if sp is None else sp.float(), w, mask)
    dim_pad = (d + 63) // 64 * 64
    pc = pack_host(fields, dim_pad) if Fd else None
    chk = Fp32Checker(pc, n_docs=N, doc_id_base=base, chunk_docs=256)
    qp = torch.nn.functional.pad(q, (0, dim_pad - d)).to(torch.bfloat16) if Fd else None
    wm = (w * mask.reshape(1, -1)).contiguous()
    s, i = chk.topk(qp, wm, k, sparse=sp, slack=16)
    assert s.shape == (Q, min(k + 16, N))
    # the checker's own top-k obeys the oracle's parity rule and its (score desc, id asc) order
    assert_topk_parity(s[:, :k].numpy(), i[:, :k].numpy(), ref_all.numpy(), k, id_offset=base)
    os_, oi = O.topk_sorted(ref_all, k)
    assert (i[:, :k] - base == oi).float().mean() > 0.99
    resc = chk.rescore(qp, wm, i[:, :k], sparse=sp)
    torch.testing.assert_close(resc, s[:, :k], rtol=1e-5, atol=1e-5)
    assert_topk_parity_at_scale(os_, oi + base, s, i, resc, k, N, base)


def test_at_scale_assertion_catches_a_dropped_winner():
    g = np.random.default_rng(0)
    Q, k, N = 2, 10, 1000
    ref_s = -np.sort(-g.standard_normal((Q, k + 8)).astype(np.float32), axis=1) + 5
    ref_i = np.stack([g.permutation(N)[: k + 8] for _ in range(Q)])
    got_s, got_i = ref_s[:, :k].copy(), ref_i[:, :k].copy()
    assert_topk_parity_at_scale(got_s, got_i, ref_s, ref_i, got_s.copy(), k, N)
    # drop rank 3, shift the tail up and append the (k+1)-th: a missing winner that does not tie the k-th score
    bad_s = np.concatenate([got_s[:, :3], ref_s[:, 4:k + 1]], axis=1)
    bad_i = np.concatenate([got_i[:, :3], ref_i[:, 4:k + 1]], axis=1)
    with pytest.raises(AssertionError, match="missing"):
        assert_topk_parity_at_scale(bad_s, bad_i, ref_s, ref_i, bad_s.copy(), k, N)
    # a wrong reported score
    wrong = got_s.copy()
    wrong[0, 2] *= 1.001
    with pytest.raises(AssertionError):
        assert_topk_parity_at_scale(wrong, got_i, ref_s, ref_i, got_s.copy(), k, N)
