"""Test-only fp32 checker for corpus-sized shards (TEST INFRASTRUCTURE, like ``oracle/``: never imported by the
product package; ``bench.py`` uses it after its timed loop to verify what it timed).

The CPU oracle (``oracle/mfar_oracle.py``) ranks every doc, which takes minutes above ~100k docs.  This module
restates the same exhaustive formulas (mfar/data/index.py:197,231 per-field ``q . v``; contrastive.py:686 ``* mask``;
weighting.py:28-29 softmax mixture; contrastive.py:696 ``torch.topk``) with plain torch on the GPU, independent of
every scoring kernel of the library:

  * the packed corpus is read back through ``mfar_corpus_unpack_rows``-equivalent indexing of ``PackedCorpus.data``
    (bit-exact bf16 -> fp32, tests/test_gpu_parity.py::test_pack_unpack_roundtrip_bit_exact pins the layout),
  * ``torch.matmul`` in fp32 with TF32 disabled (CUDA-core FFMA, no tensor cores), 64k docs per chunk,
  * an exact running top-(k + slack) ordered by (score desc, doc id asc) - the oracle's ``topk_sorted`` rule.

``assert_topk_parity_at_scale`` is ``tests/parity.py::assert_topk_parity`` for results too large to hold ``[Q,N]``
oracle scores on the host: a doc missing from the kernel's top-k must near-tie the oracle's k-th score, an extra doc
must near-tie it from below (its oracle score is re-computed from the stored vectors), reported scores must equal the
re-computed ones, rank order may differ only inside near-tie groups.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from parity import SCORE_RTOL, TIE_REL

TILE = 128


class Fp32Checker:
    def __init__(self, corpus, n_docs: Optional[int] = None, doc_id_base: int = 0, chunk_docs: int = 65536):
        """corpus: ``PackedCorpus`` (or None for a sparse-only scorer, then pass n_docs)."""
        self.pc = corpus
        self.n_docs = corpus.n_docs if corpus is not None else int(n_docs)
        self.base = int(doc_id_base)
        self.chunk = int(chunk_docs) // TILE * TILE
        if corpus is not None:
            n_tiles = (corpus.n_docs + TILE - 1) // TILE
            self.view = corpus.data.view(n_tiles, corpus.n_fields, TILE, corpus.dim_pad)

    # ---- per-field fp32 rows of a doc range / of arbitrary rows, straight from the packed bf16 tensor
    def _field_rows(self, f: int, lo: int, hi: int) -> torch.Tensor:
        t0, t1 = lo // TILE, (hi + TILE - 1) // TILE
        x = self.view[t0:t1, f].reshape(-1, self.pc.dim_pad).float()
        return x[lo - t0 * TILE: hi - t0 * TILE]

    def _mix_chunk(self, q32, w, sparse, lo, hi, field_begin, n_dense):
        Q = w.shape[0]
        acc = torch.zeros((Q, hi - lo), dtype=torch.float32, device=w.device)
        n_sparse = 0 if sparse is None else sparse.shape[1]
        for j in range(n_sparse):                         # the kernels seed the accumulator with the sparse term
            acc.addcmul_(w[:, n_dense + j: n_dense + j + 1], sparse[:, j, lo:hi].float())
        for f in range(n_dense):
            acc.addcmul_(w[:, f:f + 1], q32 @ self._field_rows(field_begin + f, lo, hi).t())
        return acc

    @torch.no_grad()
    def topk(self, q_bf16: Optional[torch.Tensor], w: torch.Tensor, k: int, sparse: Optional[torch.Tensor] = None,
             slack: int = 64, field_begin: int = 0, n_dense: Optional[int] = None
             ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Exact fp32 top-(k+slack): (scores [Q,K2], GLOBAL ids [Q,K2]) sorted by (score desc, id asc).
        w: [Q, n_dense + n_sparse] fp32 mixture weights with the mask already multiplied in."""
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            n_dense = (self.pc.n_fields if self.pc is not None else 0) if n_dense is None else n_dense
            K2 = min(k + slack, self.n_docs)
            Q = w.shape[0]
            dev = w.device
            q32 = q_bf16.float() if q_bf16 is not None else None
            best_s = torch.full((Q, 0), 0.0, device=dev)
            best_i = torch.zeros((Q, 0), dtype=torch.int64, device=dev)
            for lo in range(0, self.n_docs, self.chunk):
                hi = min(self.n_docs, lo + self.chunk)
                acc = self._mix_chunk(q32, w, sparse, lo, hi, field_begin, n_dense)
                kk = min(K2, hi - lo)
                cs, ci = torch.topk(acc, kk, dim=1)
                best_s = torch.cat([best_s, cs], dim=1)
                best_i = torch.cat([best_i, ci + lo], dim=1)
                if best_s.shape[1] > 4 * K2:
                    best_s, best_i = self._reduce(best_s, best_i, K2)
            best_s, best_i = self._reduce(best_s, best_i, K2)
            return best_s, best_i + self.base
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    @staticmethod
    def _reduce(s, i, K2):
        # (score desc, id asc): stable sort by id, then stable sort by score descending
        o = torch.sort(i, dim=1, stable=True).indices
        s, i = s.gather(1, o), i.gather(1, o)
        o = torch.sort(s, dim=1, descending=True, stable=True).indices[:, :K2]
        return s.gather(1, o), i.gather(1, o)

    @torch.no_grad()
    def rescore(self, q_bf16: Optional[torch.Tensor], w: torch.Tensor, ids: torch.Tensor,
                sparse: Optional[torch.Tensor] = None, field_begin: int = 0, n_dense: Optional[int] = None
                ) -> torch.Tensor:
        """fp32 mixture scores of the given GLOBAL ids [Q,m] (one row of ids per query)."""
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            n_dense = (self.pc.n_fields if self.pc is not None else 0) if n_dense is None else n_dense
            rows = (ids - self.base).clamp(0, self.n_docs - 1)
            Q, m = rows.shape
            acc = torch.zeros((Q, m), dtype=torch.float32, device=w.device)
            n_sparse = 0 if sparse is None else sparse.shape[1]
            for j in range(n_sparse):
                acc.addcmul_(w[:, n_dense + j: n_dense + j + 1], sparse[:, j, :].gather(1, rows).float())
            if n_dense:
                q32 = q_bf16.float()
                for f in range(n_dense):
                    x = self.view[rows // TILE, field_begin + f, rows % TILE].float()        # [Q,m,dim]
                    acc.addcmul_(w[:, f:f + 1], torch.einsum("qmd,qd->qm", x, q32))
            return acc
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev


def assert_topk_parity_at_scale(got_scores, got_ids, ref_scores, ref_ids, rescored, k, n_docs, id_base=0,
                                tie_rel=TIE_REL, rtol=SCORE_RTOL, what=""):
    """got_* [Q,k] from the CUDA path; ref_* [Q,K2>=k] = the checker's exact top-(k+slack); rescored [Q,k] = the
    checker's fp32 score of every returned id.  Same rules as ``parity.assert_topk_parity``."""
    gs = np.asarray(got_scores.detach().cpu() if torch.is_tensor(got_scores) else got_scores, dtype=np.float64)
    gi = np.asarray(got_ids.detach().cpu() if torch.is_tensor(got_ids) else got_ids, dtype=np.int64)
    rs = np.asarray(ref_scores.detach().cpu() if torch.is_tensor(ref_scores) else ref_scores, dtype=np.float64)
    ri = np.asarray(ref_ids.detach().cpu() if torch.is_tensor(ref_ids) else ref_ids, dtype=np.int64)
    rr = np.asarray(rescored.detach().cpu() if torch.is_tensor(rescored) else rescored, dtype=np.float64)
    Q = gs.shape[0]
    assert gs.shape == (Q, k) and gi.shape == (Q, k) and rs.shape[0] == Q and rs.shape[1] >= k
    n_swapped = 0
    for q in range(Q):
        tag = f"{what} q{q}"
        scale = max(1e-30, abs(rs[q, 0]), abs(rs[q, -1]))
        tol = tie_rel * scale
        ids = gi[q]
        assert len(set(ids.tolist())) == k, f"{tag}: duplicate ids"
        assert ids.min() >= id_base and ids.max() < id_base + n_docs, f"{tag}: id out of range"
        np.testing.assert_allclose(gs[q], rr[q], rtol=rtol, atol=rtol * scale, err_msg=f"{tag}: reported scores")
        assert np.all(np.diff(gs[q]) <= rtol * scale), f"{tag}: not sorted"
        kth = rs[q, k - 1]
        top = ri[q, :k]
        ref_score_of = dict(zip(ri[q].tolist(), rs[q].tolist()))
        missing = set(top.tolist()) - set(ids.tolist())
        for m in missing:
            assert ref_score_of[m] - kth <= tol, f"{tag}: doc {m} (score {ref_score_of[m]}) missing, k-th is {kth}"
        extra = set(ids.tolist()) - set(top.tolist())
        pos = {int(d): p for p, d in enumerate(ids)}
        for e in extra:
            assert kth - rr[q, pos[e]] <= tol, f"{tag}: doc {e} (score {rr[q, pos[e]]}) should not be in the top-{k}"
        assert np.all(np.abs(rr[q] - rs[q, :k]) <= tol + rtol * scale), f"{tag}: rank order differs beyond near-ties"
        n_swapped += int((ids != top).sum())
    return n_swapped
