"""The reference's CPU retrieval pipeline as a timed baseline (``bench.py --impl reference`` and ``cpu_baseline``).

TEST / MEASUREMENT INFRASTRUCTURE ONLY - never imported by the product package.

What is timed is ``RetrievalTrainingModule.trec_eval_step`` (mfar/modeling/contrastive.py:669-704) with the text
encoder replaced by a table lookup (query vectors are inputs of the hot path):

    for every field index:  retrieve_batch(query texts, top_k=100)                (672-674)
    per query:              union of the F hit lists (Python sets of doc keys)    (676-679)
                            score_batch([query], union keys) on every index       (681-683)
                            stack * mask -> LinearWeights -> torch.topk(100)      (685-696)

Two interchangeable sets of index classes run under the SAME driver:

  kind = "reference"  the reference's own ``DenseFlatIndex`` / ``MemoryMapDict`` / ``LinearWeights`` imported unmodified
                      from /root/reference through ``ref_import`` (build container only - the tree does not travel);
  kind = "port"       ``PortDenseIndex`` / ``PortSparseIndex`` below: a restatement that pays the same costs the
                      reference pays - an fp32 ``np.memmap`` per field on disk (mfar/data/util.py:35, modeling/
                      util.py:85-94), chunked ``q @ V^T`` + cat + topk per chunk (index.py:194-212), Python lists of
                      (key, score) tuples (214-222), a dict lookup per candidate key and a memmap fancy-index copy per
                      ``score_batch`` call (229-230), a re-encode of the query per call (228).  Round 1's port worked
                      on in-memory tensors with integer rows and skipped those costs.

``profiles/r2_cpu_arm_port_vs_reference.json`` holds both kinds timed on the same host by ``python
oracle/cpu_pipeline.py``.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time
from functools import reduce
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


class TableEncoder:
    """Stands in for ``SentenceTransformer.encode`` (index.py:187, 228): query text -> its stored vector."""

    def __init__(self, table: Dict[str, np.ndarray]):
        self.table = table

    def encode(self, texts, convert_to_tensor=True):
        return torch.from_numpy(np.stack([self.table[t] for t in texts]))


class PortDenseIndex:
    """Restatement of ``DenseFlatIndex`` (mfar/data/index.py:160-232) over an fp32 memmap."""

    def __init__(self, model, vectors, numeric_ids_to_keys, keys_to_numeric_ids, vector_batch_size=1048576):
        self.model, self.vectors = model, vectors
        self.id2key, self.key2id = numeric_ids_to_keys, keys_to_numeric_ids
        self.chunk = vector_batch_size

    def retrieve_batch(self, queries, top_k):
        q = torch.from_numpy(queries) if isinstance(queries, np.ndarray) else \
            self.model.encode(list(queries), convert_to_tensor=True)                  # index.py:184-187
        Q, n = q.size(0), self.vectors.shape[0]
        best_s = torch.zeros((Q, top_k), dtype=torch.float32)                          # index.py:192-193
        best_i = torch.zeros((Q, top_k), dtype=torch.int64)
        for lb in range(0, n, self.chunk):
            ub = min(n, lb + self.chunk)
            block = torch.from_numpy(self.vectors[lb:ub])                              # index.py:196
            s = torch.matmul(q, block.t())
            cat_s = torch.cat([best_s, s], dim=1)
            cat_i = torch.cat([best_i, torch.arange(lb, ub).unsqueeze(0).expand(Q, -1)], dim=1)
            top_s, pos = torch.topk(cat_s, top_k, dim=1, largest=True, sorted=True)    # index.py:203
            best_i = cat_i[torch.arange(Q).unsqueeze(1), pos]
            best_s = top_s[:, :top_k]
        rows, vals = best_i.tolist(), best_s.tolist()
        return [list(zip([self.id2key[j] for j in rows[i]], vals[i])) for i in range(len(queries))]   # index.py:214-222

    def score_batch(self, queries, keys):
        q = self.model.encode(queries, convert_to_tensor=True)                         # index.py:228
        rows = [self.key2id[k] for k in keys]                                          # index.py:229
        picked = torch.from_numpy(self.vectors[rows])                                  # index.py:230 (memmap fancy index)
        return torch.matmul(q, picked.t())


class PortSparseIndex:
    """Restatement of ``BM25sSparseIndex.retrieve_batch / score_batch`` (index.py:95-118) around precomputed
    per-query score vectors (BM25 arithmetic is an input of this path)."""

    def __init__(self, keys, score_table: Dict[str, np.ndarray]):
        self.keys, self.table = keys, score_table
        self.key2id = {k: i for i, k in enumerate(keys)}

    def retrieve_batch(self, queries, top_k):
        out = []
        for t in queries:
            s = self.table[t]
            idx = np.argpartition(-s, top_k - 1)[:top_k]
            idx = idx[np.argsort(-s[idx], kind="stable")]
            out.append([(self.keys[i], float(s[i])) for i in idx])
        return out

    def score_batch(self, queries, keys):
        rows = np.array([self.key2id.get(k, -1) for k in keys])                         # index.py:112
        missing = np.nonzero(rows < 0)[0]                                              # index.py:113
        picked = np.stack([self.table[t] for t in queries], axis=0)[:, rows]           # index.py:114-116
        picked[:, missing] = 0                                                         # index.py:117
        return torch.tensor(picked)


def eval_step(indices, layer, mask, qtexts, q_vecs, query_cond, k=100):
    """``trec_eval_step`` (contrastive.py:669-704) over the given index objects; returns per query (values, keys)."""
    all_hits = [index.retrieve_batch(qtexts, top_k=k) for index in indices]           # 672-674
    hits_ids = np.array([[[h[0] for h in hit] for hit in field] for field in all_hits])
    out = []
    for i, text in enumerate(qtexts):
        ids_set = [set(h) for h in hits_ids[:, i, :].tolist()]
        union = list(reduce(lambda a, b: a | b, ids_set))                              # 678-679
        rescored = [index.score_batch([text], union) for index in indices]            # 681-683
        all_tens = torch.stack([h.float() for h in rescored], dim=0).squeeze(1) * mask  # 685-686
        with torch.no_grad():
            scores = layer(all_tens.t(), torch.from_numpy(q_vecs[i:i + 1]) if query_cond else None)   # 694
        values, idx = torch.topk(scores, k=min(k, scores.shape[1]), dim=1)             # 696
        out.append((values.squeeze(0), [union[j] for j in idx.flatten().tolist()]))
    return out


class PortLinearWeights(torch.nn.Module):
    """``LinearWeights`` (mfar/modeling/weighting.py:3-29) restated."""

    def __init__(self, emb_size, num_fields, query_cond=False):
        super().__init__()
        self.query_cond = query_cond
        self.weight = torch.nn.Parameter(torch.ones(emb_size, num_fields))

    def forward(self, x, q):
        logits = torch.matmul(q, self.weight) if self.query_cond else self.weight.transpose(1, 0)
        return torch.sum(torch.softmax(logits, dim=1).unsqueeze(1) * x, dim=-1)


def build_indices(kind: str, tmp: str, fields: Sequence[np.ndarray], sparse: Optional[np.ndarray], q_vecs: np.ndarray,
                  W: np.ndarray, query_cond: bool):
    """Field memmaps on disk + index objects + mixture layer, built the way ``read_and_create_indices`` /
    ``on_eval_start`` do (mfar/modeling/util.py:83-101, contrastive.py:482-496)."""
    n = fields[0].shape[0] if len(fields) else sparse.shape[2]
    keys = [f"d{i}" for i in range(n)]
    key2id = {k: i for i, k in enumerate(keys)}
    qtexts = [f"q{i}" for i in range(q_vecs.shape[0])]
    enc = TableEncoder({t: q_vecs[i] for i, t in enumerate(qtexts)})
    if kind == "reference":
        import ref_import
        DenseFlatIndex, BM25sSparseIndex, MemoryMapDict, LinearWeights, _ = ref_import.load()
    indices = []
    for f, x in enumerate(fields):
        path = os.path.join(tmp, f"field{f}.npy")
        mm = np.memmap(path, dtype=np.float32, mode="w+", shape=x.shape)               # headerless fp32 (data/util.py:35)
        mm[:] = x
        mm.flush()
        del mm
        if kind == "reference":
            store = MemoryMapDict(path, keys=keys, shape=x.shape)
            indices.append(DenseFlatIndex(enc, store.file, numeric_ids_to_keys=keys, keys_to_numeric_ids=key2id))
        else:
            vec = np.memmap(path, dtype=np.float32, mode="r+", shape=x.shape)
            indices.append(PortDenseIndex(enc, vec, keys, key2id))
    n_sparse = 0 if sparse is None else sparse.shape[1]
    for j in range(n_sparse):
        table = {t: sparse[i, j] for i, t in enumerate(qtexts)}
        if kind == "reference":
            class _BM25:                                   # stands in for bm25s.BM25 (index.py:75, 99)
                def __init__(self, tb): self.tb = tb
                def get_scores(self, toks): return self.tb[toks[0] if isinstance(toks, list) else toks]
                def retrieve(self, toks, k, show_progress=False, backend_selection="numpy"):
                    s = np.stack([self.tb[t[0] if isinstance(t, list) else t] for t in toks])
                    idx = np.argpartition(-s, k - 1, axis=1)[:, :k]          # bm25s' numpy backend: partition, then sort k
                    idx = np.take_along_axis(idx, np.argsort(-np.take_along_axis(s, idx, axis=1), axis=1, kind="stable"), axis=1)
                    return idx, np.take_along_axis(s, idx, axis=1)
            indices.append(BM25sSparseIndex(keys, _BM25(table), stemmer=None))
        else:
            indices.append(PortSparseIndex(keys, table))
    F = len(fields) + n_sparse
    L = LinearWeights if kind == "reference" else PortLinearWeights
    layer = L(W.shape[0], F, query_cond=True) if query_cond else L(F, 1)
    with torch.no_grad():
        layer.weight.copy_(torch.from_numpy(W))
    return indices, layer, qtexts


def time_pipeline(kind: str, n_docs: int, n_dense: int, n_sparse: int, Q: int, dim: int, seed: int, steps: int,
                  warmup: int, tmp_root: Optional[str] = None, make_rows=None) -> List[float]:
    """Seconds per ``eval_step`` over an ``n_docs`` sample (list of ``steps`` timings, after ``warmup`` untimed ones).
    ``make_rows(n, dim, field)`` may supply the field rows (e.g. drawn on a GPU); default: torch CPU randn."""
    g = torch.Generator().manual_seed(seed)
    mu = torch.randn(dim, generator=g)
    if make_rows is None:
        def make_rows(n, d, f):
            return (torch.randn(n, d, generator=g) + 0.5 * mu).to(torch.bfloat16).float().numpy()
    fields = [make_rows(n_docs, dim, f) for f in range(n_dense)]
    q = (torch.randn(Q, dim, generator=g) + 0.5 * mu).to(torch.bfloat16).float().numpy()
    sparse = None
    if n_sparse:
        u = torch.rand(Q, n_sparse, n_docs, generator=g)
        sparse = torch.where(u < 0.95, torch.zeros(()), 4.0 * torch.rand(Q, n_sparse, n_docs, generator=g)).numpy()
    W = (0.05 * torch.randn(dim, n_dense + n_sparse, generator=g)).numpy()
    mask = torch.ones(n_dense + n_sparse, 1)
    with tempfile.TemporaryDirectory(dir=tmp_root) as tmp:
        indices, layer, qtexts = build_indices(kind, tmp, fields, sparse, q, W, True)
        del fields
        for _ in range(warmup):
            eval_step(indices, layer, mask, qtexts, q, True)
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            eval_step(indices, layer, mask, qtexts, q, True)
            times.append(time.perf_counter() - t0)
        del indices
    return times


if __name__ == "__main__":          # port vs the reference's own classes on this host (build container)
    import json
    import ref_import
    torch.set_num_threads(os.cpu_count() or 1)
    rows = []
    for n, Fd, Fs, Q in ((50_000, 8, 0, 64), (200_000, 8, 0, 64), (20_000, 8, 8, 32)):
        rec = {"n_docs": n, "n_dense": Fd, "n_sparse": Fs, "Q": Q, "cores": torch.get_num_threads()}
        for kind in (["reference"] if ref_import.available() else []) + ["port"]:
            t = time_pipeline(kind, n, Fd, Fs, Q, 768, 1234, steps=3, warmup=1)
            rec[kind + "_s_per_step"] = sorted(t)[len(t) // 2]
        rows.append(rec)
        print(json.dumps(rec), flush=True)
