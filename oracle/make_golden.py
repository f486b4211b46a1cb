"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN CLASSES (unmodified, imported
from /root/reference through oracle/ref_import.py) on small seeded inputs.

Run in the build container only:   python oracle/make_golden.py
The reference has no tests / golden vectors for this path, so these files are what pins the
oracle (and through it the CUDA path).  What is executed from the reference:
  * MemoryMapDict                         mfar/data/util.py:28-59   (headerless fp32 memmap)
  * DenseFlatIndex.retrieve_batch         mfar/data/index.py:181-222
  * DenseFlatIndex.score_batch            mfar/data/index.py:227-232
  * BM25sSparseIndex.score_batch / retrieve_batch  mfar/data/index.py:95-118
        (with a fake ``bm25s.BM25`` object serving precomputed score vectors - BM25
         arithmetic is an INPUT to this path, see oracle/mfar_oracle.py header)
  * LinearWeights.forward                 mfar/modeling/weighting.py:17-29
  * resolve_fields                        mfar/data/schema.py:96-134
and a line-by-line driver of trec_eval_step (mfar/modeling/contrastive.py:669-704; that
module itself cannot be imported without Lightning/sentence-transformers) that calls those
reference objects in the reference's order.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
from functools import reduce

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def bf16_round(a: np.ndarray) -> np.ndarray:
    return torch.from_numpy(a.astype(np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


class FakeEncoder:
    """Stands in for SentenceTransformer.encode (index.py:187, 228): text -> fixed vector."""

    def __init__(self, table):
        self.table = table

    def encode(self, texts, convert_to_tensor=True):
        return torch.from_numpy(np.stack([self.table[t] for t in texts]))


class FakeBM25:
    """Stands in for bm25s.BM25: get_scores (index.py:75) / retrieve (index.py:99)."""

    def __init__(self, table):
        self.table = table          # query text -> np.float32 [N]

    def get_scores(self, query_tokens):
        return self.table[query_tokens]

    def retrieve(self, query_tokens, k, show_progress=False, backend_selection="numpy"):
        s = np.stack([self.table[t] for t in query_tokens])
        idx = np.argsort(-s, axis=1, kind="stable")[:, :k]
        return idx, np.take_along_axis(s, idx, axis=1)


def make_case(name, seed, N, d, Fd, Fs, Q, k, query_cond, mask_zero=(), negative=False, w_ones=False):
    DenseFlatIndex, BM25sSparseIndex, MemoryMapDict, LinearWeights, _ = ref_import.load()
    rng = np.random.RandomState(seed)
    mu = rng.standard_normal(d).astype(np.float32)
    fields = bf16_round(rng.standard_normal((Fd, N, d)).astype(np.float32) + 0.5 * mu)
    q = bf16_round(rng.standard_normal((Q, d)).astype(np.float32) + 0.5 * mu)
    if negative:   # most scores below zero -> exercises the (0.0, row 0) init quirk, index.py:192-193
        fields = bf16_round(-np.abs(fields))
        q = bf16_round(np.abs(q))
        fields[:, : max(2, k // 3), :] *= -1.0
    sparse = np.where(rng.rand(Q, Fs, N) < 0.9, 0.0, rng.gamma(2.0, 2.0, (Q, Fs, N))).astype(np.float16).astype(np.float32)
    F = Fd + Fs
    if query_cond:
        W = np.ones((d, F), np.float32) if w_ones else (0.05 * rng.standard_normal((d, F))).astype(np.float32)
    else:
        W = np.ones((F, 1), np.float32) if w_ones else rng.standard_normal((F, 1)).astype(np.float32)
    mask = np.ones((F, 1), np.float32)
    for m in mask_zero:
        mask[m] = 0.0

    keys = [f"d{i}" for i in range(N)]
    qtexts = [f"q{i}" for i in range(Q)]
    enc = FakeEncoder({t: q[i] for i, t in enumerate(qtexts)})

    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        indices = []
        # ---- store + dense indices, built the way read_and_create_indices does (modeling/util.py:83-101)
        for f in range(Fd):
            path = os.path.join(tmp, f"field{f}.npy")
            with open(path, "w"):
                pass
            vec = MemoryMapDict(path, keys=keys, shape=(N, d))
            for i, key in enumerate(keys):           # contrastive.py:490
                vec[key] = fields[f, i]
            vec.close()
            vec.reopen()
            assert os.path.getsize(path) == N * d * 4
            indices.append(DenseFlatIndex(enc, vec.file, numeric_ids_to_keys=keys,
                                          keys_to_numeric_ids={k_: i for i, k_ in enumerate(keys)},
                                          vector_batch_size=max(7, N // 3)))   # several chunks
        for j in range(Fs):
            table = {t: sparse[i, j] for i, t in enumerate(qtexts)}
            indices.append(BM25sSparseIndex(keys, FakeBM25(table), stemmer=None))

        # ---- a1: per-field retrieve_batch via the ndarray path (index.py:184-185)
        rs, rr = [], []
        for f in range(Fd):
            hits = indices[f].retrieve_batch(q, top_k=k)
            rs.append([[h[1] for h in hit] for hit in hits])
            rr.append([[int(h[0][1:]) for h in hit] for hit in hits])
        out["ref_retrieve_scores"] = np.asarray(rs, np.float32).reshape(Fd, Q, k)
        out["ref_retrieve_rows"] = np.asarray(rr, np.int64).reshape(Fd, Q, k)

        # ---- a2/a3: score_batch on a candidate list (with one unknown key for the sparse side)
        cand_rows = rng.choice(N, size=min(N, 17), replace=False).tolist()
        cand_keys = [keys[r] for r in cand_rows]
        out["cand_rows"] = np.asarray(cand_rows, np.int64)
        out["ref_score_batch"] = np.stack(
            [indices[f].score_batch(qtexts, cand_keys).numpy() for f in range(Fd)]) if Fd else np.zeros((0, Q, len(cand_rows)), np.float32)
        if Fs:
            out["ref_sparse_score_batch"] = np.stack(
                [indices[Fd + j].score_batch(qtexts, cand_keys + ["__missing__"]).numpy() for j in range(Fs)])

        # ---- a5 + a7, exhaustive form: reference per-field scores of ALL docs -> * mask -> LinearWeights
        layer = LinearWeights(d, F, query_cond=True) if query_cond else LinearWeights(F, 1)
        with torch.no_grad():
            layer.weight.copy_(torch.from_numpy(W))
        all_f = torch.stack([indices[f].score_batch(qtexts, keys).float() for f in range(F)], dim=0)  # [F,Q,N]
        x = (all_f * torch.from_numpy(mask).unsqueeze(-1)).permute(1, 2, 0).contiguous()              # [Q,N,F]
        with torch.no_grad():
            mix = layer(x, torch.from_numpy(q) if query_cond else None)
        out["ref_mix_all"] = mix.numpy().astype(np.float32)                                           # [Q,N]

        # ---- a6: trec_eval_step driver (contrastive.py:669-704), reference objects, reference order
        all_hits = []
        for index in indices:
            all_hits.append(index.retrieve_batch(qtexts, top_k=k))
        hits_ids = np.array([[[h[0] for h in hit] for hit in field] for field in all_hits])
        uv, ur = [], []
        union_raises = False
        for i in range(Q):
            ids_set = [set(h) for h in hits_ids[:, i, :].tolist()]
            all_ids_set = list(reduce(lambda a, b: a | b, ids_set))
            new_hits = [index.score_batch([qtexts[i]], all_ids_set) for index in indices]
            all_tens = torch.stack([h.float() for h in new_hits], dim=0).squeeze(1)
            all_tens = all_tens * torch.from_numpy(mask)
            with torch.no_grad():
                scores = layer(all_tens.t(), torch.from_numpy(q[i:i + 1]) if query_cond else None)
            if scores.shape[1] < k:
                # duplicates of row 0 from the zero-init quirk can leave the union smaller than k;
                # the reference's torch.topk (contrastive.py:696) then raises.  Recorded, not papered over.
                union_raises = True
                break
            values, idx = torch.topk(scores, k=k, dim=1)
            uv.append(values.squeeze(0).numpy())
            ur.append([int(all_ids_set[j][1:]) for j in idx.flatten().tolist()])
        out["ref_union_raises"] = np.asarray(union_raises)
        if not union_raises:
            out["ref_union_vals"] = np.asarray(uv, np.float32)
            out["ref_union_rows"] = np.asarray(ur, np.int64)

    out.update(fields=fields, q=q, sparse=sparse, W=W, mask=mask,
               meta=np.asarray(json.dumps(dict(name=name, seed=seed, N=N, d=d, Fd=Fd, Fs=Fs, Q=Q, k=k,
                                               query_cond=bool(query_cond)))))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    # bf16-representable fp32 / fp16-representable fp32 compress well
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), **out)
    print(f"wrote {name}: N={N} d={d} Fd={Fd} Fs={Fs} Q={Q} k={k}")


def make_schema_golden():
    *_, resolve_fields = ref_import.load()
    cases = {}
    for dataset in ["prime", "mag", "amazon", "data/stark/amazon_small"]:
        for names in ["all_dense", "all_sparse", "all_dense,all_sparse", "single_dense", "single_sparse,single_dense",
                      "title_dense,title_sparse"]:
            try:
                r = resolve_fields(names, dataset)
                cases[f"{dataset}|{names}"] = [[k_, f.name, f.field_type.name, f.max_seq_length] for k_, f in r.items()]
            except ValueError as e:
                cases[f"{dataset}|{names}"] = f"ValueError"
    with open(os.path.join(GOLDEN_DIR, "resolve_fields.json"), "w") as fp:
        json.dump(cases, fp, indent=0, sort_keys=True)
    print("wrote resolve_fields.json", len(cases))


if __name__ == "__main__":
    torch.manual_seed(0)
    make_case("tiny_hybrid_qc", 11, N=300, d=64, Fd=3, Fs=2, Q=5, k=10, query_cond=True)
    make_case("tiny_hybrid_masked", 12, N=257, d=64, Fd=3, Fs=2, Q=4, k=10, query_cond=True, mask_zero=(1, 4))
    make_case("d768_dense_static", 13, N=160, d=768, Fd=2, Fs=0, Q=3, k=20, query_cond=False)
    make_case("zero_init_quirk", 14, N=60, d=16, Fd=2, Fs=0, Q=3, k=10, query_cond=False, negative=True, w_ones=True)
    make_case("uniform_ones_qc", 15, N=130, d=32, Fd=4, Fs=1, Q=2, k=8, query_cond=True, w_ones=True)
    make_schema_golden()
