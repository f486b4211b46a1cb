"""Generate tests/golden/precompute/*.npz by EXECUTING the reference's own
``mfar.commands.precompute_bm25s_scores.precompute_score_for_field`` (unmodified, its multiprocessing pool and file
writes included) over the reference's ``BM25sSparseIndex`` wrapped around a fake ``bm25s.BM25`` that serves given
score vectors (the BM25 arithmetic is an input to this step).

Run in the build container only:   python oracle/make_golden_precompute.py
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden", "precompute")


class FakeBM25:
    """bm25s.BM25.get_scores (index.py:80).  With the stubbed ``bm25s.tokenize`` (ref_import.py) the "tokens" the
    reference passes are the query text itself, so the table is keyed by text."""

    def __init__(self, table):
        self.table = table

    def get_scores(self, query_tokens):
        return self.table[query_tokens]


def _load_reference_command():
    ref_import.load()

    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
    stub("fire", Fire=lambda f: None)
    sys.modules["sentence_transformers"].__path__ = []           # let "sentence_transformers.models" resolve
    stub("sentence_transformers.models", Normalize=object, Pooling=object)
    from mfar.commands import precompute_bm25s_scores as P
    from mfar.data.index import BM25sSparseIndex
    return P, BM25sSparseIndex


def make_case(name, seed, N, Q, density, safe_frac, extremes=False):
    P, BM25sSparseIndex = _load_reference_command()
    rng = np.random.RandomState(seed)
    scores = np.where(rng.rand(Q, N) < density, rng.gamma(2.0, 2.0, (Q, N)), 0.0).astype(np.float32)
    if extremes:                                                  # f16 underflow / overflow / subnormal / negative
        scores[0, 1] = 1e-9
        scores[0, 2] = 7e4
        scores[1, 3] = 3e-6
        scores[1, 4] = -2.5
        scores[2 % Q, N - 1] = 65519.9
    qids = (1000 + 37 * np.arange(Q)).astype(np.int64)
    rng.shuffle(qids)
    safe = set(int(d) for d in rng.choice(N + 50, size=max(1, int(safe_frac * N)), replace=False))
    if extremes:
        safe |= {1, 2, 3, 4, N - 1}
    safe = sorted(safe)
    texts = [f"query text {i}" for i in range(Q)]
    index = BM25sSparseIndex([str(i) for i in range(N)], FakeBM25({t: scores[i] for i, t in enumerate(texts)}),
                             stemmer=None)
    train_queries = {int(q): t for q, t in zip(qids, texts)}
    with tempfile.TemporaryDirectory() as tmp:
        P.precompute_score_for_field(index, set(safe), train_queries, tmp, "f_sparse")
        keys = np.load(os.path.join(tmp, "f_sparse_keys_bm25.npy"))
        vals = np.load(os.path.join(tmp, "f_sparse_vals_bm25.npy"))
    assert keys.dtype == np.int32 and vals.dtype == np.float16
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), scores=scores, qids=qids,
                        safe=np.asarray(safe, np.int64), ref_keys=keys.reshape(-1, 2), ref_vals=vals,
                        meta=np.asarray(json.dumps(dict(name=name, seed=seed, N=N, Q=Q))))
    print(f"wrote {name}: N={N} Q={Q} nnz={len(vals)}")


if __name__ == "__main__":
    make_case("pre_small", 1, N=300, Q=5, density=0.2, safe_frac=0.5)
    make_case("pre_extremes", 2, N=70, Q=3, density=0.3, safe_frac=0.9, extremes=True)
    make_case("pre_multiseg", 3, N=9001, Q=4, density=0.05, safe_frac=0.3)
    make_case("pre_empty", 4, N=130, Q=2, density=0.0, safe_frac=0.5)
