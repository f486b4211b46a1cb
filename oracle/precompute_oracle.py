"""CPU oracle for the producer of the precomputed-BM25 score files (numpy).

TEST INFRASTRUCTURE ONLY - imported by ``tests/`` (and ``oracle/make_golden_precompute.py``), never by the product
package.  Restates, from dense per-query score vectors (``bm25s.BM25.get_scores`` output, an INPUT here):

  mfar/data/index.py:78-84                           BM25sSparseIndex.get_scores_sparse
  mfar/commands/precompute_bm25s_scores.py:12-30     precompute_score_for_field

Pinned by ``tests/golden/precompute/*.npz``, produced by EXECUTING the reference's own
``precompute_score_for_field`` (its 64-process pool included) around a fake ``bm25s.BM25`` object that serves the
score vectors (``oracle/make_golden_precompute.py``).
"""
from __future__ import annotations

from typing import Dict, Iterable, Mapping, Tuple

import numpy as np


def get_scores_sparse(dense_scores: np.ndarray, safe_docs) -> Dict[int, np.float32]:
    """index.py:78-84: ``{i: s[i] for i in range(N) if s[i] != 0}`` filtered by ``doc_id in safe_docs`` - ascending
    doc id (dict insertion order)."""
    nz = np.nonzero(dense_scores != 0)[0]
    return {int(i): dense_scores[i] for i in nz if int(i) in safe_docs}


def precompute_score_for_field(score_rows: Mapping[int, np.ndarray], all_candidate_docs: Iterable[int]
                               ) -> Tuple[np.ndarray, np.ndarray]:
    """precompute_bm25s_scores.py:12-30 without the file writes: ``score_rows`` maps query id -> dense fp32 score
    vector in ``train_queries`` order.  Returns (int32 [nnz,2] (qid, doc id), float16 [nnz])."""
    safe = set(int(d) for d in all_candidate_docs)
    output_keys, output_vals = [], []
    for qid, row in score_rows.items():
        sparse = get_scores_sparse(np.asarray(row), safe)
        output_keys.extend((int(qid), int(doc_id)) for doc_id in sparse.keys())
        output_vals.extend(np.float16(score) for score in sparse.values())
    keys = np.array(output_keys, dtype=np.int32).reshape(-1, 2)
    vals = np.array(output_vals, dtype=np.float16)
    return keys, vals
