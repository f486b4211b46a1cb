"""Producer of the precomputed per-field BM25 score files (mfar/commands/precompute_bm25s_scores.py).

``precompute_score_for_field`` keeps the reference's signature and writes the same two files per field
(``{output_path}/{field_key}_keys_bm25.npy`` int32 [nnz,2] = (query id, doc id) and ``{field_key}_vals_bm25.npy``
float16 [nnz], precompute_bm25s_scores.py:26-30) in the same order (queries in dict order, docs ascending) - the files
``PrecomputedSparseScores.load`` / the reference's ``read_sparse_scores`` (mfar/modeling/util.py:151-173) read back.
Where the reference maps ``get_scores_sparse`` over a 64-process pool and walks Python dicts
(precompute_bm25s_scores.py:17-24), here a batch of queries is scored against the HBM-resident postings and the rows
are filtered + compacted on the device (``BM25sSparseIndex.get_scores_sparse_batch``).
"""
from __future__ import annotations

from typing import Dict, Iterable, Mapping, Optional, Sequence, Tuple

import numpy as np

from ..data.index import BM25sSparseIndex
from ..data.schema import resolve_fields
from ..data.typedef import FieldType


def precompute_score_for_field(index: BM25sSparseIndex, all_candidate_docs: Iterable[int],
                               train_queries: Mapping[int, str], output_path: Optional[str], field_key: str,
                               batch_size: int = 256) -> Tuple[np.ndarray, np.ndarray]:
    """precompute_bm25s_scores.py:12-30.  Returns the (keys, vals) arrays it saved (``output_path=None``: only
    returns them)."""
    index.set_safe_docs(all_candidate_docs)
    qids = list(train_queries.keys())
    texts = list(train_queries.values())
    keys, vals = [np.zeros((0, 2), np.int32)], [np.zeros((0,), np.float16)]
    for b in range(0, len(qids), batch_size):
        k, v = index.get_scores_sparse_batch(texts[b:b + batch_size], [int(q) for q in qids[b:b + batch_size]])
        keys.append(k)
        vals.append(v)
    output_keys_array = np.concatenate(keys).astype(np.int32, copy=False)
    output_vals_array = np.concatenate(vals).astype(np.float16, copy=False)
    if output_path is not None:
        np.save(f"{output_path}/{field_key}_keys_bm25.npy", output_keys_array)
        np.save(f"{output_path}/{field_key}_vals_bm25.npy", output_vals_array)
    return output_keys_array, output_vals_array


def read_queries_and_positives(data_path: str, partition: str = "train") -> Tuple[Dict[int, str], set]:
    """``{partition}.queries`` (id \\t text) and the doc ids of ``{partition}.qrels`` (qid \\t _ \\t doc \\t _),
    precompute_bm25s_scores.py:55-68."""
    queries: Dict[int, str] = {}
    with open(f"{data_path}/{partition}.queries", "r") as f:
        for line in f:
            idx, query = line.strip().split("\t")
            queries[int(idx)] = query
    pos_docs = set()
    with open(f"{data_path}/{partition}.qrels", "r") as f:
        for line in f:
            _, _, doc_id, _ = line.strip().split("\t")
            pos_docs.add(int(doc_id))
    return queries, pos_docs


def candidate_docs(negative_sampling_index: BM25sSparseIndex, queries: Sequence[str], pos_docs: set,
                   top_k: int = 150, batch_size: int = 256) -> set:
    """Top-150 BM25 docs of every train query (possible negatives) united with the positives
    (precompute_bm25s_scores.py:73-82; doc keys are the integer doc ids)."""
    cand = set(pos_docs)
    for b in range(0, len(queries), batch_size):
        for hits in negative_sampling_index.retrieve_batch(queries[b:b + batch_size], top_k=top_k):
            cand.update(int(doc_id) for doc_id, _ in hits)
    return cand


def main(data_path: str, dataset_name: str, output_path: str, index_path: str,
         fields_str: str = "all_sparse,single_sparse", device: str = "cuda", batch_size: int = 256) -> None:
    """precompute_bm25s_scores.py:32-88 over saved index directories: every sparse field's
    ``{index_path}/{field_key}_sparse_index`` (what create_bm25s_index.py:23-24 writes; the reference rebuilds the
    indices from the corpus text here, which needs its document formatter - outside this path)."""
    fields = resolve_fields(fields_str, dataset_name)
    if any(field.field_type == FieldType.DENSE for field in fields.values()):
        raise ValueError("Dense fields are not supported in this script.")        # precompute_bm25s_scores.py:42-43
    train_queries, pos_docs = read_queries_and_positives(data_path, "train")
    neg_index = BM25sSparseIndex.load(f"{index_path}/single_sparse_sparse_index", device=device)
    all_candidate_docs = candidate_docs(neg_index, list(train_queries.values()), pos_docs, batch_size=batch_size)
    del neg_index
    for field_key in fields:
        index = BM25sSparseIndex.load(f"{index_path}/{field_key}_sparse_index", device=device)
        k, _ = precompute_score_for_field(index, all_candidate_docs, train_queries, output_path, field_key, batch_size)
        print(f"{len(k)} scores written to {output_path}/{field_key}.scores")
        del index
