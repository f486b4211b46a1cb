"""Wiring of the vector store to the indices (mfar/modeling/util.py:73-108)."""
from __future__ import annotations

import csv
import json
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Tuple

from ..data.index import BM25sSparseIndex, DenseFlatIndex, PrecomputedSparseIndex
from ..data.typedef import Field, FieldType
from ..data.util import MemoryMapDict


def read_corpus(path: str) -> Iterable[Tuple[str, object]]:
    """TREC corpus TSV: ``doc_id \\t json`` per line (mfar/data/trec.py:96-105)."""
    with open(path, "r") as f:
        for row in csv.reader(f, delimiter="\t"):
            if len(row) < 2:
                yield row[0], ""
            else:
                try:
                    yield row[0], json.loads(row[1])
                except Exception:
                    yield row[0], row[1]


class FieldSparseScores:
    """One sparse field's precomputed pairs, sorted by (query id, doc row).  Stands where the reference keeps a
    ``{qid: {doc_id: score}}`` dict (mfar/modeling/util.py:111-148): ``qid in f``, ``f.get(qid, {})`` behave like that
    dict; ``score_matrix`` is the vectorised form of ``score_batch_with_cache`` (index.py:120-125)."""

    def __init__(self, keys, vals):
        import numpy as np
        keys = np.asarray(keys, dtype=np.int32).reshape(-1, 2)
        vals = np.asarray(vals)
        assert len(keys) == len(vals)                                  # modeling/util.py:163
        order = np.lexsort((keys[:, 1], keys[:, 0]))                   # stable: of duplicate pairs the last one stays last
        self.qid = np.ascontiguousarray(keys[order, 0])
        self.doc = np.ascontiguousarray(keys[order, 1])
        self.val = np.ascontiguousarray(vals[order])

    def _span(self, qid):
        import numpy as np
        return int(np.searchsorted(self.qid, qid, "left")), int(np.searchsorted(self.qid, qid, "right"))

    def __contains__(self, qid) -> bool:
        lo, hi = self._span(qid)
        return hi > lo

    def __len__(self) -> int:
        return len(self.val)

    def get(self, qid, default=None):
        lo, hi = self._span(qid)
        if hi == lo:
            return default
        return dict(zip(self.doc[lo:hi].tolist(), self.val[lo:hi].tolist()))     # later duplicates overwrite earlier

    def score_matrix(self, query_ids, doc_rows):
        """fp32 [len(query_ids), len(doc_rows)]: the stored score of every (query, doc) pair, 0 where absent."""
        import numpy as np
        doc_rows = np.asarray(doc_rows, dtype=np.int64)
        out = np.zeros((len(query_ids), len(doc_rows)), dtype=np.float32)
        for r, qid in enumerate(query_ids):
            lo, hi = self._span(qid)
            if hi == lo:
                continue
            d = self.doc[lo:hi]
            pos = np.searchsorted(d, doc_rows, "right") - 1                          # last duplicate wins
            hit = (pos >= 0) & (d[np.maximum(pos, 0)] == doc_rows)
            out[r, hit] = self.val[lo:hi][pos[hit]].astype(np.float32)
        return out


class PrecomputedSparseScores:
    """Precomputed per-field BM25 scores in the reference's file layout
    (``{scores_path}/{field_key}_keys_bm25.npy`` int32 [nnz,2] = (query id, doc row) and
    ``{field_key}_vals_bm25.npy`` float16 [nnz]; written by mfar/commands/precompute_bm25s_scores.py:21-30, read by
    ``read_sparse_scores``, mfar/modeling/util.py:151-173).  Instead of the reference's nested Python dicts the pairs
    stay as sorted arrays; the object still behaves like the reference's ``{field: {qid: {doc: score}}}`` mapping where
    the losses use it (``.values()``, ``[field]``, ``qid in by_field``), ``batch(query_ids)`` slices out the pairs of
    one query batch in the COO form ``MultiFieldRetriever.search(sparse_coo=...)`` consumes."""

    def __init__(self, per_field: Dict[str, Tuple["np.ndarray", "np.ndarray"]], field_keys: List[str]):
        self.field_keys = list(field_keys)
        self._f = {fk: FieldSparseScores(*per_field[fk]) for fk in self.field_keys}

    @classmethod
    def load(cls, scores_path: str, field_info: Dict[str, Field]) -> "PrecomputedSparseScores":
        import numpy as np
        sparse = [k for k, f in field_info.items() if f.field_type == FieldType.SPARSE]
        return cls({k: (np.load(f"{scores_path}/{k}_keys_bm25.npy"), np.load(f"{scores_path}/{k}_vals_bm25.npy"))
                    for k in sparse}, sparse)

    # mapping protocol of the reference's return value
    def __getitem__(self, field_key: str) -> FieldSparseScores:
        return self._f[field_key]

    def __contains__(self, field_key) -> bool:
        return field_key in self._f

    def __len__(self) -> int:
        return len(self._f)

    def __bool__(self) -> bool:
        return len(self._f) > 0

    def keys(self):
        return list(self.field_keys)

    def values(self):
        return [self._f[k] for k in self.field_keys]

    def items(self):
        return [(k, self._f[k]) for k in self.field_keys]

    def lookup(self, field_key: str, qid: int, doc_row: int) -> float:
        """``sparse_scores[field].get(qid, {}).get(doc, 0)`` of the reference (index.py:120-125)."""
        return float(self._f[field_key].score_matrix([qid], [doc_row])[0, 0])

    def batch(self, query_ids, device="cuda"):
        """-> (keys int32 [nnz,2] (row in batch, doc row), vals [nnz], field_offsets [Fs+1]) on ``device``."""
        import numpy as np
        import torch
        ks, vs, offs = [], [], [0]
        for fk in self.field_keys:
            f = self._f[fk]
            for row, qid in enumerate(query_ids):
                lo, hi = f._span(qid)
                if hi > lo:
                    ks.append(np.stack([np.full(hi - lo, row, np.int32), f.doc[lo:hi]], axis=1))
                    vs.append(f.val[lo:hi])
            offs.append(sum(len(x) for x in vs))
        keys = np.concatenate(ks) if ks else np.zeros((0, 2), np.int32)
        vals = np.concatenate(vs) if vs else np.zeros((0,), np.float16)
        return torch.from_numpy(keys).to(device), torch.from_numpy(vals).to(device), offs


def read_sparse_scores(scores_path: str, field_info: Dict[str, Field]) -> PrecomputedSparseScores:
    """``read_sparse_scores`` (mfar/modeling/util.py:151-173): the training-time loader of the precomputed BM25 scores.
    Returns an object usable wherever the reference passes its nested dicts (``HybridContrastiveLoss`` /
    ``score_batch_with_cache``) and as the COO source of the retrieval path."""
    return PrecomputedSparseScores.load(scores_path, field_info)


def read_and_create_indices(corpus_path: str, dataset_name: str, field_info: Dict[str, Field], temp_dir: str,
                            encoder, device="cuda", sparse_scores: Optional[Dict[str, Dict[str, object]]] = None,
                            sparse_texts: Optional[Dict[str, Dict[str, str]]] = None):
    """Same contract as the reference (modeling/util.py:73-108): returns
    ``(corpus, vectors_dict, indices_dict)``; for every dense field an (empty, to-be-filled)
    headerless fp32 memmap ``{temp_dir}/{field.name}.npy`` wrapped in MemoryMapDict + DenseFlatIndex.
    Sparse fields: with ``sparse_texts[field_key] = {doc key: formatted field text}`` (what the reference's
    ``format_documents`` yields, modeling/util.py:102-103 - text formatting is outside this path) a device-resident
    ``BM25sSparseIndex`` is built (modeling/util.py:104-105); otherwise a PrecomputedSparseIndex over
    ``sparse_scores[field_key]``."""
    corpus = list(read_corpus(corpus_path))
    keys: List[str] = [x[0] for x in corpus]
    key_to_row = {k: i for i, k in enumerate(keys)}
    vectors_dict, indices_dict = {}, {}
    for field_key, field in field_info.items():
        if field.field_type == FieldType.DENSE:
            dim = encoder.get_sentence_embedding_dimension()           # sparse-only callers pass encoder=None
            v_file = f"{temp_dir}/{field.name}.npy"                    # name, not key (modeling/util.py:85)
            Path(v_file).parents[0].mkdir(parents=True, exist_ok=True)
            with open(v_file, "w"):
                pass                                                   # truncate (modeling/util.py:88-89)
            vectors = MemoryMapDict(v_file, keys=keys, shape=(len(corpus), dim))
            vectors_dict[field_key] = vectors
            indices_dict[field_key] = DenseFlatIndex(encoder, vectors.file, numeric_ids_to_keys=keys,
                                                     keys_to_numeric_ids=key_to_row, device=device)
        elif field.field_type == FieldType.SPARSE:
            if sparse_texts is not None and field_key in sparse_texts:
                idx = BM25sSparseIndex.create(sparse_texts[field_key], dataset_name=dataset_name, device=device)
            else:
                idx = PrecomputedSparseIndex(keys, (sparse_scores or {}).get(field_key, {}), device=device)
            idx.name = field.name
            indices_dict[field_key] = idx
    return corpus, vectors_dict, indices_dict
