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


class PrecomputedSparseScores:
    """Precomputed per-field BM25 scores in the reference's file layout
    (``{scores_path}/{field_key}_keys_bm25.npy`` int32 [nnz,2] = (query id, doc row) and
    ``{field_key}_vals_bm25.npy`` float16 [nnz]; written by mfar/commands/precompute_bm25s_scores.py:21-30, read by
    ``read_sparse_scores``, mfar/modeling/util.py:151-173).  Instead of the reference's nested Python dicts the
    pairs stay as arrays sorted by query id; ``batch(query_ids)`` slices out the pairs of one query batch in the
    COO form ``MultiFieldRetriever.search(sparse_coo=...)`` consumes."""

    def __init__(self, per_field: Dict[str, Tuple["np.ndarray", "np.ndarray"]], field_keys: List[str]):
        import numpy as np
        self.field_keys = list(field_keys)
        self._f = {}
        for fk in self.field_keys:
            keys, vals = per_field[fk]
            keys = np.asarray(keys, dtype=np.int32).reshape(-1, 2)
            vals = np.asarray(vals)
            assert len(keys) == len(vals)                              # modeling/util.py:163
            order = np.argsort(keys[:, 0], kind="stable")
            self._f[fk] = (np.ascontiguousarray(keys[order, 0]), np.ascontiguousarray(keys[order, 1]),
                           np.ascontiguousarray(vals[order]))

    @classmethod
    def load(cls, scores_path: str, field_info: Dict[str, Field]) -> "PrecomputedSparseScores":
        import numpy as np
        sparse = [k for k, f in field_info.items() if f.field_type == FieldType.SPARSE]
        return cls({k: (np.load(f"{scores_path}/{k}_keys_bm25.npy"), np.load(f"{scores_path}/{k}_vals_bm25.npy"))
                    for k in sparse}, sparse)

    def lookup(self, field_key: str, qid: int, doc_row: int) -> float:
        """``sparse_scores[field].get(qid, {}).get(doc, 0)`` of the reference (index.py:120-125)."""
        import numpy as np
        q, d, v = self._f[field_key]
        lo, hi = np.searchsorted(q, qid, "left"), np.searchsorted(q, qid, "right")
        hit = np.nonzero(d[lo:hi] == doc_row)[0]
        return float(v[lo + hit[-1]]) if len(hit) else 0.0             # dict semantics: the last duplicate wins

    def batch(self, query_ids, device="cuda"):
        """-> (keys int32 [nnz,2] (row in batch, doc row), vals [nnz], field_offsets [Fs+1]) on ``device``."""
        import numpy as np
        import torch
        ks, vs, offs = [], [], [0]
        for fk in self.field_keys:
            q, d, v = self._f[fk]
            for row, qid in enumerate(query_ids):
                lo, hi = np.searchsorted(q, qid, "left"), np.searchsorted(q, qid, "right")
                if hi > lo:
                    ks.append(np.stack([np.full(hi - lo, row, np.int32), d[lo:hi]], axis=1))
                    vs.append(v[lo:hi])
            offs.append(sum(len(x) for x in vs))
        keys = np.concatenate(ks) if ks else np.zeros((0, 2), np.int32)
        vals = np.concatenate(vs) if vs else np.zeros((0,), np.float16)
        return torch.from_numpy(keys).to(device), torch.from_numpy(vals).to(device), offs


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
