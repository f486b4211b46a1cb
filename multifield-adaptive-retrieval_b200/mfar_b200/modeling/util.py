"""Wiring of the vector store to the indices (mfar/modeling/util.py:73-108)."""
from __future__ import annotations

import csv
import json
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Tuple

from ..data.index import DenseFlatIndex, PrecomputedSparseIndex
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


def read_and_create_indices(corpus_path: str, dataset_name: str, field_info: Dict[str, Field], temp_dir: str,
                            encoder, device="cuda", sparse_scores: Optional[Dict[str, Dict[str, object]]] = None):
    """Same contract as the reference (modeling/util.py:73-108): returns
    ``(corpus, vectors_dict, indices_dict)``; for every dense field an (empty, to-be-filled)
    headerless fp32 memmap ``{temp_dir}/{field.name}.npy`` wrapped in MemoryMapDict + DenseFlatIndex.
    Sparse fields get a PrecomputedSparseIndex over ``sparse_scores[field_key]`` (BM25 itself -
    third-party bm25s in the reference - is an input to this path)."""
    corpus = list(read_corpus(corpus_path))
    keys: List[str] = [x[0] for x in corpus]
    key_to_row = {k: i for i, k in enumerate(keys)}
    vectors_dict, indices_dict = {}, {}
    dim = encoder.get_sentence_embedding_dimension()
    for field_key, field in field_info.items():
        if field.field_type == FieldType.DENSE:
            v_file = f"{temp_dir}/{field.name}.npy"                    # name, not key (modeling/util.py:85)
            Path(v_file).parents[0].mkdir(parents=True, exist_ok=True)
            with open(v_file, "w"):
                pass                                                   # truncate (modeling/util.py:88-89)
            vectors = MemoryMapDict(v_file, keys=keys, shape=(len(corpus), dim))
            vectors_dict[field_key] = vectors
            indices_dict[field_key] = DenseFlatIndex(encoder, vectors.file, numeric_ids_to_keys=keys,
                                                     keys_to_numeric_ids=key_to_row, device=device)
        elif field.field_type == FieldType.SPARSE:
            idx = PrecomputedSparseIndex(keys, (sparse_scores or {}).get(field_key, {}), device=device)
            idx.name = field.name
            indices_dict[field_key] = idx
    return corpus, vectors_dict, indices_dict
