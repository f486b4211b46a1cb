"""Per-field doc-vector store: the reference's "temp-dir doc vectors" (``MemoryMapDict``, mfar/data/util.py:28-59).

On disk it is the reference's format, byte for byte: a HEADERLESS raw fp32 matrix ``[N, d]`` (despite the ``.npy``
suffix the reference gives the file, mfar/modeling/util.py:85) of exactly ``N*d*4`` bytes, row i = i-th corpus doc.
In this package the store is the SOURCE the HBM-resident ``PackedCorpus`` is packed from, so besides the mapping
protocol ``trec_eval_step``'s callers use (``store[key]``, ``store[key] = vec``, ``.file``, ``.close()``,
``.reopen()``, contrastive.py:488-494) it hands out row numbers and pinned, chunked row blocks for the H2D + pack
kernel path (``rows_of``, ``iter_row_blocks``, ``pack_into``).
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np


class MemoryMapDict:
    """``MemoryMapDict(path, keys, shape, mode="r+", dtype=np.float32)`` - same constructor as the reference."""

    def __init__(self, path: str, keys: Iterable[str], shape: Tuple[int, ...], mode: str = "r+", dtype=np.float32):
        self.path = path
        self.doc_keys: List[str] = list(keys)
        self.n_rows, self.dim = int(shape[0]), int(shape[1])
        self.dtype = np.dtype(dtype)
        self._row_index: Optional[Dict[str, int]] = None     # built on first keyed access
        self.file = self._map(mode)

    def _map(self, mode: str) -> np.memmap:
        return np.memmap(self.path, dtype=self.dtype, mode=mode, shape=(self.n_rows, self.dim))

    # ------------------------------------------------------------------ key -> row
    def row_of(self, key: str) -> int:
        if self._row_index is None:
            self._row_index = {k: i for i, k in enumerate(self.doc_keys)}
        return self._row_index[key]                           # KeyError for an unknown doc key, as in the reference

    def rows_of(self, keys: Sequence[str]) -> np.ndarray:
        return np.fromiter((self.row_of(k) for k in keys), dtype=np.int64, count=len(keys))

    # ------------------------------------------------------------------ mapping protocol of the reference class
    def __getitem__(self, key: str) -> np.ndarray:
        return self.file[self.row_of(key)]

    def __setitem__(self, key: str, vector) -> None:
        self.file[self.row_of(key)] = vector

    def __contains__(self, key) -> bool:
        try:
            self.row_of(key)
            return True
        except KeyError:
            return False

    def __delitem__(self, key: str) -> None:
        raise NotImplementedError("rows of a doc-vector store cannot be removed")   # the reference raises the same

    def __iter__(self) -> Iterator[str]:
        return iter(self.doc_keys)

    def __len__(self) -> int:
        return self.n_rows

    def keys(self):
        return list(self.doc_keys)

    def close(self) -> None:
        """Write dirty pages back (the reference calls this after the encode pass, contrastive.py:492)."""
        self.file.flush()

    def reopen(self) -> None:
        """Map the file again (the reference rebinds ``index.vectors`` to the new map, contrastive.py:493-494)."""
        self.file = self._map("r+")

    # ------------------------------------------------------------------ feeding the HBM-resident corpus
    def iter_row_blocks(self, block_rows: int = 131072) -> Iterator[Tuple[int, np.ndarray]]:
        """(first row, contiguous fp32 block) pairs covering the file in order."""
        for lo in range(0, self.n_rows, block_rows):
            yield lo, np.ascontiguousarray(self.file[lo:min(self.n_rows, lo + block_rows)])

    def pack_into(self, corpus, field: int, block_rows: int = 131072) -> None:
        """Stream this field into column ``field`` of a ``PackedCorpus`` (H2D + ``mfar_corpus_pack_rows`` per block)."""
        import torch
        for lo, block in self.iter_row_blocks(block_rows):
            corpus.load_rows(field, lo, torch.from_numpy(block))
