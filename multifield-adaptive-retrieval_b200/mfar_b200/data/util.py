"""MemoryMapDict - the "temp-dir doc vectors" store.

Same on-disk format and interface as mfar/data/util.py:28-59: a HEADERLESS raw
``np.memmap(path, float32, shape=(N, d))`` (despite the ``.npy`` suffix the reference gives
the file, mfar/modeling/util.py:85), addressed by document key.  File size is exactly N*d*4.
"""
from __future__ import annotations

from typing import Iterable, Iterator, MutableMapping, Tuple

import numpy as np


class MemoryMapDict(MutableMapping):

    def __init__(self, path: str, keys: Iterable[str], shape: Tuple[int, ...], mode: str = "r+",
                 dtype=np.float32):
        self._keys = {key: i for i, key in enumerate(keys)}
        self._path = path
        self._shape = tuple(shape)
        self._dtype = dtype
        self.file = np.memmap(path, dtype=dtype, mode=mode, shape=self._shape)

    def __getitem__(self, key: str) -> np.ndarray:
        return self.file[self._keys[key], :]

    def __setitem__(self, key: str, value: np.ndarray) -> None:
        self.file[self._keys[key], :] = value

    def __delitem__(self, key: str) -> None:
        raise NotImplementedError

    def __iter__(self) -> Iterator[str]:
        return iter(self._keys)

    def __len__(self) -> int:
        return self._shape[0]

    def __contains__(self, item) -> bool:
        return item in self._keys

    def row_of(self, key: str) -> int:
        return self._keys[key]

    def close(self) -> None:
        self.file.flush()

    def reopen(self) -> None:
        self.file = np.memmap(self._path, dtype=self._dtype, mode="r+", shape=self._shape)
