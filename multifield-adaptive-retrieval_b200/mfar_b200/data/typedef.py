"""Field / FieldType - the two schema types the scoring path needs.

Mirrors mfar/data/typedef.py:69-122 (the reference's Query/Document/Corpus text containers
are outside the hot path and not rebuilt).
"""
from __future__ import annotations

import json
from enum import Enum
from typing import Optional


class FieldType(Enum):
    SPARSE = 1
    DENSE = 2


class Field:
    """One scorer column: ``key`` (e.g. "title_dense"), ``name`` (e.g. "title", which also
    names the vector file ``{temp_dir}/{name}.npy``), its type and token budget."""

    __slots__ = ("key", "name", "field_type", "max_seq_length", "dataset")

    def __init__(self, key: str, name: str, field_type: FieldType, max_seq_length: int = 512,
                 dataset: Optional[str] = None):
        self.key = key
        self.name = name
        self.field_type = field_type
        self.max_seq_length = max_seq_length
        self.dataset = dataset

    def serialize(self) -> dict:
        return {"key": self.key, "name": self.name, "field_type": self.field_type.name,
                "max_seq_length": self.max_seq_length, "dataset": self.dataset}

    @classmethod
    def deserialize(cls, data: dict) -> "Field":
        return cls(data["key"], data["name"], FieldType[data["field_type"]], data["max_seq_length"],
                   data["dataset"])

    def __repr__(self) -> str:
        return json.dumps({"name": self.name, "field_type": self.field_type.name,
                           "max_seq_length": self.max_seq_length})

    def __copy__(self):
        return Field(self.key, self.name, self.field_type, self.max_seq_length, self.dataset)

    def __deepcopy__(self, memo):
        return self.__copy__()
