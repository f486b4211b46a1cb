"""Schema and text-container types of the scoring path (mfar/data/typedef.py).

``Field`` / ``FieldType`` (typedef.py:69-122) order the scorer columns; ``Query`` / ``Document`` / ``Corpus``
(typedef.py:13-67, 125-172) are the plain containers the callers either side of the path pass around
(``IndexNegativeSampler``, ``BM25sSparseIndex.create``) - kept without the reference's JSON mixin, gzip readers and
STaRK text formatting, which are data preparation.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from enum import Enum
from typing import Any, Dict, Iterator, List, Optional


@dataclass
class Query:                     # typedef.py:13-17
    _id: str
    text: str
    metadata: Any = None


@dataclass
class Document:                  # typedef.py:31-36
    _id: str
    text: str
    title: Optional[str] = None
    metadata: Any = None


@dataclass
class Corpus:                    # typedef.py:125-172
    docs: List[Document]
    dataset_name: Optional[str] = None

    def __post_init__(self):
        self.key_to_id = {doc._id: i for i, doc in enumerate(self.docs)}

    def keys(self) -> Iterator[str]:
        return (doc._id for doc in self.docs)

    def __len__(self) -> int:
        return len(self.docs)

    def get_text_by_id(self, doc_id: int) -> str:
        return self.docs[doc_id].text

    def get_text_by_key(self, key: str) -> str:
        return self.docs[self.key_to_id[key]].text

    def get_doc_by_id(self, doc_id: int) -> Document:
        return self.docs[doc_id]

    def get_doc_by_key(self, key: str) -> Document:
        try:
            return self.docs[self.key_to_id[key]]
        except KeyError:
            raise KeyError(f"Key {key} not found in corpus.")

    def pairs(self):
        return ((doc._id, doc.text) for doc in self.docs)

    @classmethod
    def from_docs_dict(cls, docs_dict: Dict[Any, str], dataset_name: Optional[str] = None) -> "Corpus":
        return cls([Document(key, text) for key, text in docs_dict.items()], dataset_name)


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
