"""Hard-negative mining over an ``Index`` (mfar/data/negative_sampler.py) - the dataloader-side caller of the
retrieval path (``Index.retrieve``, negative_sampler.py:43).

Same classes, constructor arguments and sampling rule as the reference: retrieve ``n_retrieve`` docs, drop the
query's positives, keep the ``n_bottom`` lowest-scoring survivors, draw ``n_sample`` of them with ``random.sample``.
``sample_batch`` - a per-query loop in the reference ("TODO: implement batch sampling", negative_sampler.py:61-63) -
issues ONE ``retrieve_batch`` for the whole batch (one fused GPU pass) and consumes the ``random`` stream in the same
per-query order, so it returns what the reference's loop returns for the same seed.
"""
from __future__ import annotations

import random
from abc import ABC, abstractmethod
from typing import AbstractSet, List, Mapping, Sequence, Tuple

from .index import Index
from .typedef import Document, Query


Positives = Mapping[str, AbstractSet[str]]        # query id -> keys of its relevant docs


class NegativeSampler(ABC):
    """Interface of negative_sampler.py:10-20: ``n_sample`` docs per query, drawn for one query or a batch."""

    n_sample: int

    @abstractmethod
    def sample(self, query: Query, pos_for_each_qid: Positives) -> List[Document]: ...

    def sample_batch(self, queries: List[Query], pos_for_each_qid: Positives) -> List[List[Document]]:
        return [self.sample(q, pos_for_each_qid) for q in queries]


class IndexNegativeSampler(NegativeSampler):
    """negative_sampler.py:22-63.  ``documents`` maps doc key -> text (missing keys give an empty text)."""

    def __init__(self, index: Index, documents: Mapping[str, str], n_retrieve: int = 50, n_bottom: int = 5,
                 n_sample: int = 1):
        self.index, self.documents = index, documents
        self.n_retrieve, self.n_bottom, self._n_sample = n_retrieve, n_bottom, n_sample

    @property
    def n_sample(self) -> int:
        return self._n_sample

    def _retrieve_negatives(self, text: str, top_k: int, positives: AbstractSet[str],
                            hits: Sequence[Tuple[str, float]] = None) -> List[Tuple[str, float]]:
        """Retrieved (key, score) pairs minus the query's positives (negative_sampler.py:41-45)."""
        if hits is None:
            hits = self.index.retrieve(text, top_k=top_k)
        return [(key, score) for key, score in hits if key not in positives]

    def _pick(self, query: Query, positives: AbstractSet[str], first_hits=None) -> List[Document]:
        cands = self._retrieve_negatives(query.text, self.n_retrieve, positives, first_hits)
        if not cands:                    # everything retrieved was relevant: look deeper once (negative_sampler.py:46-52)
            cands = self._retrieve_negatives(query.text, len(positives) + self.n_bottom, positives)
        cands.sort(key=lambda pair: pair[1], reverse=True)               # stable, like the reference's list.sort
        bottom = [key for key, _ in cands[-self.n_bottom:]]              # the n_bottom lowest-scoring survivors
        chosen = random.sample(range(len(bottom)), self.n_sample)        # same draw as negative_sampler.py:55
        return [Document(bottom[i], self.documents.get(bottom[i], "")) for i in chosen]

    def sample(self, query: Query, pos_for_each_qid: Positives) -> List[Document]:
        return self._pick(query, pos_for_each_qid[query._id])

    def sample_batch(self, queries: List[Query], pos_for_each_qid: Positives) -> List[List[Document]]:
        """ONE ``retrieve_batch`` (one fused GPU pass) for the whole batch; the per-query filtering, the rare deeper
        retry and the ``random`` draws then run in query order, so the result equals the reference's loop."""
        all_hits = self.index.retrieve_batch([q.text for q in queries], top_k=self.n_retrieve)
        return [self._pick(q, pos_for_each_qid[q._id], hits) for q, hits in zip(queries, all_hits)]
