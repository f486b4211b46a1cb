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
from abc import ABC
from typing import AbstractSet, List, Mapping, Sequence, Tuple

from .index import Index
from .typedef import Document, Query


class NegativeSampler(ABC):
    @property
    def n_sample(self) -> int:
        raise NotImplementedError

    def sample(self, query: Query, pos_for_each_qid: Mapping[str, AbstractSet[str]]) -> List[Document]:
        raise NotImplementedError

    def sample_batch(self, queries: List[Query], pos_for_each_qid: Mapping[str, AbstractSet[str]]) -> List[List[Document]]:
        raise NotImplementedError


class IndexNegativeSampler(NegativeSampler):
    def __init__(self, index: Index, documents: Mapping[str, str], n_retrieve: int = 50, n_bottom: int = 5,
                 n_sample: int = 1):
        self.index = index
        self.documents = documents
        self.n_retrieve = n_retrieve
        self.n_bottom = n_bottom
        self._n_sample = n_sample

    @property
    def n_sample(self) -> int:
        return self._n_sample

    @staticmethod
    def _negatives(hits: Sequence[Tuple[str, float]], positives: AbstractSet[str]) -> List[Tuple[str, float]]:
        return [(doc_id, score) for doc_id, score in hits if doc_id not in positives]   # negative_sampler.py:41-45

    def _draw(self, cands: List[Tuple[str, float]]) -> List[Document]:
        cands.sort(key=lambda x: x[1], reverse=True)                                    # negative_sampler.py:53 (stable)
        neg_cand_ids = [doc_id for doc_id, _ in cands[-self.n_bottom:]]
        picked = [neg_cand_ids[i] for i in random.sample(range(len(neg_cand_ids)), self.n_sample)]
        return [Document(i, self.documents.get(i, "")) for i in picked]

    def sample(self, query: Query, pos_for_each_qid: Mapping[str, AbstractSet[str]]) -> List[Document]:
        pos = pos_for_each_qid[query._id]
        cands = self._negatives(self.index.retrieve(query.text, top_k=self.n_retrieve), pos)
        if len(cands) == 0:                                                             # negative_sampler.py:46-52
            cands = self._negatives(self.index.retrieve(query.text, top_k=len(pos) + self.n_bottom), pos)
        return self._draw(cands)

    def sample_batch(self, queries: List[Query], pos_for_each_qid: Mapping[str, AbstractSet[str]]) -> List[List[Document]]:
        all_hits = self.index.retrieve_batch([q.text for q in queries], top_k=self.n_retrieve)
        out = []
        for q, hits in zip(queries, all_hits):
            pos = pos_for_each_qid[q._id]
            cands = self._negatives(hits, pos)
            if len(cands) == 0:                         # every retrieved doc was a positive: the deeper retry, per query
                cands = self._negatives(self.index.retrieve(q.text, top_k=len(pos) + self.n_bottom), pos)
            out.append(self._draw(cands))
        return out
