"""QRes line format written by the retrieval step (mfar/data/trec.py:35-59)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class QRes:
    query_id: str
    doc_id: str
    sim: float
    run_id: str = "0"
    _iter: str = "0"
    _rank: int = 0

    def __str__(self) -> str:   # "qid\t0\tdoc\t0\tsim\t0", trec.py:49-50
        return f"{self.query_id}\t{self._iter}\t{self.doc_id}\t{self._rank}\t{self.sim}\t{self.run_id}"

    @classmethod
    def from_str(cls, s: str) -> "QRes":
        query_id, _iter, doc_id, _rank, sim, run_id = s.split()
        return cls(query_id, doc_id, float(sim), run_id, _iter, int(_rank))
