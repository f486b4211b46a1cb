"""CPU oracle for the BM25 sparse-field scorer (numpy, fp64/fp32 exactly as stated below).

TEST INFRASTRUCTURE ONLY - imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs, never by the product package.

PARITY UNPINNED.  The arithmetic lives in the third-party package ``bm25s`` (pinned 0.1.10 in the reference's
``poetry.lock:768-769`` / ``pyproject.toml:26``).  It is neither vendored under /root/reference nor installed in
this image, and the reference holds no test or golden vector for it, so this file restates bm25s's *published*
algorithm (bm25s 0.1.x ``scoring.py`` / ``__init__.py`` / ``selection.py``) and anchors on the reference's call
sites only:

  mfar/data/index.py:138-140   bm25s.BM25(method="lucene", k1=1.2, b=0.75); index.index(doc_tokens)
  mfar/data/index.py:72-76     index.get_scores(query_tokens)                -> fp32 [N]
  mfar/data/index.py:88-103    index.retrieve(query_tokens, k, backend_selection="numpy")
  mfar/data/index.py:64        bm25s.tokenize(text, stopwords="en", stemmer=...)

Published algorithm restated here
  * df[t]   = number of docs containing token t;  l_d = token count of doc d (after stop-word removal);
    l_avg = mean_d l_d (float64)
  * idf[t]  = ln(1 + (N - df + 0.5) / (df + 0.5))                 float64 math, STORED in a float32 array
  * tfc     = tf / (k1 * (1 - b + b * l_d / l_avg) + tf)          float64  ("lucene" == "robertson" tf part:
                                                                   no (k1 + 1) numerator)
  * score(t, d) = float32( float64(idf32[t]) * tfc )              stored in a float32 CSC matrix
    (column = token, rows = docs in ascending doc order)
  * get_scores(query) : zeros float32 [N]; for every query token that is in the vocabulary, IN QUERY ORDER AND
    WITH REPEATS, ``np.add.at(scores, indices[s:e], data[s:e])`` - float32 accumulation
  * retrieve(k)       : argpartition top-k, sorted descending (tie order unspecified); k > N raises ValueError
  * tokenize          : lower-case, regex ``(?u)\\b\\w\\w+\\b``, drop the 33 Lucene English stop words, optional
                        stemmer applied per unique token
"""
from __future__ import annotations

import math
import re
from collections import Counter
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

# bm25s.tokenization.STOPWORDS_EN (0.1.x): Lucene's default English stop-word set
STOPWORDS_EN = ("a", "an", "and", "are", "as", "at", "be", "but", "by", "for", "if", "in", "into", "is", "it", "no",
                "not", "of", "on", "or", "such", "that", "the", "their", "then", "there", "these", "they", "this",
                "to", "was", "will", "with")
_TOKEN_RE = re.compile(r"(?u)\b\w\w+\b")


def tokenize(text: str, stopwords: Sequence[str] = STOPWORDS_EN,
             stemmer: Optional[Callable[[str], str]] = None) -> List[str]:
    """bm25s.tokenize(text, stopwords="en", stemmer=...)[0] with return_ids=False (index.py:64)."""
    stop = set(stopwords)
    toks = [t for t in _TOKEN_RE.findall(text.lower()) if t not in stop]
    if stemmer is not None:
        toks = [stemmer(t) for t in toks]
    return toks


def build_index(corpus_token_ids: Sequence[Sequence[int]], n_vocab: int, k1: float = 1.2, b: float = 0.75
                ) -> Dict[str, np.ndarray]:
    """``BM25(method="lucene", k1, b).index(tokens)`` (index.py:138-140): the float32 CSC score matrix."""
    n_docs = len(corpus_token_ids)
    doc_len = np.array([len(d) for d in corpus_token_ids], dtype=np.float64)
    l_avg = float(doc_len.mean()) if n_docs else 0.0
    df = np.zeros(n_vocab, dtype=np.int64)
    for d in corpus_token_ids:
        for t in set(d):
            df[t] += 1
    idf = np.zeros(n_vocab, dtype=np.float32)
    for t in range(n_vocab):
        idf[t] = math.log(1.0 + (n_docs - df[t] + 0.5) / (df[t] + 0.5))
    cols: List[List[Tuple[int, np.float32]]] = [[] for _ in range(n_vocab)]
    for doc, toks in enumerate(corpus_token_ids):
        cnt = Counter(toks)
        voc = np.array(list(cnt.keys()), dtype=np.int64)
        tf = np.array(list(cnt.values()), dtype=np.float64)
        tfc = tf / (k1 * ((1.0 - b) + b * len(toks) / l_avg) + tf)
        sc = (idf[voc].astype(np.float64) * tfc).astype(np.float32)
        for t, s in zip(voc.tolist(), sc):
            cols[t].append((doc, s))
    indptr = np.zeros(n_vocab + 1, dtype=np.int64)
    for t in range(n_vocab):
        indptr[t + 1] = indptr[t] + len(cols[t])
    indices = np.zeros(int(indptr[-1]), dtype=np.int32)
    data = np.zeros(int(indptr[-1]), dtype=np.float32)
    for t in range(n_vocab):
        for i, (doc, s) in enumerate(cols[t]):
            indices[indptr[t] + i] = doc
            data[indptr[t] + i] = s
    return {"data": data, "indices": indices, "indptr": indptr, "num_docs": n_docs, "n_vocab": n_vocab,
            "l_avg": l_avg, "k1": k1, "b": b}


def get_scores(index: Dict[str, np.ndarray], query_token_ids: Sequence[int]) -> np.ndarray:
    """``BM25.get_scores`` (called at index.py:75): sequential float32 add.at per query token, repeats included.
    Token ids outside the vocabulary stand for tokens that are not in ``vocab_dict`` and are skipped."""
    scores = np.zeros(index["num_docs"], dtype=np.float32)
    indptr, indices, data = index["indptr"], index["indices"], index["data"]
    for t in query_token_ids:
        if 0 <= t < index["n_vocab"]:
            s, e = int(indptr[t]), int(indptr[t + 1])
            np.add.at(scores, indices[s:e], data[s:e])
    return scores


def retrieve(index: Dict[str, np.ndarray], queries_token_ids: Sequence[Sequence[int]], k: int
             ) -> Tuple[np.ndarray, np.ndarray]:
    """``BM25.retrieve(..., backend_selection="numpy")`` (index.py:92,99): (doc rows [Q,k], scores [Q,k]),
    sorted descending; ties broken by ascending doc row here (unspecified upstream)."""
    if k > index["num_docs"]:
        raise ValueError("k of the top-k retrieval exceeds the number of documents")
    rows, vals = [], []
    for q in queries_token_ids:
        s = get_scores(index, q)
        order = np.lexsort((np.arange(len(s)), -s.astype(np.float64)))[:k]
        rows.append(order)
        vals.append(s[order])
    return np.stack(rows), np.stack(vals)


def shard(index: Dict[str, np.ndarray], lo: int, hi: int) -> Dict[str, np.ndarray]:
    """Postings of docs [lo, hi) only, rows rebased to lo - corpus statistics (idf, l_avg) stay global."""
    keep = (index["indices"] >= lo) & (index["indices"] < hi)
    c = np.concatenate([[0], np.cumsum(keep.astype(np.int64))])
    per_tok = c[index["indptr"][1:]] - c[index["indptr"][:-1]]
    indptr = np.concatenate([[0], np.cumsum(per_tok)]).astype(np.int64)
    out = dict(index)
    out.update(data=index["data"][keep], indices=(index["indices"][keep] - lo).astype(np.int32), indptr=indptr,
               num_docs=hi - lo)
    return out
