"""Index API of the scoring path (mfar/data/index.py).

``Index``                  - the ABC (index.py:21-37)
``DenseFlatIndex``         - exhaustive flat dense index over ONE field (index.py:160-232); same
                             constructor and return types, every dot product / top-k on the GPU
``BM25sSparseIndex``       - the reference's sparse index (index.py:39-157) over a device-resident
                             ``DeviceBM25`` (postings in HBM, scatter-add kernel) instead of ``bm25s.BM25``
``PrecomputedSparseIndex`` - sparse index fed with PRECOMPUTED per-query full-corpus score vectors
                             (what ``get_scores`` returns); does the gather / top-k the reference does
                             around them (index.py:95-118)
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Dict, Generic, List, Optional, Sequence, Tuple, TypeVar, Union

import numpy as np
import torch

from .. import _native as nv

Key = TypeVar("Key")
Query = TypeVar("Query")


class Index(ABC, Generic[Key, Query]):
    """Anything that can be searched (index.py:21-37)."""

    @abstractmethod
    def retrieve(self, query: Query, top_k: int) -> Sequence[Tuple[Key, float]]:
        raise NotImplementedError

    def retrieve_batch(self, queries: Sequence[Query], top_k: int) -> Sequence[Sequence[Tuple[Key, float]]]:
        return [self.retrieve(q, top_k) for q in queries]


class DenseFlatIndex(Index[str, str]):
    """Flat exhaustive index over one dense field.

    ``vectors`` is the field's [N, d] fp32 array (the reference passes ``MemoryMapDict.file``,
    mfar/modeling/util.py:96-101).  It is packed to bf16 on the device lazily and re-packed when
    ``.vectors`` is rebound (the reference rebinds it after the corpus is encoded,
    contrastive.py:493-494).  ``device`` must be a CUDA device; ``vector_batch_size`` is accepted
    for signature compatibility (the fused kernel streams the whole field in one pass)."""

    def __init__(self, model, vectors, numeric_ids_to_keys: Sequence[str], keys_to_numeric_ids: Dict[str, int],
                 device: Union[str, torch.device] = "cuda", vector_batch_size: int = 1048576,
                 normalize: bool = False):
        self.model = model
        self._vectors = vectors
        self.numeric_ids_to_key = numeric_ids_to_keys
        self.key_to_numeric_ids = keys_to_numeric_ids
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DenseFlatIndex (mfar_b200) runs on a CUDA sm_100 device only; no CPU path")
        self.vector_batch_size = vector_batch_size
        self.normalize = normalize
        self._retriever = None

    # rebinding .vectors invalidates the packed copy (contrastive.py:493-494)
    @property
    def vectors(self):
        return self._vectors

    @vectors.setter
    def vectors(self, value) -> None:
        self._vectors = value
        self._retriever = None

    def _ensure(self):
        if self._retriever is None:
            from ..modeling.retrieval import MultiFieldRetriever, PackedCorpus
            from ..modeling.weighting import LinearWeights
            pc = PackedCorpus.from_fields([self._vectors], self.device, self.normalize)
            mix = LinearWeights(1, 1).to(self.device)          # single field: weight softmax == 1
            self._retriever = MultiFieldRetriever(pc, mix, top_k=100)
        return self._retriever

    def _encode(self, queries) -> torch.Tensor:
        if isinstance(queries, np.ndarray):                    # index.py:184-185
            return torch.from_numpy(queries)
        if torch.is_tensor(queries):
            return queries
        return self.model.encode(list(queries), convert_to_tensor=True)   # index.py:187

    def retrieve(self, query, top_k: int):
        if isinstance(query, np.ndarray) or torch.is_tensor(query):   # a pre-encoded vector [d] or [1,d]
            return self.retrieve_batch(query.reshape(1, -1), top_k)[0]
        return self.retrieve_batch([query], top_k)[0]                 # index.py:178-179

    def retrieve_batch(self, queries, top_k: int) -> List[List[Tuple[str, float]]]:
        r = self._ensure()
        k_eff = min(int(top_k), r.n_docs)
        s, i = r.per_field_topk(self._encode(queries), None, k_eff, zero_init=True)
        if k_eff < top_k:
            # more hits requested than the field holds: the reference's running top-k starts as top_k entries of
            # (0.0, row 0) (index.py:192-193) and those that no real score displaces stay - after the zero-init rule
            # every kept score is >= 0, so they are exactly the tail
            pad = int(top_k) - k_eff
            s = torch.nn.functional.pad(s, (0, pad), value=0.0)
            i = torch.nn.functional.pad(i, (0, pad), value=0)
        rows, vals = i[0].cpu().tolist(), s[0].cpu().tolist()
        return [list(zip([self.numeric_ids_to_key[j] for j in rows[q]], vals[q])) for q in range(len(rows))]

    def score(self, query, keys: Sequence[str]) -> torch.Tensor:
        return self.score_batch([query], keys)[0]

    def score_batch(self, queries, keys: Sequence[str]) -> torch.Tensor:
        """[Q, C] fp32 scores of the given keys (index.py:227-232); unknown key -> KeyError, as there."""
        r = self._ensure()
        rows = torch.tensor([self.key_to_numeric_ids[k] for k in keys], dtype=torch.int64)
        return r.score_candidates(self._encode(queries), rows)[0]


class BM25sSparseIndex(Index[str, str]):
    """``BM25sSparseIndex`` (index.py:39-157) with the BM25 score matrix in HBM.

    Same constructor, methods and return types; ``index`` is a ``mfar_b200.data.bm25.DeviceBM25`` (stands where
    ``bm25s.BM25`` stands).  ``get_scores`` keeps the reference's per-query cache (index.py:71, ``lru_cache``)."""

    def __init__(self, keys: List[str], index, stemmer=None, index_limit: int = 5000, safe_docs=None):
        self.keys = keys
        self.key_to_id = {key: i for i, key in enumerate(keys)}
        self.index = index
        self.stemmer = stemmer
        self.index_limit = index_limit
        self.safe_docs = safe_docs if safe_docs is not None else {}
        self.name = None
        self._score_cache: Dict[str, np.ndarray] = {}
        self._cache_size = 2 ** 15                                   # index.py:71
        self._safe_bits = None                                       # device bitmap of safe_docs (lazy)

    def set_safe_docs(self, safe_docs):
        self.safe_docs = safe_docs
        self._safe_bits = None

    @staticmethod
    def tokenize_single(query: str, stopwords="en", stemmer=None, return_ids: bool = False) -> List[str]:
        """index.py:55-64 with return_ids=False: the token list of one query string."""
        from .bm25 import tokenize
        return tokenize(query, stopwords, stemmer)[0]

    def tokenize(self, queries, stopwords="en", stemmer=None, return_ids: bool = False):
        """index.py:55-69 with return_ids=False: token list for a str, list of token lists for a sequence."""
        from .bm25 import tokenize
        if isinstance(queries, str):
            return tokenize(queries, stopwords, stemmer)[0]
        return tokenize(list(queries), stopwords, stemmer)

    def get_scores(self, query: str) -> np.ndarray:
        """fp32 [N] BM25 scores of one query against the whole field corpus (index.py:72-76)."""
        hit = self._score_cache.get(query)
        if hit is not None:
            return hit
        tokens = self.tokenize(query, stopwords="en", stemmer=self.stemmer)
        score = self.index.get_scores(tokens) if tokens else np.zeros(self.index.num_docs, dtype=np.float32)
        if len(self._score_cache) >= self._cache_size:
            self._score_cache.pop(next(iter(self._score_cache)))
        self._score_cache[query] = score
        return score

    def get_scores_sparse(self, query: str) -> Dict[int, float]:
        """index.py:78-84: nonzero scores of the docs in ``safe_docs``."""
        dense = self.get_scores(query)
        return {int(i): dense[i] for i in np.nonzero(dense)[0] if int(i) in self.safe_docs}

    def get_scores_sparse_batch(self, queries: Sequence[str], query_ids: Optional[Sequence[int]] = None,
                                vals_dtype=torch.float16) -> Tuple[np.ndarray, np.ndarray]:
        """``get_scores_sparse`` (index.py:78-84) for a whole batch, in the array form the reference's
        ``precompute_score_for_field`` turns the dicts into (precompute_bm25s_scores.py:19-27): keys int32 [nnz,2] =
        (query id, doc id), vals float16 [nnz]; entries != 0 whose doc id is in ``safe_docs``, queries in the given
        order, docs ascending.  The score rows never leave the device: one scatter pass writes them
        (``mfar_bm25_scores``), two passes filter + compact them (``mfar_sparse_coo_count/write``) and only the nnz
        pairs cross PCIe.  ``query_ids`` defaults to the row numbers."""
        from .bm25 import rows_to_coo, safe_docs_bitmap
        if self._safe_bits is None:
            total = self.index.doc_base + self.index.num_docs
            self._safe_bits = torch.from_numpy(safe_docs_bitmap(self.safe_docs, total).view(np.int32)).to(
                self.index.device)
        tokens = self.tokenize(list(queries), stopwords="en", stemmer=self.stemmer)
        sv = self.index.get_scores_batch(tokens)                     # [Q, N] view of the fp32 [Q, 1, ld] rows
        qids = None if query_ids is None else torch.tensor([int(x) for x in query_ids], dtype=torch.int32)
        keys, vals = rows_to_coo(sv, self.index.num_docs, self._safe_bits, qids, self.index.doc_base, vals_dtype)
        return keys.cpu().numpy(), vals.cpu().numpy()

    def retrieve(self, query: str, top_k: int):
        return self.retrieve_batch([query], top_k)[0]                # index.py:86-93

    def retrieve_batch(self, queries: Sequence[str], top_k: int):
        """index.py:95-103: [(key, score)] x top_k per query, best first."""
        tokens = self.tokenize(list(queries), stopwords="en", stemmer=self.stemmer)
        rows, scores = self.index.retrieve(tokens, k=top_k)
        return [[(self.keys[rows[i, j]], scores[i, j]) for j in range(rows.shape[1])] for i in range(rows.shape[0])]

    def score(self, query: str, keys: Sequence[str]) -> np.ndarray:
        doc_ids = np.array([self.key_to_id[key] for key in keys])    # index.py:105-109 (unknown key: KeyError)
        return self.get_scores(query)[doc_ids]

    def score_batch(self, queries: Sequence[str], keys: Sequence[str]) -> torch.Tensor:
        """[Q, C] scores of the given keys; keys missing from the index score 0 (index.py:111-118)."""
        rows = torch.tensor([self.key_to_id.get(k, -1) for k in keys], dtype=torch.int64, device=self.index.device)
        tokens = self.tokenize(list(queries), stopwords="en", stemmer=self.stemmer)
        sv = self.index.get_scores_batch(tokens)                     # [Q, N] on the device
        out = sv[:, rows.clamp(min=0)]
        return (out * (rows >= 0).to(out.dtype)).cpu()

    def score_batch_with_cache(self, query_ids: List[int], keys: Sequence[str], sparse_scores: Dict) -> torch.Tensor:
        """index.py:120-125: lookup in precomputed ``{qid: {doc_id: score}}`` dicts, missing -> 0."""
        doc_ids = [self.key_to_id[key] for key in keys]
        if hasattr(sparse_scores, "score_matrix"):                   # FieldSparseScores: sorted arrays, vectorised
            return torch.from_numpy(sparse_scores.score_matrix(query_ids, doc_ids))
        all_doc_scores = [sparse_scores.get(qid, {}) for qid in query_ids]
        return torch.tensor([[s.get(d, 0) for d in doc_ids] for s in all_doc_scores])

    @classmethod
    def create(cls, corpus, stemmer=None, dataset_name: Optional[str] = "", device="cuda"):
        """index.py:134-145.  ``corpus``: anything with ``keys()`` and ``docs`` (objects with ``.text``), or a
        ``{key: text}`` dict."""
        from .bm25 import DeviceBM25, tokenize
        if isinstance(corpus, dict):
            keys, texts = list(corpus.keys()), list(corpus.values())
        else:
            keys, texts = list(corpus.keys()), [d.text for d in corpus.docs]
        index = DeviceBM25(k1=1.2, b=0.75, method="lucene", device=device)
        index.index(tokenize(texts, stopwords="en", stemmer=stemmer))
        index_limit = 5000 if dataset_name == "amazon" else 12000
        return cls(keys, index, stemmer, index_limit)

    def save(self, path: str):
        import json
        self.index.save(f"{path}/index")
        with open(f"{path}/keys.json", "w") as f:
            json.dump(self.keys, f)

    @classmethod
    def load(cls, path: str, stemmer=None, device="cuda"):
        import json
        from .bm25 import DeviceBM25
        with open(f"{path}/keys.json", "r") as f:
            keys = json.load(f)
        return cls(keys, DeviceBM25.load(f"{path}/index", mmap=True, device=device), stemmer)


class PrecomputedSparseIndex(Index[str, str]):
    """Sparse field index fed with precomputed score vectors.

    ``scores`` maps a query (text or id) to its full-corpus fp32/fp16 score vector [N] - what
    ``BM25sSparseIndex.get_scores`` returns (index.py:72-76)."""

    def __init__(self, keys: List[str], scores: Dict[str, "np.ndarray"], device: Union[str, torch.device] = "cuda"):
        self.keys = keys
        self.key_to_id = {key: i for i, key in enumerate(keys)}
        self.scores = scores
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PrecomputedSparseIndex (mfar_b200) runs on a CUDA sm_100 device only")
        self.name = None

    def get_scores(self, query) -> np.ndarray:
        return self.scores[query]

    def _stack(self, queries) -> torch.Tensor:
        return torch.from_numpy(np.stack([np.asarray(self.scores[q]) for q in queries])).to(self.device)

    def retrieve(self, query, top_k: int):
        return self.retrieve_batch([query], top_k)[0]

    def retrieve_batch(self, queries, top_k: int):
        from ..modeling.retrieval import MultiFieldRetriever
        from ..modeling.weighting import LinearWeights
        sv = self._stack(queries)                                                    # [Q,N]
        r = MultiFieldRetriever(None, LinearWeights(1, 1).to(self.device), n_sparse=1, top_k=top_k,
                                n_docs=sv.shape[1], device=self.device)
        s, i = r.per_field_topk(None, sv.unsqueeze(1), top_k, zero_init=False)
        rows, vals = i[0].cpu().tolist(), s[0].cpu().tolist()
        return [[(self.keys[j], v) for j, v in zip(rows[q], vals[q])] for q in range(len(rows))]

    def score_batch(self, queries, keys: Sequence[str]) -> torch.Tensor:
        """[Q, C]; keys missing from the index score 0 (index.py:112-117)."""
        rows = torch.tensor([self.key_to_id.get(k, -1) for k in keys], dtype=torch.int64, device=self.device)
        sv = self._stack(queries).float()
        out = sv[:, rows.clamp(min=0)]
        return out * (rows >= 0).to(out.dtype)
