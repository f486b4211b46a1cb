"""Import the reference's own hot-path classes, unmodified, from /root/reference.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py and the optional cross-check test).
/root/reference exists only in the build container, never on the GPU box, so callers must
treat ``available()`` == False as "skip".

Five third-party modules the reference imports at module scope are not installed here
(bm25s, more_itertools, sentence_transformers, mashumaro, pytorch_lightning); none of them
is touched by the code paths we execute, so empty stubs are registered first
(SURVEY.md appendix C).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MFAR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mfar"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


def load():
    """Returns (DenseFlatIndex, BM25sSparseIndex, MemoryMapDict, LinearWeights, resolve_fields)."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True           # the tree is read-only
    # bm25s.tokenize stub: "tokens" of a query are the query string itself, so a fake BM25
    # object can look the precomputed score vector up by query text (index.py:64-65, 74-75).
    _stub("bm25s", BM25=object, tokenize=lambda q, **k: [q])
    _stub("more_itertools", chunked=lambda it, n: None)
    _stub("sentence_transformers", SentenceTransformer=object)
    _stub("mashumaro")
    _stub("mashumaro.mixins")
    _stub("mashumaro.mixins.json", DataClassJSONMixin=object)
    _stub("pytorch_lightning")
    _stub("pytorch_lightning.loggers", MLFlowLogger=object)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from mfar.data.index import DenseFlatIndex, BM25sSparseIndex
    from mfar.data.util import MemoryMapDict
    from mfar.modeling.weighting import LinearWeights
    from mfar.data.schema import resolve_fields
    return DenseFlatIndex, BM25sSparseIndex, MemoryMapDict, LinearWeights, resolve_fields
