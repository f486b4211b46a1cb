"""Golden vectors for the precomputed-BM25 score loader, produced by EXECUTING the reference's own
``read_sparse_scores`` (mfar/modeling/util.py:151-173) and ``BM25sSparseIndex.score_batch_with_cache``
(mfar/data/index.py:120-125) - imported unmodified from /root/reference with stub modules (oracle/ref_import.py plus a
``sentence_transformers.models`` stub for the import line mfar/modeling/util.py:14).

Run in the build container only:   python oracle/make_golden_sparse_scores.py
Writes tests/golden/loader/sparse_scores.npz: the key / value files of two sparse fields (incl. duplicate pairs, for which
the reference's dict keeps the LAST value) and the [Q, C] matrices the reference's cached scorer returns for them.

Note (not mirrored): ``_create_sparse_index_from_npy`` chunks its input with start indices 0, 1, 2, ... instead of
multiples of its CHUNK_SIZE (mfar/modeling/util.py:138-139), so for more than ~2**17 pairs per field it silently drops
the tail; the golden stays below that size.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden", "loader", "sparse_scores.npz")


def main():
    _, BM25sSparseIndex, *_ = ref_import.load()
    m = types.ModuleType("sentence_transformers.models")
    m.Normalize = m.Pooling = object
    sys.modules["sentence_transformers.models"] = m
    sys.modules["sentence_transformers"].__path__ = []
    import mfar.modeling.util as U
    from mfar.data.typedef import Field, FieldType
    rng = np.random.RandomState(7)
    finfo = {"a_dense": Field("a_dense", "a", FieldType.DENSE), "a_sparse": Field("a_sparse", "a", FieldType.SPARSE),
             "b_sparse": Field("b_sparse", "b", FieldType.SPARSE)}
    out = {}
    n_docs = 500
    keys_all = [str(i) for i in range(n_docs)]
    query_ids = [3, 11, 0, 42, 19]                      # 42: a query without any stored score
    cand = [str(x) for x in rng.choice(n_docs, size=40, replace=False)]
    with tempfile.TemporaryDirectory() as tmp:
        for fk in ("a_sparse", "b_sparse"):
            pairs = np.stack([rng.randint(0, 20, size=400), rng.randint(0, n_docs, size=400)], axis=1).astype(np.int32)
            pairs = np.concatenate([pairs, pairs[:25]])                  # duplicates: the later value must win
            vals = rng.gamma(2.0, 2.0, size=len(pairs)).astype(np.float16)
            np.save(os.path.join(tmp, f"{fk}_keys_bm25.npy"), pairs)
            np.save(os.path.join(tmp, f"{fk}_vals_bm25.npy"), vals)
            out[f"{fk}_keys"], out[f"{fk}_vals"] = pairs, vals
        ref = U.read_sparse_scores(tmp, finfo)
        assert sorted(ref) == ["a_sparse", "b_sparse"]
        for fk in ref:
            index = BM25sSparseIndex(keys_all, index=None, stemmer=None)
            out[f"{fk}_cached"] = index.score_batch_with_cache(query_ids, cand, ref[fk]).numpy().astype(np.float32)
            out[f"{fk}_has_qid"] = np.asarray([qid in ref[fk] for qid in query_ids])
    out["meta"] = np.asarray(json.dumps({"query_ids": query_ids, "cand": cand, "n_docs": n_docs}))
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN)


if __name__ == "__main__":
    main()
