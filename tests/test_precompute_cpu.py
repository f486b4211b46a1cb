"""Producer of the precomputed-BM25 score files: oracle vs goldens produced by the reference's own
``precompute_score_for_field`` (oracle/make_golden_precompute.py), host-side bitmap / file round trip, C-ABI argument
checks.  No GPU needed."""
import glob
import os

import numpy as np
import pytest

import precompute_oracle as PO

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "precompute", "*.npz")))


def _load(path):
    z = np.load(path, allow_pickle=False)
    rows = {int(q): z["scores"][i] for i, q in enumerate(z["qids"])}
    return z, rows


def test_goldens_present():
    assert len(GOLDEN) == 4


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_reference_precompute_score_for_field(path):
    z, rows = _load(path)
    keys, vals = PO.precompute_score_for_field(rows, z["safe"].tolist())
    assert keys.dtype == np.int32 and vals.dtype == np.float16
    np.testing.assert_array_equal(keys, z["ref_keys"])
    np.testing.assert_array_equal(vals.view(np.uint16), z["ref_vals"].view(np.uint16))       # bit-exact, inf included


def test_extremes_golden_holds_the_f16_edge_cases():
    z, _ = _load([p for p in GOLDEN if p.endswith("pre_extremes.npz")][0])
    vals = z["ref_vals"]
    assert np.isinf(vals).any()              # 7e4 / 65519.9 overflow to inf in np.float16
    assert (vals == 0).any()                 # 1e-9 underflows to 0 but the pair is still written (filter is on fp32)
    assert (vals < 0).any()                  # the filter is `!= 0`, not `> 0`


def test_safe_docs_bitmap():
    from mfar_b200.data.bm25 import safe_docs_bitmap
    bits = safe_docs_bitmap({0, 31, 32, 95, 1000, -3}, 96)
    assert bits.dtype == np.uint32 and bits.shape == (3,)
    assert bits.tolist() == [(1 << 0) | (1 << 31), 1 << 0, 1 << 31]
    assert safe_docs_bitmap(set(), 10).tolist() == [0]
    got = safe_docs_bitmap(range(0, 70, 3), 70)
    want = np.zeros(70, bool)
    want[::3] = True
    assert np.array_equal(np.unpackbits(got.view(np.uint8), bitorder="little")[:70].astype(bool), want)


def test_written_files_round_trip_through_the_reader(tmp_path):
    """oracle writer -> the reference's file layout -> PrecomputedSparseScores (reader of a4) -> dict semantics."""
    from mfar_b200.data.typedef import Field, FieldType
    from mfar_b200.modeling.util import PrecomputedSparseScores
    z, rows = _load([p for p in GOLDEN if p.endswith("pre_small.npz")][0])
    keys, vals = PO.precompute_score_for_field(rows, z["safe"].tolist())
    np.save(tmp_path / "f_sparse_keys_bm25.npy", keys)
    np.save(tmp_path / "f_sparse_vals_bm25.npy", vals)
    store = PrecomputedSparseScores.load(str(tmp_path), {"f_sparse": Field("f_sparse", "f", FieldType.SPARSE)})
    safe = set(z["safe"].tolist())
    for qid, row in rows.items():
        for doc in (0, 7, 123, 299):
            want = float(np.float16(row[doc])) if (row[doc] != 0 and doc in safe) else 0.0
            assert store.lookup("f_sparse", qid, doc) == want


def test_c_abi_argument_checks_without_a_gpu():
    from mfar_b200 import _native as nv
    lib = nv.lib()
    assert lib.mfar_sparse_coo_offsets_len(0, 10) == 0 and lib.mfar_sparse_coo_offsets_len(3, 0) == 0
    assert lib.mfar_sparse_coo_offsets_len(3, 4096) == 3 * 1 + 1
    assert lib.mfar_sparse_coo_offsets_len(3, 4097) == 3 * 2 + 1
    assert lib.mfar_sparse_coo_count(0, 10, 1, 10, 0, 0, 0, 0) == 1                      # null pointers: MFAR_ERR_ARG
    assert lib.mfar_sparse_coo_count(256, 5, 1, 10, 0, 0, 256, 0) == 1                   # ld < n_docs
    assert lib.mfar_sparse_coo_count(256, 10, 1, 10, 0, 2**31 - 5, 256, 0) == 2          # doc ids beyond int32
    assert lib.mfar_sparse_coo_write(256, 10, 1, 10, 0, 0, 0, 256, 0, 0, nv.F16, 0) == 1  # null outputs
    assert lib.mfar_sparse_coo_write(256, 10, 1, 10, 0, 0, 0, 256, 256, 256, nv.BF16, 0) == 1


class _FakeSparseIndex:
    """Host stand-in for BM25sSparseIndex (the device class is covered by test_gpu_precompute.py): serves given score
    rows through the same two methods the command calls."""

    def __init__(self, rows_by_text):
        self.rows, self.safe, self.batches = rows_by_text, None, []

    def set_safe_docs(self, safe_docs):
        self.safe = set(safe_docs)

    def get_scores_sparse_batch(self, queries, query_ids=None):
        self.batches.append(len(queries))
        return PO.precompute_score_for_field({qid: self.rows[q] for qid, q in zip(query_ids, queries)}, self.safe)

    def retrieve_batch(self, queries, top_k):
        return [[(str(d), float(self.rows[q][d])) for d in np.argsort(-self.rows[q], kind="stable")[:top_k]] for q in queries]


def test_command_host_logic_batches_orders_and_files(tmp_path):
    from mfar_b200.commands import precompute_bm25s_scores as C
    z, rows = _load([p for p in GOLDEN if p.endswith("pre_small.npz")][0])
    qids = [int(q) for q in z["qids"]]
    texts = {qid: f"text {i}" for i, qid in enumerate(qids)}
    index = _FakeSparseIndex({texts[qid]: rows[qid] for qid in qids})
    keys, vals = C.precompute_score_for_field(index, z["safe"].tolist(), texts, str(tmp_path), "f_sparse", batch_size=2)
    assert index.batches == [2, 2, 1]                                   # 5 queries in batches of 2, order kept
    np.testing.assert_array_equal(keys, z["ref_keys"])                  # == what the reference's function wrote
    np.testing.assert_array_equal(vals.view(np.uint16), z["ref_vals"].view(np.uint16))
    np.testing.assert_array_equal(np.load(tmp_path / "f_sparse_keys_bm25.npy"), z["ref_keys"])
    assert np.load(tmp_path / "f_sparse_vals_bm25.npy").dtype == np.float16
    # train.queries / train.qrels readers and the candidate set (top-k negatives united with the positives)
    (tmp_path / "train.queries").write_text("".join(f"{qid}\t{texts[qid]}\n" for qid in qids))
    (tmp_path / "train.qrels").write_text(f"{qids[0]}\t0\t7\t1\n{qids[1]}\t0\t299\t1\n")
    queries, pos = C.read_queries_and_positives(str(tmp_path))
    assert queries == texts and pos == {7, 299}
    cand = C.candidate_docs(index, list(queries.values()), pos, top_k=3, batch_size=2)
    want = {7, 299}
    for q in queries.values():
        want |= set(np.argsort(-index.rows[q], kind="stable")[:3].tolist())
    assert cand == want
    with pytest.raises(ValueError):
        C.main(str(tmp_path), "mag", str(tmp_path), str(tmp_path), fields_str="all_dense")
