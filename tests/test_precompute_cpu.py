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
