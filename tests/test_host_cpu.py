"""Host-side logic and the C-ABI surface - no GPU needed."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_library_loads_and_exports_every_header_symbol():
    from mfar_b200 import _native as nv
    lib = nv.lib()                                   # raises if the .so is missing: no fallback
    header = open(os.path.join(ROOT, "include", "mfar_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(mfar_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mfar_b200.h but not exported"
    assert declared == set(nv.PROTOTYPES), "ctypes prototype table out of sync with the header"
    assert lib.mfar_abi_version() == 1
    assert lib.mfar_status_string(0) == b"ok"
    assert b"sm_100" in lib.mfar_status_string(3)


def test_geometry_helpers_without_gpu():
    from mfar_b200 import _native as nv
    lib = nv.lib()
    assert lib.mfar_corpus_packed_elems(129375, 22, 768) == 1011 * 22 * 128 * 768
    assert lib.mfar_corpus_packed_elems(128, 1, 64) == 128 * 64
    assert lib.mfar_corpus_packed_elems(-1, 1, 64) == -1
    small = lib.mfar_score_topk_workspace_bytes(1, 100, 10_000_000, 0)
    big = lib.mfar_score_topk_workspace_bytes(512, 100, 10_000_000, 0)
    assert 0 < small < big < (1 << 30)
    assert lib.mfar_score_topk_workspace_bytes(64, 100, 129375, 22) > 64 * 129375 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_gpu():
    from mfar_b200 import _native as nv
    assert nv.lib().mfar_device_check(-1) != 0
    out = (ctypes.c_float * 4)()
    W = (ctypes.c_float * 4)()
    rc = nv.lib().mfar_mixture_weights(None, ctypes.addressof(W), None, 1, 0, 4, 0, ctypes.addressof(out), None)
    assert rc != 0
    with pytest.raises(RuntimeError):
        nv.check(rc, "mixture_weights")


def test_wrappers_refuse_cpu_tensors():
    from mfar_b200.data.index import DenseFlatIndex
    from mfar_b200.modeling.retrieval import PackedCorpus
    from mfar_b200.modeling.weighting import LinearWeights
    with pytest.raises(RuntimeError):
        PackedCorpus(10, 1, 64, device="cpu")
    with pytest.raises(RuntimeError):
        DenseFlatIndex(None, np.zeros((4, 8), np.float32), ["a"] * 4, {}, device="cpu")
    layer = LinearWeights(8, 3, query_cond=True)
    with pytest.raises(RuntimeError):
        layer(torch.zeros(2, 5, 3), torch.zeros(2, 8))


def test_resolve_fields_matches_reference_golden():
    from mfar_b200 import resolve_fields
    cases = json.load(open(os.path.join(GOLDEN, "resolve_fields.json")))
    assert len(cases) >= 20
    for key, expect in cases.items():
        dataset, names = key.split("|")
        if expect == "ValueError":
            with pytest.raises(ValueError):
                resolve_fields(names, dataset)
            continue
        got = resolve_fields(names, dataset)
        assert [[k, f.name, f.field_type.name, f.max_seq_length] for k, f in got.items()] == expect
    with pytest.raises(NotImplementedError):
        resolve_fields("all_dense", "nope")
    f = resolve_fields("all_dense,all_sparse", "prime")
    assert len(f) == 44 and list(f)[:22] == sorted(list(f)[:22]) and list(f)[22:] == sorted(list(f)[22:])
    assert len(resolve_fields("all_dense", "mag")) == 5 and len(resolve_fields("all_dense", "amazon")) == 8
    assert resolve_fields("expression.absent_dense", "prime")["expression absent_dense"].name == "expression absent"


def test_memory_map_dict_is_headerless_fp32(tmp_path):
    from mfar_b200 import MemoryMapDict
    path = str(tmp_path / "title.npy")
    open(path, "w").close()
    keys = [f"d{i}" for i in range(7)]
    m = MemoryMapDict(path, keys, (7, 16))
    for i, k in enumerate(keys):
        m[k] = np.full(16, i, np.float32)
    m.close()
    m.reopen()
    assert os.path.getsize(path) == 7 * 16 * 4                      # no .npy header (modeling/util.py:85-94)
    raw = np.fromfile(path, dtype=np.float32).reshape(7, 16)
    assert (raw[:, 0] == np.arange(7)).all() and (m["d3"] == 3).all()
    assert len(m) == 7 and "d2" in m and "zz" not in m and list(m) == keys
    with pytest.raises(NotImplementedError):
        del m["d0"]
    with pytest.raises(KeyError):
        m["zz"]
    assert m.rows_of(["d5", "d0"]).tolist() == [5, 0]
    blocks = list(m.iter_row_blocks(3))
    assert [lo for lo, _ in blocks] == [0, 3, 6] and blocks[1][1].shape == (3, 16) and blocks[1][1].flags["C_CONTIGUOUS"]


def test_read_and_create_indices_wiring(tmp_path, monkeypatch):
    """store + index wiring (modeling/util.py:73-108) - construction only, no compute."""
    from mfar_b200 import resolve_fields
    from mfar_b200.data import index as index_mod
    from mfar_b200.modeling import util

    class Enc:
        def get_sentence_embedding_dimension(self):
            return 32
    corpus = tmp_path / "corpus.tsv"
    corpus.write_text("".join(f"doc{i}\t{json.dumps({'title': 't%d' % i})}\n" for i in range(5)))
    fields = resolve_fields("title_dense,brand_dense,title_sparse", "amazon")
    # device="cuda" string is accepted without touching the GPU at construction time
    c, vd, idx = util.read_and_create_indices(str(corpus), "amazon", fields, str(tmp_path / "vec"), Enc())
    assert [k for k, _ in c] == [f"doc{i}" for i in range(5)]
    assert list(idx) == ["brand_dense", "title_dense", "title_sparse"]
    assert os.path.getsize(tmp_path / "vec" / "title.npy") == 5 * 32 * 4      # file named by field.name
    assert isinstance(idx["title_dense"], index_mod.DenseFlatIndex)
    assert isinstance(idx["title_sparse"], index_mod.PrecomputedSparseIndex)
    vd["title_dense"]["doc2"] = np.ones(32, np.float32)
    assert vd["title_dense"].file[2].sum() == 32


def test_key_packing_roundtrip_and_order():
    from mfar_b200.dist import decode_keys, encode_keys
    s = np.array([3.5, -1.25, 0.0, -0.0, 1e-30, -1e30, 3.5], np.float32)
    i = np.array([7, 1, 2, 3, 4, 5, 6], np.int64)
    k = encode_keys(s, i)
    ds, di = decode_keys(k)
    assert (di == i).all() and (ds == s).all()
    order = np.argsort(-k.astype(np.float64), kind="stable")        # coarse check via exact uint compare below
    srt = sorted(range(len(k)), key=lambda j: int(k[j]), reverse=True)
    # score desc, then id asc: (3.5,6) before (3.5,7)
    assert srt[:2] == [6, 0] and srt[-1] == 5
    es, ei = decode_keys(np.zeros(2, np.uint64))
    assert np.isneginf(es).all() and (ei == -1).all()


def test_shard_range_partitions_exactly():
    from mfar_b200.dist import shard_range
    for n in (1, 7, 128, 957192, 10_000_000):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def test_qres_line_format():
    from mfar_b200.data.trec import QRes
    line = str(QRes("q1", "d9", 0.5))
    assert line == "q1\t0\td9\t0\t0.5\t0"
    assert QRes.from_str(line).doc_id == "d9"


def test_precomputed_sparse_scores_loader_matches_reference_dict_semantics(tmp_path):
    """{field}_keys_bm25.npy int32 [nnz,2] + {field}_vals_bm25.npy f16 [nnz] (precompute_bm25s_scores.py:21-30):
    lookups behave like the reference's nested dicts (modeling/util.py:112-173, index.py:120-125: missing -> 0),
    and batch() emits (row-in-batch, doc) pairs grouped by field."""
    from mfar_b200.data.typedef import Field, FieldType
    from mfar_b200.modeling.util import PrecomputedSparseScores
    rng = np.random.RandomState(0)
    finfo = {"a_dense": Field("a_dense", "a", FieldType.DENSE), "a_sparse": Field("a_sparse", "a", FieldType.SPARSE),
             "b_sparse": Field("b_sparse", "b", FieldType.SPARSE)}
    ref = {}
    for fk in ("a_sparse", "b_sparse"):
        qids = rng.randint(0, 20, size=300)
        docs = rng.randint(0, 1000, size=300)
        pairs = np.unique(np.stack([qids, docs], axis=1), axis=0)
        pairs = pairs[rng.permutation(len(pairs))].astype(np.int32)
        vals = rng.gamma(2.0, 2.0, size=len(pairs)).astype(np.float16)
        np.save(tmp_path / f"{fk}_keys_bm25.npy", pairs)
        np.save(tmp_path / f"{fk}_vals_bm25.npy", vals)
        d = {}
        for (qi, di), v in zip(pairs.tolist(), vals.tolist()):          # the reference's {qid: {doc: score}}
            d.setdefault(qi, {})[di] = v
        ref[fk] = d
    store = PrecomputedSparseScores.load(str(tmp_path), finfo)
    assert store.field_keys == ["a_sparse", "b_sparse"]
    for fk in store.field_keys:
        for qid in (0, 3, 19, 25):
            for doc in (0, 5, 999) + tuple(ref[fk].get(qid, {}).keys())[:3]:
                assert store.lookup(fk, qid, doc) == ref[fk].get(qid, {}).get(doc, 0)
    batch_q = [7, 25, 3]
    keys, vals, offs = store.batch(batch_q, device="cpu")
    assert keys.dtype == torch.int32 and keys.shape[1] == 2 and len(offs) == 3 and offs[0] == 0 and offs[-1] == len(vals)
    for j, fk in enumerate(store.field_keys):
        seg_k, seg_v = keys[offs[j]:offs[j + 1]].numpy(), vals[offs[j]:offs[j + 1]].numpy()
        want = {(row, doc): v for row, qid in enumerate(batch_q) for doc, v in ref[fk].get(qid, {}).items()}
        got = {(int(r), int(d)): float(v) for (r, d), v in zip(seg_k, seg_v)}
        assert got == {k_: float(np.float16(v)) for k_, v in want.items()}


def test_mask_sweep_plan_follows_mask_fields_command():
    """mfar/commands/mask_fields.py:142-170: baseline, each field, all sparse, all dense, each field name."""
    from mfar_b200.data.schema import resolve_fields
    from mfar_b200.modeling.retrieval import MultiFieldRetriever
    finfo = resolve_fields("all_dense,all_sparse", "amazon")
    F = len(finfo)
    plan = MultiFieldRetriever.mask_sweep_plan(finfo)
    names = sorted({f.name for f in finfo.values()})
    assert len(plan) == 1 + F + 2 + len(names)
    assert plan[0] == ("baseline", [])
    assert [p[1] for p in plan[1:1 + F]] == [[i] for i in range(F)]
    assert plan[1 + F] == ("all_sparse", list(range(F // 2, F))) and plan[2 + F] == ("all_dense", list(range(F // 2)))
    for (label, idx), name in zip(plan[3 + F:], names):
        assert label == f"name:{name}" and len(idx) == 2            # one dense + one sparse column per name
    assert len(MultiFieldRetriever.mask_sweep_plan(resolve_fields("all_dense", "mag"))) == 1 + 5 + 1 + 5


def test_every_package_module_imports_without_gpu():
    import importlib
    for name in ("mfar_b200", "mfar_b200._native", "mfar_b200.dist", "mfar_b200.synth", "mfar_b200.data.bm25",
                 "mfar_b200.data.index", "mfar_b200.data.schema", "mfar_b200.data.trec", "mfar_b200.data.typedef",
                 "mfar_b200.data.util", "mfar_b200.modeling.losses", "mfar_b200.modeling.retrieval",
                 "mfar_b200.modeling.util", "mfar_b200.modeling.weighting"):
        importlib.import_module(name)


def test_training_scorer_refuses_cpu_tensors_and_bad_layouts():
    from mfar_b200 import _native as nv
    from mfar_b200.modeling import losses as L
    with pytest.raises(RuntimeError):
        L.field_components(torch.zeros(2, 8), torch.zeros(3, 2, 8), 1.0)
    assert L._layout(torch.zeros(5, 3, 8)) == (5, 3, 1, 24, 8, 0)
    assert L._layout(torch.zeros(5, 3, 4, 8)) == (20, 3, 4, 96, 32, 8)
    with pytest.raises(ValueError):
        L._layout(torch.zeros(5, 8))
    lib = nv.lib()
    fake = 0x1000
    # argument checks precede device work: E not a multiple of 4, E > 1024, N not a multiple of inner, T <= 0
    assert lib.mfar_field_components_fwd(fake, 2, 6, fake, 4, 2, 1, 12, 6, 0, 1.0, fake, None) == 2
    assert lib.mfar_field_components_fwd(fake, 2, 2048, fake, 4, 2, 1, 4096, 2048, 0, 1.0, fake, None) == 2
    assert lib.mfar_field_components_fwd(fake, 2, 8, fake, 5, 2, 2, 32, 16, 8, 1.0, fake, None) == 1
    assert lib.mfar_field_components_fwd(fake, 2, 8, fake, 4, 2, 1, 16, 8, 0, 0.0, fake, None) == 1
    assert lib.mfar_field_components_bwd(fake, 2, 8, fake, 4, 2, 1, 16, 8, 0, 1.0, fake, None, None, None) == 1
    assert lib.mfar_mixture_bwd(fake, fake, fake, fake, 3, fake, 2, 4, 8, 2, 1, None, fake, None, fake, None) == 2


def test_header_is_valid_c_and_links_against_the_library(tmp_path):
    """include/mfar_b200.h must be consumable from plain C (the drop-in boundary is a C ABI): compile a C translation
    unit that takes the address of every declared entry point, link it against the .so, run it."""
    import shutil
    import subprocess
    from mfar_b200 import _native as nv
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    header = open(os.path.join(ROOT, "include", "mfar_b200.h")).read()
    names = sorted(set(re.findall(r"\b(mfar_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", header, flags=re.S))))
    src = tmp_path / "abi.c"
    src.write_text('#include "mfar_b200.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\nint main(void) {\n  fn_t fns[] = {\n' +
                   "".join(f"    (fn_t)&{n},\n" for n in names) +
                   '  };\n  if (mfar_abi_version() != MFAR_ABI_VERSION) return 2;\n'
                   '  printf("%d %s\\n", (int)(sizeof fns / sizeof fns[0]), mfar_status_string(MFAR_ERR_ARCH));\n  return 0;\n}\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(nv.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-l:libmfar_b200.so", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split(" ", 1)
    assert int(out[0]) == len(names) == len(nv.PROTOTYPES) and "sm_100" in out[1]


def test_sweep_case_generators_stay_inside_the_kernels_envelope():
    """tools/fuzz_parity.py draws (the GPU sweep itself needs a B200): k <= N, at least one field, the tensor-core
    requests only where the C ABI accepts them (dim % 64 == 0, a dense field; query-stationary: dim <= 768)."""
    import os
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_parity as F
    rng = np.random.RandomState(0)
    for _ in range(500):
        c = F.draw_case(rng)
        assert 1 <= c["k"] <= c["N"] and c["k"] <= 128 and c["Fd"] + c["Fs"] >= 1
        if c["impl"] in ("tcgen05", "tcgen05_qs"):
            assert c["Fd"] >= 1 and c["d"] % 64 == 0
        if c["impl"] == "tcgen05_qs":
            assert c["d"] <= 768
        assert c["base"] + c["N"] <= 2 ** 32
        a = F.draw_api_case(rng)
        assert a["Fd"] >= 1 and 2 <= a["N"] and a["k"] <= a["N"] and a["impl"] == "auto"
        b = F.draw_bm25_case(rng)
        assert b["N"] >= 1 and b["V"] >= 1 and b["Fs"] >= 1
        t = F.draw_train_case(rng)
        assert t["E"] % 4 == 0 and t["E"] <= 1024
    assert set(F.MODES) == {"kernels", "api", "bm25", "train"}


def test_weighted_shard_ranges_cover_the_corpus_and_follow_the_weights():
    from mfar_b200.dist import shard_range, weighted_shard_ranges
    n = 10_000_000
    eq = weighted_shard_ranges(n, [1.0] * 8)
    assert eq[0][0] == 0 and eq[-1][1] == n and all(a[1] == b[0] for a, b in zip(eq, eq[1:]))
    assert all(abs((hi - lo) - n / 8) <= 128 for lo, hi in eq)
    assert all(lo % 128 == 0 for lo, _ in eq)
    assert all(abs(lo - shard_range(n, r, 8)[0]) <= 64 for r, (lo, _) in enumerate(eq))
    w = [1.00, 0.96, 1.02, 0.99, 1.03, 0.97, 1.0, 1.01]
    ws = weighted_shard_ranges(n, w)
    assert ws[0][0] == 0 and ws[-1][1] == n and all(a[1] == b[0] for a, b in zip(ws, ws[1:]))
    sizes = [hi - lo for lo, hi in ws]
    for s_, w_ in zip(sizes, w):
        assert abs(s_ - n * w_ / sum(w)) <= 256
    # degenerate inputs: a tiny corpus, a zero weight
    tiny = weighted_shard_ranges(100, [1, 1, 1, 1])
    assert tiny[0][0] == 0 and tiny[-1][1] == 100 and all(lo <= hi for lo, hi in tiny)
    z = weighted_shard_ranges(1000, [1.0, 0.0, 1.0], align=1)
    assert z[1][1] - z[1][0] <= 1 and z[-1][1] == 1000


def test_rebalanced_boundaries_converge_to_equal_time_inside_the_margins():
    from mfar_b200.dist import held_ranges, rebalanced_boundaries, weighted_shard_ranges
    n, R = 10_000_000, 8
    base = [r[0] for r in weighted_shard_ranges(n, [1.0] * R)] + [n]
    margin = 75_008
    held = held_ranges(base, margin, n)
    assert held[0][0] == 0 and held[-1][1] == n and all(lo % 128 == 0 for lo, _ in held)
    assert all(h[0] == b - margin for h, b in zip(held[1:], base[1:]))
    docs_per_ms = [201.6, 198.4, 195.3, 201.6, 186.6, 201.6, 195.3, 201.6]    # what each "GPU" really does
    cuts = list(base)
    for _ in range(4):
        t = [(cuts[r + 1] - cuts[r]) / docs_per_ms[r] for r in range(R)]
        cuts = rebalanced_boundaries(cuts, t, base, margin)
        assert cuts[0] == 0 and cuts[-1] == n and all(a < b for a, b in zip(cuts, cuts[1:]))
        assert all(c % 128 == 0 and abs(c - b) <= margin for c, b in zip(cuts[1:-1], base[1:-1]))
        for r in range(R):                                   # the active range stays inside what the rank holds
            assert held[r][0] <= cuts[r] and cuts[r + 1] <= held[r][1]
    t = [(cuts[r + 1] - cuts[r]) / docs_per_ms[r] for r in range(R)]
    assert (max(t) - min(t)) / min(t) < 1e-3
    # a rank far slower than the margin allows: boundaries stop at the margin instead of leaving the resident range
    slow = rebalanced_boundaries(list(base), [1.0, 3.0] + [1.0] * 6, base, margin, damping=1.0)
    assert slow[1] == base[1] + margin and slow[2] >= base[2] - margin
    # equal times leave equal boundaries alone
    assert rebalanced_boundaries(list(base), [5.0] * R, base, margin) == base


def test_sparse_score_store_matches_the_reference_loader_golden(tmp_path):
    """tests/golden/loader/sparse_scores.npz was produced by the reference's own ``read_sparse_scores`` +
    ``score_batch_with_cache`` (oracle/make_golden_sparse_scores.py): same [Q,C] matrices (missing -> 0, the later of
    duplicate pairs wins), same ``qid in by_field`` answers, through the store and through ``BM25sSparseIndex``."""
    import json as _json
    from mfar_b200.data.index import BM25sSparseIndex
    from mfar_b200.data.typedef import Field, FieldType
    from mfar_b200.modeling.util import read_sparse_scores
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "loader", "sparse_scores.npz"))
    meta = _json.loads(str(z["meta"]))
    finfo = {"a_dense": Field("a_dense", "a", FieldType.DENSE), "a_sparse": Field("a_sparse", "a", FieldType.SPARSE),
             "b_sparse": Field("b_sparse", "b", FieldType.SPARSE)}
    for fk in ("a_sparse", "b_sparse"):
        np.save(tmp_path / f"{fk}_keys_bm25.npy", z[f"{fk}_keys"])
        np.save(tmp_path / f"{fk}_vals_bm25.npy", z[f"{fk}_vals"])
    store = read_sparse_scores(str(tmp_path), finfo)
    assert store.keys() == ["a_sparse", "b_sparse"] and len(store.values()) == 2 and bool(store)
    keys_all = [str(i) for i in range(meta["n_docs"])]
    index = BM25sSparseIndex(keys_all, index=None, stemmer=None)
    for fk in store.keys():
        by_field = store[fk]
        assert [qid in by_field for qid in meta["query_ids"]] == z[f"{fk}_has_qid"].tolist()
        got = index.score_batch_with_cache(meta["query_ids"], meta["cand"], by_field)
        assert got.dtype == torch.float32 and np.array_equal(got.numpy(), z[f"{fk}_cached"])
        # dict-style access agrees as well
        row = by_field.get(meta["query_ids"][0], {})
        assert all(np.float32(row.get(int(c), 0)) == z[f"{fk}_cached"][0, j] for j, c in enumerate(meta["cand"]))
