"""Property tests (hypothesis) of the host-side mirrors the device code relies on: the packed (score, doc id) key
order, doc-range sharding, the safe-set bitmap, the COO batch slicer, the oracle of the pair producer.  No GPU."""
import numpy as np
from hypothesis import given, settings, strategies as st

from mfar_b200.dist import decode_keys, encode_keys, shard_range

finite_f32 = st.floats(width=32, allow_nan=False, allow_infinity=False)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(finite_f32, st.integers(0, 2 ** 32 - 2)), min_size=2, max_size=40))
def test_key_order_is_score_desc_then_doc_asc(pairs):
    s = np.array([p[0] for p in pairs], np.float32)
    d = np.array([p[1] for p in pairs], np.int64)
    keys = encode_keys(s, d)
    order = np.argsort(keys)[::-1]                                  # max key first
    want = np.lexsort((d, -s.astype(np.float64)))                   # score descending, then doc id ascending
    # -0.0 and +0.0 are distinct keys (sign bit) but equal floats: compare through the key's own decode
    ks, kd = decode_keys(keys[order])
    assert np.array_equal(np.sort(kd), np.sort(d))
    assert np.all(np.diff(ks.astype(np.float64)) <= 0)
    same = s[want].astype(np.float64) == ks.astype(np.float64)
    assert same.all()
    for a, b in zip(order[:-1], order[1:]):                         # equal scores (same bit pattern): lower id first
        if s[a].tobytes() == s[b].tobytes():
            assert d[a] < d[b] or (d[a] == d[b])
    rs, rd = decode_keys(keys)
    assert np.array_equal(rs.view(np.uint32), s.view(np.uint32)) and np.array_equal(rd, d)
    assert (keys != 0).all()                                        # 0 is reserved for "empty slot"


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 10 ** 7), st.integers(1, 8))
def test_shard_ranges_tile_the_corpus(n, world):
    cuts = [shard_range(n, r, world) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
    sizes = [hi - lo for lo, hi in cuts]
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 3000), st.data())
def test_safe_bitmap_membership(total, data):
    from mfar_b200.data.bm25 import safe_docs_bitmap
    ids = data.draw(st.sets(st.integers(-5, total + 40), max_size=200))
    bits = safe_docs_bitmap(ids, total)
    assert bits.dtype == np.uint32 and len(bits) == (total + 31) // 32
    got = np.unpackbits(bits.view(np.uint8), bitorder="little")[:total].astype(bool)
    want = np.zeros(total, bool)
    for i in ids:
        if 0 <= i < total:
            want[i] = True
    assert np.array_equal(got, want)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 6), st.integers(1, 300), st.floats(0.0, 1.0), st.integers(0, 2 ** 31 - 1))
def test_pair_producer_oracle_invariants(Q, N, density, seed):
    import precompute_oracle as PO
    rng = np.random.RandomState(seed)
    rows = {10 * q + 3: np.where(rng.rand(N) < density, rng.randn(N) * 5, 0.0).astype(np.float32) for q in range(Q)}
    safe = set(rng.choice(N, size=max(1, N // 2), replace=False).tolist())
    keys, vals = PO.precompute_score_for_field(rows, safe)
    assert keys.shape == (len(vals), 2) and keys.dtype == np.int32 and vals.dtype == np.float16
    assert len(vals) == sum(int(((r != 0) & np.isin(np.arange(N), list(safe))).sum()) for r in rows.values())
    qpos = {q: i for i, q in enumerate(rows)}
    order = [(qpos[int(q)], int(d)) for q, d in keys]
    assert order == sorted(order)                                   # queries in dict order, docs ascending
    for (q, d), v in zip(keys, vals):
        assert int(d) in safe and rows[int(q)][d] != 0 and np.float16(rows[int(q)][d]).tobytes() == v.tobytes()


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 5), st.integers(1, 4), st.integers(0, 2 ** 31 - 1))
def test_coo_batch_slicer_matches_dict_lookup(n_q, n_f, seed):
    """PrecomputedSparseScores.batch (device COO input of a query batch) vs the reference's nested-dict lookup
    (index.py:120-125): every (query row, field, doc) of the batch carries the stored value, nothing else appears."""
    from mfar_b200.modeling.util import PrecomputedSparseScores
    rng = np.random.RandomState(seed)
    fks = [f"f{j}_sparse" for j in range(n_f)]
    per_field, dicts = {}, {}
    all_qids = list(range(100, 100 + 3 * n_q, 3))
    for fk in fks:
        n = rng.randint(0, 60)
        q = rng.choice(all_qids + [999], size=n)
        d = rng.randint(0, 50, size=n)
        pairs = {}
        for a, b in zip(q, d):
            pairs[(int(a), int(b))] = np.float16(rng.rand() * 9)    # unique pairs (the files hold no duplicates)
        ks = np.array(list(pairs.keys()), np.int32).reshape(-1, 2)
        vs = np.array(list(pairs.values()), np.float16)
        per_field[fk] = (ks, vs)
        dicts[fk] = pairs
    store = PrecomputedSparseScores(per_field, fks)
    batch_q = [all_qids[i] for i in rng.permutation(len(all_qids))[: max(1, n_q - 1)]]
    keys, vals, offs = store.batch(batch_q, device="cpu")
    keys, vals = keys.numpy(), vals.numpy()
    assert len(offs) == n_f + 1 and offs[0] == 0 and offs[-1] == len(keys)
    for j, fk in enumerate(fks):
        seg_k, seg_v = keys[offs[j]:offs[j + 1]], vals[offs[j]:offs[j + 1]]
        got = {(batch_q[int(r)], int(d)): v for (r, d), v in zip(seg_k, seg_v)}
        want = {(q, d): v for (q, d), v in dicts[fk].items() if q in batch_q}
        assert got == want
