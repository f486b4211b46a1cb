"""Oracle-RANKED parity at the sizes BASELINE.json names (C2..C5), for every tensor-core kernel.

The small-shape tests (test_gpu_parity.py) give each CTA one or two corpus tiles, so the top-k admission machinery -
list overflow -> ``warp_select_list``, the pooled rank-r bound, the shared ``gthr`` thresholds (csrc/common.cuh) -
hardly runs there.  Here every shard is corpus-sized and the result is ranked against ``tests/checker.py`` (plain
torch fp32 on the GPU, TF32 off, independent of the library's scoring kernels) with the full near-tie rule of
``tests/parity.py``: a doc the kernel dropped must near-tie the oracle's k-th score.

Reference semantics checked: mfar/data/index.py:181-222 (per-field exhaustive scores), mfar/modeling/contrastive.py
:686-696 (mask, mixture, top-100), mfar/modeling/weighting.py:17-29.
"""
import gc

import pytest
import torch

from checker import Fp32Checker, assert_topk_parity_at_scale

pytestmark = pytest.mark.gpu

DEV = "cuda"
K = 100
D = 768

# name: (n_docs, n_dense, n_sparse, doc_id_base, query_cond, masked field or None, seed, largest batch)
CASES = {
    "c2_prime_hybrid": (129_375, 22, 22, 0, True, 3, 101, 200),
    "c3_mag": (700_244, 5, 0, 0, True, None, 102, 512),
    "c4_amazon_hybrid": (957_192, 8, 8, 0, True, None, 103, 512),
    "c5_shard_of_8": (1_250_000, 8, 0, 3_750_000, True, None, 104, 512),      # rank 3's doc range of the 8-GPU run
    "c5_single": (10_000_000, 1, 0, 0, False, None, 105, 512),
    "c5_all": (10_000_000, 8, 0, 0, True, None, 1234, 512),
    # threshold-seeding edges (csrc/capi.cu): a ragged last tile, doc ids near 2^32, fp32 sparse rows behind the
    # advanced pointers; Q=40 seeds with one prefix level (doc-stationary), Q=300 with two (query-stationary, 37 lists)
    "c6_seed_edges": (330_001, 2, 2, 3_900_000_000, True, 1, 106, 300, torch.float32),
}

# (case, batch, impl).  impl "tcgen05" = doc-stationary kernel, "tcgen05_qs" = query-stationary, "auto" = what ships.
RUNS = []
for _c, _batches in (("c2_prime_hybrid", (1, 64, 200)), ("c3_mag", (1, 512)), ("c4_amazon_hybrid", (64, 512)),
                     ("c5_shard_of_8", (1, 512)), ("c5_single", (1, 64, 128, 512)), ("c5_all", (1, 64, 512)),
                     ("c6_seed_edges", (40, 300))):
    for _b in _batches:
        for _impl in ("tcgen05", "tcgen05_qs", "auto"):
            RUNS.append((_c, _b, _impl))

_cache = {}


def _free():
    _cache.clear()
    gc.collect()
    torch.cuda.empty_cache()


def _make_sparse(Q, Fs, n, seed, dtype=torch.float16):
    from mfar_b200 import synth as S
    ld = (n + 63) // 64 * 64                       # 128-byte rows: gathered inside the scoring epilogue
    out = torch.zeros((Q, Fs, ld), dtype=dtype, device=DEV)
    for q0 in range(0, Q, 32):                     # bounded temporaries (fp32 [32,Fs,N] x 4)
        q1 = min(Q, q0 + 32)
        out[q0:q1, :, :n] = S.make_sparse(q1 - q0, Fs, n, seed + q0, DEV)
    return out


def _case(name):
    """Corpus, queries, weights and the checker's exact top-(k+slack) for the largest batch - built once per case;
    smaller batches are prefixes of the same query set."""
    if name in _cache:
        return _cache[name]
    _free()
    from mfar_b200 import synth as S
    from mfar_b200.modeling.retrieval import MultiFieldRetriever, PackedCorpus
    from mfar_b200.modeling.weighting import LinearWeights
    n, Fd, Fs, base, qc, masked, seed, Qmax = CASES[name][:8]
    sp_dtype = CASES[name][8] if len(CASES[name]) > 8 else torch.float16
    need = n * Fd * D * 2 + Qmax * Fs * n * 2 + 8e9
    free, _ = torch.cuda.mem_get_info()
    if free < need:
        pytest.skip(f"needs ~{need / 1e9:.0f} GB of free HBM")
    pc = PackedCorpus(n, Fd, D, DEV)
    S.fill_packed_corpus(pc, seed=seed)
    mu = S.corpus_mean(D, seed, DEV)
    q = S.make_queries(Qmax, D, mu, seed + 1, DEV)
    F = Fd + Fs
    layer = LinearWeights(D, F, query_cond=True) if qc else LinearWeights(F, 1)
    with torch.no_grad():
        layer.weight.copy_(S.make_mixture(D, F, seed + 2, query_cond=qc))
    layer = layer.to(DEV)
    sp = _make_sparse(Qmax, Fs, n, seed + 3, sp_dtype) if Fs else None
    r = MultiFieldRetriever(pc, layer, n_sparse=Fs, top_k=K, doc_id_base=base)
    if masked is not None:
        r.mask_field([masked])
    # mixture weights restated with plain torch (weighting.py:25-28, contrastive.py:686), not the library's kernel
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    logits = q.float() @ layer.weight.detach().float() if qc else layer.weight.detach().float().t().expand(Qmax, F)
    torch.backends.cuda.matmul.allow_tf32 = prev
    w = (torch.softmax(logits, dim=1) * r.mask.reshape(1, F)).contiguous()
    chk = Fp32Checker(pc, doc_id_base=base)
    ref_s, ref_i = chk.topk(q, w, K, sparse=sp, slack=64)
    torch.cuda.synchronize()
    _cache[name] = dict(r=r, q=q, w=w, sp=sp, chk=chk, ref_s=ref_s, ref_i=ref_i, n=n, base=base)
    return _cache[name]


@pytest.mark.parametrize("case,batch,impl", RUNS, ids=[f"{c}-Q{b}-{i}" for c, b, i in RUNS])
def test_ranked_against_fp32_checker(case, batch, impl):
    c = _case(case)
    r, q, sp = c["r"], c["q"][:batch], None if c["sp"] is None else c["sp"][:batch]
    s, i = r.search(q, q.float(), sp, impl=impl)
    torch.cuda.synchronize()
    resc = c["chk"].rescore(q, c["w"][:batch], i, sparse=sp)
    assert_topk_parity_at_scale(s, i, c["ref_s"][:batch], c["ref_i"][:batch], resc, K, c["n"], c["base"],
                                what=f"{case} Q={batch} {impl}")


def test_c5_all_shards_and_host_call_equal_unsharded():
    """10M x 8: (a) two doc-range shards over the same packed tensor, merged by ``mfar_topk_merge``, reproduce the
    unsharded result bit for bit; (b) ``mfar_search_host`` (pinned host buffers) == the device call."""
    from mfar_b200.dist import merge_keys
    from mfar_b200.modeling.retrieval import MultiFieldRetriever, PackedCorpus
    c = _case("c5_all")
    r, q = c["r"], c["q"][:140]
    s_c, i_c, _ = r.search(q, q.float(), return_keys=True)
    pc = r.corpus
    cut = 128 * 40_000
    parts = []
    for lo, hi in ((0, cut), (cut, pc.n_docs)):
        sh = MultiFieldRetriever(pc.window(lo, hi - lo), r.mixture, top_k=K, doc_id_base=lo)
        parts.append(sh.search(q, q.float(), return_keys=True)[2])
    s_m, i_m = merge_keys(torch.stack(parts), K)
    assert torch.equal(i_m, i_c) and torch.equal(s_m, s_c)
    s_h, i_h = r.search_host(q.cpu().pin_memory(), q.float().cpu().pin_memory())
    assert torch.equal(s_h, s_c.cpu()) and torch.equal(i_h, i_c.cpu())


def test_zz_release_hbm():
    _free()
