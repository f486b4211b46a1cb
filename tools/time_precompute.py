#!/usr/bin/env python
"""Times the device producer of the precomputed-BM25 score files (csrc/sparse_coo.cu) on an Amazon-shaped sparse
field and the CPU restatement of the reference's procedure beside it (needs a B200; not part of pytest).

    python tools/time_precompute.py [--docs 957192] [--batch 64] [--out gpurun_out/precompute_timing.json]

Per batch of Q queries: BM25 postings scatter (mfar_bm25_scores) -> count + scan -> write (mfar_sparse_coo_*), timed
with CUDA events on the launch stream; the CPU side is precompute_oracle.precompute_score_for_field over the same score
rows for a few queries (the reference walks Python dicts per query, precompute_bm25s_scores.py:19-24).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "multifield-adaptive-retrieval_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=957192)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--safe-frac", type=float, default=0.1)
    ap.add_argument("--cpu-queries", type=int, default=4)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "precompute_timing.json"))
    args = ap.parse_args()
    from mfar_b200 import _native as nv
    from mfar_b200 import synth
    from mfar_b200.data.bm25 import BM25FieldSet, rows_to_coo, safe_docs_bitmap
    import precompute_oracle as PO
    dev = torch.device("cuda", 0)
    N, Q = args.docs, args.batch
    field = synth.make_bm25_field(N, 300, dev)
    fs = BM25FieldSet([field])
    rng = np.random.default_rng(0)
    safe = rng.choice(N, size=int(args.safe_frac * N), replace=False)
    bits = torch.from_numpy(safe_docs_bitmap(safe.tolist(), N).view(np.int32)).to(dev)
    batches = [synth.make_bm25_query_entries(Q, 1, 200 + i).to(dev) for i in range(4)]
    qids = torch.arange(Q, dtype=torch.int32, device=dev)

    def one(ent):
        sv = fs.field_scores(ent, Q)[:, 0, :N]
        return sv, rows_to_coo(sv, N, bits, qids, 0)

    for i in range(3):
        sv, (k, v) = one(batches[i % 4])
    torch.cuda.synchronize()
    nnz = int(k.shape[0])
    # stage timings: scatter | count+scan | write  (events on the current stream; the nnz read-back syncs in between)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_sc, t_cnt, t_wr, t_all = [], [], [], []
    lib = nv.lib()
    for i in range(args.steps):
        ent = batches[i % 4]
        w0 = time.perf_counter()
        ev[0].record()
        sv = fs.field_scores(ent, Q)[:, 0, :N]
        ev[1].record()
        offs = torch.empty(lib.mfar_sparse_coo_offsets_len(Q, N), dtype=torch.int64, device=dev)
        nv.check(lib.mfar_sparse_coo_count(nv.ptr(sv), sv.stride(0), Q, N, nv.ptr(bits), 0, nv.ptr(offs), nv.stream()))
        ev[2].record()
        n = int(offs[-1].item())
        keys = torch.empty((n, 2), dtype=torch.int32, device=dev)
        vals = torch.empty((n,), dtype=torch.float16, device=dev)
        ev[2].record()
        nv.check(lib.mfar_sparse_coo_write(nv.ptr(sv), sv.stride(0), Q, N, nv.ptr(bits), nv.ptr(qids), 0, nv.ptr(offs),
                                           nv.ptr(keys), nv.ptr(vals), nv.F16, nv.stream()))
        ev[3].record()
        kh, vh = keys.cpu(), vals.cpu()
        t_all.append(time.perf_counter() - w0)
        t_sc.append(ev[0].elapsed_time(ev[1]))
        t_wr.append(ev[2].elapsed_time(ev[3]))
    # count stage timed separately (its end event above is re-recorded after the read-back)
    for i in range(args.steps):
        sv = fs.field_scores(batches[i % 4], Q)[:, 0, :N]
        offs = torch.empty(lib.mfar_sparse_coo_offsets_len(Q, N), dtype=torch.int64, device=dev)
        ev[0].record()
        nv.check(lib.mfar_sparse_coo_count(nv.ptr(sv), sv.stride(0), Q, N, nv.ptr(bits), 0, nv.ptr(offs), nv.stream()))
        ev[1].record()
        torch.cuda.synchronize()
        t_cnt.append(ev[0].elapsed_time(ev[1]))
    rows_bytes = Q * N * 4
    med = lambda x: float(np.median(x))  # noqa: E731
    # CPU: the reference's procedure over the same rows (score vectors given), a few queries
    svh = sv[: args.cpu_queries].cpu().numpy()
    safe_set = set(safe.tolist())
    t0 = time.perf_counter()
    ck, cv = PO.precompute_score_for_field({i: svh[i] for i in range(args.cpu_queries)}, safe_set)
    t_cpu = (time.perf_counter() - t0) / args.cpu_queries
    gk, gv = rows_to_coo(sv[: args.cpu_queries], N, bits, qids[: args.cpu_queries], 0)
    same = bool(np.array_equal(gk.cpu().numpy(), ck) and np.array_equal(gv.cpu().numpy().view(np.uint16), cv.view(np.uint16)))
    res = {
        "workload": f"precompute BM25 score pairs: 1 sparse field, {N} docs, Q={Q} queries/batch, safe set {args.safe_frac:.0%}",
        "nnz_per_batch": nnz, "scatter_ms": med(t_sc), "count_scan_ms": med(t_cnt), "write_ms": med(t_wr),
        "count_gbs": rows_bytes / (med(t_cnt) * 1e-3) / 1e9, "write_gbs": (rows_bytes + nnz * 10) / (med(t_wr) * 1e-3) / 1e9,
        "batch_wall_ms_incl_d2h": med(t_all) * 1e3, "queries_per_s": Q / med(t_all),
        "cpu_port_ms_per_query": t_cpu * 1e3, "cpu_port_queries_per_s": 1.0 / t_cpu, "cpu_cores": 1,
        "gpu_equals_cpu_port_bitwise": same,
    }
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
