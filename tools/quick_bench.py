#!/usr/bin/env python
"""Developer timing loop: ONE corpus build, several batch sizes / kernel choices, device-resident inputs.

    python tools/quick_bench.py --workload amazon_full --batches 64,128,512 [--impls auto] [--docs N] [--iters 10]

Prints one JSON line per (batch, impl): step ms (CUDA events around the whole search call, mean over --iters after 3
warm-ups) and the scoring kernel's own ms (the library's event pair).  Not a bench.py replacement: no e2e, no clocks,
no roofline bookkeeping - it exists so that an A/B of a kernel change costs seconds of GPU time, not a corpus rebuild
per data point.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multifield-adaptive-retrieval_b200"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="amazon_full")
    ap.add_argument("--batches", default="64,128,512")
    ap.add_argument("--impls", default="auto")
    ap.add_argument("--docs", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--tag", default="")
    ap.add_argument("--sparse-fields", type=int, default=-1, help="override the workload's sparse field count")
    ap.add_argument("--dense-fields", type=int, default=-1, help="override the workload's dense field count")
    args = ap.parse_args()
    from mfar_b200 import _native as nv
    from mfar_b200 import synth
    from mfar_b200.modeling.retrieval import MultiFieldRetriever, PackedCorpus
    from mfar_b200.modeling.weighting import LinearWeights
    n, Fd, Fs = synth.SHAPES[args.workload]
    n = args.docs or n
    Fs = Fs if args.sparse_fields < 0 else args.sparse_fields
    Fd = Fd if args.dense_fields < 0 else args.dense_fields
    dev = torch.device("cuda", 0)
    pc = None
    if Fd:
        pc = PackedCorpus(n, Fd, 768, dev)
        synth.fill_packed_corpus(pc, seed=args.seed)
    mu = synth.corpus_mean(768, args.seed, dev)
    layer = LinearWeights(768, Fd + Fs, query_cond=True)
    with torch.no_grad():
        layer.weight.copy_(synth.make_mixture(768, Fd + Fs, args.seed + 1))
    r = MultiFieldRetriever(pc, layer.to(dev), n_sparse=Fs, top_k=100, n_docs=n, device=dev)
    batches = [int(x) for x in args.batches.split(",") if x]
    qmax = max(batches)
    q_all = synth.make_queries(qmax, 768, mu, args.seed + 100, dev)
    sp_all = None
    if Fs:
        ld = (n + 63) // 64 * 64
        sp_all = torch.zeros((qmax, Fs, ld), dtype=torch.float16, device=dev)
        for q0 in range(0, qmax, 32):
            q1 = min(qmax, q0 + 32)
            sp_all[q0:q1, :, :n] = synth.make_sparse(q1 - q0, Fs, n, args.seed + 200 + q0, dev)
    for Q in batches:
        q, sp = q_all[:Q], (None if sp_all is None else sp_all[:Q].contiguous())
        qe = q.float()
        for impl in args.impls.split(","):
            for _ in range(3):
                r.search(q, qe, sp, impl=impl)
            torch.cuda.synchronize()
            nv.check(nv.lib().mfar_profile_enable(1))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                r.search(q, qe, sp, impl=impl)
            e1.record()
            torch.cuda.synchronize()
            buf = (ctypes.c_float * 256)()
            k = nv.lib().mfar_profile_collect(ctypes.addressof(buf), 256)
            nv.lib().mfar_profile_enable(0)
            km = sum(buf[i] for i in range(k)) / max(k, 1)
            step = e0.elapsed_time(e1) / args.iters
            flops = 2.0 * Q * n * Fd * 768
            byts = n * Fd * 768 * 2 + Q * n * Fs * 2
            print(json.dumps({"tag": args.tag, "workload": args.workload, "docs": n, "batch": Q, "impl": impl,
                              "step_ms": round(step, 4), "kernel_ms": round(km, 4), "launches": r.last_launches,
                              "tflops": round(flops / km / 1e9, 1) if km else None,
                              "gbs": round(byts / km / 1e6, 1) if km else None}), flush=True)


if __name__ == "__main__":
    main()
