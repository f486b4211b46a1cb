#!/usr/bin/env python
"""Where does a sharded step go?  torchrun --nproc-per-node N tools/dist_step_breakdown.py [--docs D] [--batch Q]

Times, per rank and with CUDA events, the same shard three ways: (a) the local search only (graph replay, no exchange),
(b) local search + complete exchange (graph replay, what bench.py times), (c) the exchange kernel alone on fixed keys.
Prints one JSON line from rank 0 with every rank's numbers."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=0)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=60)
    a = ap.parse_args()
    args = argparse.Namespace(gpus=int(os.environ.get("WORLD_SIZE", "1")), steps=a.steps, warmup=5, impl="ours",
                              workload="scale_10m_all", batch=a.batch, docs=a.docs, kernel="auto", extra_batches="",
                              others="none", sparse_mode="precomputed", graph=True, balance="off",
                              exchange_mode="complete", cpu_budget_s=0.0, seed=1234)
    ctx = bench.Ctx(args)
    wl = bench.Workload(ctx, "scale_10m_all", a.docs)
    pool = wl.make_batches(a.batch, 2)
    from mfar_b200.modeling.retrieval import GraphedSearch
    local = GraphedSearch(wl.retr, a.batch, sparse="none")
    shard = GraphedSearch(wl.retr, a.batch, sparse="none", sharded=wl.sharded)

    def loop(fn, n):
        for i in range(5):
            fn(i)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        e1.synchronize()
        ctx.barrier()
        return ctx.per_rank(e0.elapsed_time(e1) / n)

    out = {"world": ctx.world, "docs_per_rank": wl.hi - wl.lo, "batch": a.batch}
    for rep in range(2):
        out[f"local_only_ms_{rep}"] = loop(lambda i: local(pool[i % 2][0], pool[i % 2][1]), a.steps)
        out[f"with_exchange_ms_{rep}"] = loop(lambda i: shard(pool[i % 2][0], pool[i % 2][1]), a.steps)
    if ctx.exchange is not None:
        _, _, keys = wl.retr.search(pool[0][0], pool[0][1], None, return_keys=True)
        out["exchange_alone_ms"] = loop(lambda i: ctx.exchange.merge(keys, bench.TOPK), 200)
    if ctx.rank == 0:
        print(json.dumps(out))
    ctx.dist.destroy_process_group() if ctx.world > 1 else None


if __name__ == "__main__":
    main()
