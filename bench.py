#!/usr/bin/env python
"""bench.py - queries/sec of exhaustive multi-field top-100 on synthetic STaRK-shaped corpora.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload scale_10m_all] [--batch 512]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU pipeline (oracle port) on host cores

One "step" = one batch of Q queries scored against the whole corpus (all fields), mixed, top-100.
Multi-GPU = strong scaling: the SAME global corpus is doc-range sharded over the ranks, per-shard
top-k keys are all-gathered (NCCL) and merged on every rank.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "multifield-adaptive-retrieval_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "queries/sec for multi-field top-100"
NOMINAL_HBM_GBS = 8000.0        # north_star: "roughly 8 TB/s" (fractions against both, labelled)
NOMINAL_BF16_TFLOPS = 2250.0    # dense bf16, B200 data sheet
DIM = 768
TOPK = 100
GEN_CHUNK = 65536


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ---------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------- corpus
def build_shard(n_total, n_fields, lo, hi, seed, device):
    """Rank-local PackedCorpus holding global docs [lo, hi).  Generation is per global 64k-doc chunk so every
    shard count sees the same global corpus."""
    from mfar_b200 import synth
    from mfar_b200.modeling.retrieval import PackedCorpus
    pc = PackedCorpus(hi - lo, n_fields, DIM, device)
    mu = synth.corpus_mean(DIM, seed, device)
    c0, c1 = lo // GEN_CHUNK, (hi - 1) // GEN_CHUNK
    for c in range(c0, c1 + 1):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + c + 1)
        clo, chi = c * GEN_CHUNK, min(n_total, (c + 1) * GEN_CHUNK)
        a, b = max(lo, clo), min(hi, chi)
        for f in range(n_fields):
            rows = synth.make_field_rows(chi - clo, DIM, mu, g, device)
            pc.load_rows(f, a - lo, rows[a - clo:b - clo])
    torch.cuda.synchronize()
    return pc, mu


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def algorithmic_work(n_docs, n_dense, n_sparse, Q, sparse_bytes=2):
    """bytes = N*Fd*d*2 (+ Q*N*Fs*b_s) + small; flops = 2*Q*N*Fd*d."""
    bytes_ = n_docs * n_dense * DIM * 2 + Q * n_docs * n_sparse * sparse_bytes + Q * DIM * 2 + Q * TOPK * 12
    flops = 2.0 * Q * n_docs * n_dense * DIM
    return bytes_, flops


# ---------------------------------------------------------------------------------------- CPU legs
def cpu_reference_leg(n_total, n_dense, n_sparse, Q, seed, budget_s, steps, warmup):
    """The reference's CPU pipeline (oracle port of trec_eval_step: per-field retrieve_batch -> union -> rescore ->
    mask -> mixture -> top-100, fp32 torch on all host threads) on a bounded doc sample; q/s is extrapolated
    per-doc-linearly to the full corpus and labelled so.  (Inputs are drawn on the GPU when there is one, only to
    make the sample quickly; everything timed runs on the host cores.)"""
    import mfar_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    gdev = "cuda" if torch.cuda.is_available() else "cpu"
    g = torch.Generator(device=gdev).manual_seed(seed)
    mu = torch.randn(DIM, generator=g, device=gdev)

    def draw(*shape):
        return torch.randn(*shape, generator=g, device=gdev)

    q = O.round_bf16((draw(Q, DIM) + 0.5 * mu).cpu())
    W = (0.05 * draw(DIM, n_dense + n_sparse)).cpu()

    def make(n):
        fields = [(draw(n, DIM) + 0.5 * mu).to(torch.bfloat16).float().cpu() for _ in range(n_dense)]
        sp = None
        if n_sparse:
            u = torch.rand(Q, n_sparse, n, generator=g, device=gdev)
            sp = torch.where(u < 0.95, torch.zeros((), device=gdev),
                             4.0 * torch.rand(Q, n_sparse, n, generator=g, device=gdev)).cpu()
        return fields, sp

    def run_once(fields, sp):
        t0 = time.perf_counter()
        O.union_rescore(q, fields, sp, q, W, True, None, TOPK)
        return time.perf_counter() - t0

    def run_exhaustive(fields, sp):
        t0 = time.perf_counter()
        O.exhaustive_topk(q, fields, sp, O.mixture_weights(q, W, True), None, TOPK)
        return time.perf_counter() - t0

    # two-point calibration t(n) = a + b*n, then size the sample to the time budget
    n1, n2 = min(n_total, 4096), min(n_total, 32768)
    f1, s1 = make(n1)
    run_once(f1, s1)
    t1 = run_once(f1, s1)
    f2, s2 = make(n2)
    t2 = run_once(f2, s2)
    b = max((t2 - t1) / max(1, n2 - n1), 1e-9)
    a = max(t1 - b * n1, 0.0)
    per_call = max(budget_s / max(1, steps + warmup + 1), 0.5)
    n_sample = int((per_call - a) / b) if per_call > a else n2
    ram_cap = int(16e9 / (max(1, n_dense) * DIM * 4 + Q * n_sparse * 4))
    n_sample = max(TOPK, min(n_sample, ram_cap, n_total))
    fields, sp = make(n_sample)
    for _ in range(warmup):
        run_once(fields, sp)
    times = [run_once(fields, sp) for _ in range(steps)]
    t_ex = run_exhaustive(fields, sp)
    scale = n_total / n_sample
    t_step = statistics.median(times)
    return {
        "value": Q / (t_step * scale), "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
        "sample": (f"oracle union_rescore (faithful trec_eval_step port), fp32 torch CPU, {n_sample} of {n_total} docs x "
                   f"{n_dense}+{n_sparse} fields, Q={Q}, median of {steps}; extrapolated per-doc-linearly x{scale:.1f}"),
        "exhaustive_value": Q / (t_ex * scale), "ms_per_step_sample": t_step * 1e3, "n_sample": n_sample,
        "os_cpu_count": os.cpu_count(), "cpu_model": cpu_model(),
    }, t_step * scale


# ---------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="scale_10m_all")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--docs", type=int, default=0, help="override the workload's doc count (debug)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tcgen05", "tcgen05_qs"])
    ap.add_argument("--extra-batches", default="1,64", help="also measured on the device-resident path at N=1")
    ap.add_argument("--sparse-mode", default="precomputed", choices=["precomputed", "bm25"],
                    help="sparse fields as precomputed [Q,Fs,N] f16 score tensors (north_star (2)) or scored on the "
                         "device from query tokens against HBM-resident BM25 postings (SURVEY 8f-3)")
    ap.add_argument("--graph", action="store_true",
                    help="N=1: replay the device-resident step as one CUDA graph (GraphedSearch); the scoring kernel's "
                         "own duration for the roofline is taken from a short eager pass")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0)
    ap.add_argument("--seed", type=int, default=1234)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    from mfar_b200 import synth
    n_total, n_dense, n_sparse = synth.SHAPES[args.workload]
    if args.docs:
        n_total = args.docs
    Q = args.batch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": f"{args.workload}: {n_total} docs x {n_dense} dense + {n_sparse} sparse fields x {DIM}-d bf16, "
                          f"exhaustive hybrid top-{TOPK}, query-conditioned mixture",
              "n_docs": n_total, "n_dense": n_dense, "n_sparse": n_sparse, "dim": DIM, "batch": Q, "top_k": TOPK,
              "sharding": f"doc-range x{world}"}
    shard_mb = (n_total // world) * max(n_dense, 1) * DIM * 2 / 1e6
    config["cache"] = (f"corpus shard {shard_mb:.0f} MB >> 126 MB L2 (inputs larger than L2)" if shard_mb > 2 * 126 else
                       f"corpus shard {shard_mb:.0f} MB is NOT larger than the 126 MB L2 and is not flushed between steps: "
                       "small-workload numbers are L2-assisted")
    bm25_mode = args.sparse_mode == "bm25" and n_sparse > 0
    if n_sparse:
        config["sparse_input"] = ("device BM25: query tokens -> postings scatter-add (mfar_score_topk_bm25)" if bm25_mode
                                  else "precomputed [Q,Fs,N] f16 score tensor")

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        cb, t_full = cpu_reference_leg(n_total, n_dense, n_sparse, Q, args.seed, max(args.cpu_budget_s, 20.0) * 3,
                                       args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "queries/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_full * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    # Native libraries (NCCL's version banner, ...) write to fd 1; the contract is ONE JSON line on stdout, so fd 1
    # points at stderr until that line is printed.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA sm_100 device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from mfar_b200 import _native as nv
    from mfar_b200.dist import ShardedRetriever, shard_range
    from mfar_b200.modeling.retrieval import MultiFieldRetriever
    from mfar_b200.modeling.weighting import LinearWeights

    lo, hi = shard_range(n_total, rank, world)
    t_setup = time.perf_counter()
    if n_dense:
        pc, mu = build_shard(n_total, n_dense, lo, hi, args.seed, device)
    else:                                             # sparse-only scorer: nothing to pack
        pc, mu = None, synth.corpus_mean(DIM, args.seed, device)
    F = n_dense + n_sparse
    layer = LinearWeights(DIM, F, query_cond=True)
    with torch.no_grad():
        layer.weight.copy_(synth.make_mixture(DIM, F, args.seed + 1))
    bm25_fields = None
    if bm25_mode:
        bm25_fields = [synth.make_bm25_field(n_total, args.seed + 300 + j, device, doc_range=(lo, hi))
                       for j in range(n_sparse)]
        torch.cuda.synchronize()
    retr = MultiFieldRetriever(pc, layer.to(device), n_sparse=n_sparse, top_k=TOPK, doc_id_base=lo, impl=args.kernel,
                               sparse_indices=bm25_fields, n_docs=hi - lo, device=device)
    exchange, exchange_kind = None, "none (1 GPU)"
    if world > 1:
        from mfar_b200.dist import PeerExchange
        if os.environ.get("MFAR_EXCHANGE", "p2p") == "p2p":
            try:
                exchange = PeerExchange(q_cap=max(1024, Q), k_cap=128, device=device)
                exchange_kind = "fused NVLink peer-memory exchange+merge kernel (mfar_topk_exchange_merge)"
            except Exception as e:  # noqa: BLE001  (symmetric memory unavailable: NCCL all-gather + merge kernel)
                exchange_kind = f"nccl all_gather + merge kernel (peer exchange unavailable: {type(e).__name__}: {e})"[:300]
        else:
            exchange_kind = "nccl all_gather + merge kernel"
    sharded = ShardedRetriever(retr, exchange=exchange)
    setup_s = time.perf_counter() - t_setup

    def make_batches(q_count, n_pool=4):
        pool = []
        for i in range(n_pool):
            qv = synth.make_queries(q_count, DIM, mu, args.seed + 100 + i, device)
            if bm25_mode:
                sp, ent = None, synth.make_bm25_query_entries(q_count, n_sparse, args.seed + 200 + i).to(device)
            else:
                sp, ent = synth.make_sparse(q_count, n_sparse, hi - lo, args.seed + 200 + i, device, pitch=64), None
            pool.append((qv, qv.float(), sp, ent))
        return pool

    def batch_postings(ent):
        """postings the batch touches on this shard (for the algorithmic byte count)"""
        total = 0
        for j, f in enumerate(bm25_fields):
            t = ent[ent[:, 1] == j][:, 2].long()
            ip = f.scores["indptr"]
            total += int((ip[t + 1] - ip[t]).sum().item())
        return total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphs = {}

    def step(batch):
        qv, qe, sp, ent = batch
        if args.graph and world == 1:
            gs = graphs.get(qv.shape[0])
            if gs is None:
                from mfar_b200.modeling.retrieval import GraphedSearch
                gs = graphs[qv.shape[0]] = GraphedSearch(
                    retr, qv.shape[0], sparse="bm25" if bm25_mode else ("dense" if n_sparse else "none"),
                    max_entries=0 if ent is None else ent.shape[0], sparse_ld=None if sp is None else sp.shape[2])
            gs(qv, qe, sparse=sp, entries=ent)
            return gs.launches
        sharded.search(qv, qe, sp, sparse_tokens=ent)
        return retr.last_launches + 1 + 1                   # + mixture-weights kernel + cross-shard merge kernel

    def run_device(pool, steps, warmup, profile=False):
        if args.graph and world == 1 and profile:           # kernel duration from a short eager pass
            nv.check(nv.lib().mfar_profile_enable(1))
            for i in range(3):
                qv, qe, sp, ent = pool[i % len(pool)]
                sharded.search(qv, qe, sp, sparse_tokens=ent)
            torch.cuda.synchronize()
            buf0 = (ctypes.c_float * 256)()
            n0 = nv.lib().mfar_profile_collect(ctypes.addressof(buf0), 256)
            eager_kern_ms = [buf0[i] for i in range(max(n0, 0))]
            nv.lib().mfar_profile_enable(0)
            profile = False
        else:
            eager_kern_ms = None
        for i in range(warmup):
            step(pool[i % len(pool)])
        barrier()
        if profile:
            nv.check(nv.lib().mfar_profile_enable(1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = 0
        ncu_range = profile and os.environ.get("MFAR_NCU_RANGE") == "1"   # ncu --profile-from-start off
        if ncu_range:
            torch.cuda.profiler.start()
        e0.record()
        for i in range(steps):
            launches += step(pool[i % len(pool)])
        e1.record()
        barrier()
        if ncu_range:
            torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        kern_ms = []
        if profile:
            buf = (ctypes.c_float * 256)()
            n = nv.lib().mfar_profile_collect(ctypes.addressof(buf), 256)
            kern_ms = [buf[i] for i in range(max(n, 0))]
            nv.lib().mfar_profile_enable(0)
        if eager_kern_ms is not None:
            kern_ms = eager_kern_ms
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, launches, kern_ms

    def run_e2e(pool_host, steps, warmup):
        """Through the C-ABI host-buffer call: pinned host inputs -> H2D -> mixture + scoring + top-k -> D2H."""
        out_s = torch.empty((Q, TOPK), dtype=torch.float32).pin_memory()
        out_i = torch.empty((Q, TOPK), dtype=torch.int64).pin_memory()

        def one(i):
            qh, qeh, sph, enth = pool_host[i % len(pool_host)]
            if bm25_mode:
                retr.search_host_bm25(qh, qeh, enth, out_scores=out_s, out_ids=out_i)
            else:
                retr.search_host(qh, qeh, sph, out_scores=out_s, out_ids=out_i)
        for i in range(warmup):
            one(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one(i)
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    peaks = load_peaks()
    pool_dev = make_batches(Q)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches, kern_ms = run_device(pool_dev, args.steps, args.warmup, profile=True)
    clocks = sampler.stop() if rank == 0 else None

    # e2e: N=1 goes through mfar_search_host; N>1 adds the (device) key exchange + merge per step
    pool_host = [(qv.cpu().pin_memory(), qe.cpu().pin_memory(),
                  None if sp is None else sp[:, :, :hi - lo].contiguous().cpu().pin_memory(),   # host call: pitch = N
                  None if ent is None else ent.cpu().pin_memory())
                 for qv, qe, sp, ent in pool_dev]
    if world == 1:
        ms_e2e = run_e2e(pool_host, args.steps, args.warmup)
    else:
        def e2e_multi():
            for i in range(args.warmup):
                one_e2e(i)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                one_e2e(i)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()

        def one_e2e(i):
            qh, qeh, sph, enth = pool_host[i % len(pool_host)]
            qv = qh.to(device, non_blocking=True)
            qe = qeh.to(device, non_blocking=True)
            sp = None if sph is None else sph.to(device, non_blocking=True)
            ent = None if enth is None else enth.to(device, non_blocking=True)
            s, ids = sharded.search(qv, qe, sp, sparse_tokens=ent)
            s.cpu(); ids.cpu()
        ms_e2e = e2e_multi()
    if bm25_mode:
        h2d = Q * DIM * 2 + Q * DIM * 4 + pool_dev[0][3].shape[0] * 12
    else:
        h2d = Q * DIM * 2 + Q * DIM * 4 + (Q * n_sparse * (hi - lo) * 2 if n_sparse else 0)
    d2h = Q * TOPK * 12

    # roofline of the dominant kernel (the fused scoring kernel), per launch, this rank's shard
    n_shard = hi - lo
    a_bytes, a_flops = algorithmic_work(n_shard, n_dense, 0 if bm25_mode else n_sparse, Q)
    sparse_stage = None
    if bm25_mode:
        # per batch: 8 B read per posting touched + the fp32 base[Q,N] row block zeroed, accumulated (L2 atomics) and
        # read back once by the scoring epilogue
        postings = statistics.mean(batch_postings(b[3]) for b in pool_dev)
        sparse_bytes = postings * 8 + 2 * Q * n_shard * 4
        a_bytes += sparse_bytes
        sparse_stage = {"postings_per_batch": postings, "algorithmic_bytes": sparse_bytes,
                        "entries_per_batch": int(pool_dev[0][3].shape[0])}
    if n_dense == 0:
        # sparse-only scorer: the timed kernel is the streaming top-k over the fp32 base[Q,N] rows - its own algorithmic
        # traffic is that block read once (the pre-mix / BM25 scatter that wrote it are separate launches)
        a_bytes = Q * n_shard * 4 + Q * TOPK * 12
    k_ms = statistics.mean(kern_ms) if kern_ms else None
    hbm_bound = Q < 200 or n_dense == 0
    if k_ms:
        if hbm_bound:
            achieved = a_bytes / (k_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "frac_of_nominal_peak": achieved / NOMINAL_HBM_GBS,
                    "nominal_peak": NOMINAL_HBM_GBS}
        else:
            achieved = a_flops / (k_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["bf16_tflops_sustained"],
                    "frac_of_burst_peak": achieved / peaks["bf16_tflops"],
                    "frac_of_nominal_peak": achieved / NOMINAL_BF16_TFLOPS, "nominal_peak": NOMINAL_BF16_TFLOPS,
                    "hbm_gbs_same_launch": a_bytes / (k_ms * 1e-3) / 1e9}
        kname = {"simt": "score_simt_kernel", "tcgen05": "score_tc_kernel", "tcgen05_qs": "score_qs_kernel"}.get(
            args.kernel, "score_qs_kernel" if Q > 64 else "score_tc_kernel")
        if n_dense == 0:
            kname = "topk_rows_kernel"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath) and world == 1 and not args.docs:
            traffic = json.load(open(tpath)).get(f"{args.workload}|{Q}")
        roof.update({"traffic": traffic, "traffic_source": "ncu --set full capture, profiles/ncu_traffic.json" if traffic else None,
                     "kernel": kname,
                     "kernel_ms": k_ms,
                     "kernel_share_of_step": (k_ms * args.steps if args.graph else k_ms * len(kern_ms)) / ms_total
                     if ms_total else None,
                     "algorithmic_bytes_per_launch": a_bytes, "algorithmic_flops_per_launch": a_flops,
                     "peak_source": peaks["source"] + (" (sustained: kernel timed inside a long step)" if not hbm_bound else " copy bandwidth")})
    else:
        roof = None

    extra = []
    if world == 1 and rank == 0 and args.extra_batches:
        for qb in [int(x) for x in args.extra_batches.split(",") if x]:
            if qb == Q:
                continue
            pool = make_batches(qb, 2)
            ms_b, _, km = run_device(pool, max(5, args.steps // 2), 3, profile=True)
            steps_b = max(5, args.steps // 2)
            bb, ff = algorithmic_work(n_shard, n_dense, 0 if bm25_mode else n_sparse, qb)
            kk = statistics.mean(km) if km else None
            extra.append({"batch": qb, "value": qb * steps_b / (ms_b * 1e-3), "ms_per_step": ms_b / steps_b,
                          "kernel_ms": kk,
                          "hbm_gbs": bb / (kk * 1e-3) / 1e9 if kk else None,
                          "hbm_frac": bb / (kk * 1e-3) / 1e9 / peaks["hbm_gbs"] if kk else None,
                          "tflops": ff / (kk * 1e-3) / 1e12 if kk else None})
            del pool

    cpu_base = None
    if rank == 0 and world == 1 and args.cpu_budget_s > 0:
        cpu_base, _ = cpu_reference_leg(n_total, n_dense, n_sparse, Q, args.seed, args.cpu_budget_s, 3, 1)

    if rank == 0:
        line = {
            "metric": METRIC, "value": Q * args.steps / (ms_total * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "roofline": roof, "cpu_baseline": cpu_base,
            "e2e": {"value": Q * args.steps / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "path": "mfar_search_host (C ABI, pinned host buffers)" if world == 1 else
                            "pinned host -> device copies + sharded search + D2H of the merged top-k"},
            "sparse_stage": sparse_stage, "gpu_launches": launches, "exchange": exchange_kind, "clocks": clocks, "other_batches": extra, "setup_s": setup_s,
            "kernel_impl": args.kernel, "cuda_graph": bool(args.graph and world == 1),
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
