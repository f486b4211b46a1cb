#!/usr/bin/env python
"""Readable summary of a bench.py JSON line:  python tools/showbench.py <file>"""
import json
import sys

d = json.load(open(sys.argv[1]))
r = d.get("roofline") or {}


def row(b, indent="  "):
    f = b.get("frac")
    print(f"{indent}Q={b['batch']:<4d} {b['value']:>10.1f} q/s  {b['ms_per_step']:8.3f} ms/step  {str(b.get('kernel')):<17s} "
          f"{(b.get('kernel_ms') or 0):8.3f} ms  {b.get('bound')}: {(b.get('achieved') or 0):7.1f} {b.get('unit')} "
          f"frac {f if f is None else round(f, 3)}  share {round(b.get('kernel_share_of_step') or 0, 3)}")


print(f"{d['config']['workload'][:60]}  N={d['n_gpus']}  [{d['config'].get('sharding')}]")
print(f"  Q={d['config']['batch']}: {d['value']:.1f} q/s  {d['ms_per_step']:.3f} ms/step  kernel {r.get('kernel_ms', 0):.3f} ms  "
      f"{r.get('achieved', 0):.1f} {r.get('unit')} frac {r.get('frac', 0):.3f} (burst {r.get('frac_of_burst_peak')})  "
      f"share {r.get('kernel_share_of_step', 0):.3f}  e2e {d['e2e']['value']:.1f} q/s")
print(f"  clocks {d.get('clocks')}\n  parity {d.get('parity_check')}")
cb = d.get("cpu_baseline")
if cb:
    print(f"  cpu_baseline {cb['value']:.3f} q/s  kind={cb['kind']} cores={cb['cores']}  a={cb.get('fit_a_s')} b={cb.get('fit_b_s_per_doc')}")
for b in d.get("other_batches", []):
    row(b)
for o in d.get("other_workloads", []):
    if "error" in o:
        print(f"  {o['workload']}: ERROR {o['error']}")
        continue
    pc = o.get("parity_check") or {}
    print(f"  {o['workload']} ({o['n_docs']} docs, {o['n_dense']}+{o['n_sparse']} fields, shard {o['shard_docs']})  parity ok={pc.get('ok')} q={pc.get('queries')} {pc.get('error', '')}")
    for b in o["batches"]:
        row(b, "    ")
    u = o.get("union_rescore")
    if u:
        print("    union_rescore:", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in u.items() if k != "mode"})
