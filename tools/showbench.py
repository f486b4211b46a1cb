import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("Q=%d: %.1f q/s  %.2f ms  achieved %.1f %s frac %.3f  e2e %.1f clocks %s" % (d["config"]["batch"], d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["unit"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"]))
for b in d["other_batches"]: print("  Q=%d: %.1f q/s kernel %.2f ms  hbm %.0f GB/s (%.3f)  %.0f TF" % (b["batch"], b["value"], b["kernel_ms"], b["hbm_gbs"], b["hbm_frac"], b["tflops"]))
if d.get("cpu_baseline"): print("  cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
