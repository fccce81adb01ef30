import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", round(d["e2e"]["value"], 1), "clocks", d.get("clocks"))
print("roofline", json.dumps(d["roofline"]))
if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for k, v in (d.get("kernel_breakdown") or {}).items():
    print(f"{k:22s} n={v['launches_per_step']:3d} ms={v['ms_per_step']:.3f} share={v['share_of_uno_kernels']:.3f} GB/s={v['GBps']:.0f} TF={v['TFLOPs']:.1f}")
