"""Condensed view of a bench.py JSON line (last line of the file given)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])


def head(tag, r):
    e = r.get("e2e") or {}
    print(f"== {tag}: {r.get('value', 0):.1f} samples/s  {r.get('ms_per_step', 0):.3f} ms/step  e2e {e.get('value', 0):.1f}  launches {r.get('gpu_launches')}"
          f"  exec {r.get('execution')} {r.get('graph_error') or ''}  batch {r.get('per_gpu_batch', (r.get('config') or {}).get('per_gpu_batch'))}")
    if r.get("roofline"):
        ro = r["roofline"]
        print(f"   roofline {ro['kernel']} {ro['achieved']:.0f} GB/s frac {ro['frac']:.3f} traffic {ro.get('traffic')}  uno kernels {ro.get('uno_kernel_ms_per_step', 0):.3f} ms")
    for k, v in list((r.get("kernel_breakdown") or {}).items())[:14]:
        print(f"   {k:22s} n={v['launches_per_step']:3d} ms={v['ms_per_step']:.3f} share={v['share_of_uno_kernels']:.3f} GB/s={v['GBps'] or 0:.0f} frac={v.get('hbm_frac') or 0:.2f}")
    for l in (r.get("spectral_levels") or []):
        if isinstance(l, dict) and "level" in l:
            print(f"   {l['level']:75s} {l['ms_per_call']:.3f} ms n={l['launches_per_call']:.0f} hbm {l['hbm_frac']:.3f} tc {l['tensor_frac']:.4f}")


head("headline " + (d.get("config") or {}).get("workload", ""), d)
print("clocks", d.get("clocks"), "cpu_baseline", d.get("cpu_baseline"))
for key in ("strong", "weak"):
    if d.get(key):
        print(key, {k: d[key][k] for k in ("per_gpu_batch", "value", "ms_per_step")})
for wl, r in (d.get("secondary") or {}).items():
    if "value" in r:
        head("secondary " + wl, r)
        print("   cpu", r.get("cpu_baseline"))
    else:
        for mode, rr in r.items():
            if isinstance(rr, dict) and "value" in rr:
                head(f"secondary {wl} [{mode}]", rr)
            else:
                print("secondary", wl, mode, rr)
print("reference_gpu", json.dumps(d.get("reference_gpu")))
sw = d.get("sweep") or {}
for r in sw.get("rows", []):
    print(f"   sweep S={r['in'][0]:3d} m={r['modes'][0]:2d} C={r['Ci']:3d} B={r['B']:4d} fwd {r['fwd_ms']:.3f} ms {r['fwd_frac']:.3f}  bwd {r['bwd_ms']:.3f} ms {r['bwd_frac']:.3f}  vs stock {r['speedup_fwd']:.2f}x / {r['speedup_bwd']:.2f}x")
if "error" in sw:
    print("sweep error", sw["error"])
