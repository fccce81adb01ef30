"""Condense an ncu launch list (--csv, gpu__time_duration.sum) into per-kernel totals and shares."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    tot[name][0] += 1
    tot[name][1] += ns
total = sum(v[1] for v in tot.values())
print(f"# {path}: {sum(v[0] for v in tot.values())} launches, {total / 1e6:.3f} ms summed kernel time (serialised, cold cache)")
print(f"{'share':>7} {'ms':>9} {'n':>5}  kernel")
for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{ns / total:7.3%} {ns / 1e6:9.3f} {n:5d}  {k[:110]}")
