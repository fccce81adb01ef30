"""Hottest SASS instructions of an `ncu --set full --import-source on` capture (ncu -i X.ncu-rep --page source --csv > src.csv):
share of warp-stall samples per instruction with its dominant stall reason.   python tools/ncu_hot_sass.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[2:]):
    try:
        data.append((float(r[ix["Warp Stall Sampling (All Samples)"]]), n, r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1.0
print(f"# {len(data)} instructions, {tot:.0f} samples")
for v, n, r in sorted(data, key=lambda x: -x[0])[:top]:
    why = max(stalls, key=lambda h: float(r[ix[h]] or 0))
    print(f"{v / tot * 100:5.1f}%  #{n:5d}  {why[6:]:>14s}  exec {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:110]}")
