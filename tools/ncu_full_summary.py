"""Condense the raw page of an `ncu --set full` report (ncu -i X.ncu-rep --page raw --csv) into one block per launch:
duration, DRAM bytes and achieved fraction of the measured HBM peak, occupancy, issue activity, pipe utilisation (LSU, FMA,
tensor), L2 hit rate and the top warp-stall reasons.   python tools/ncu_full_summary.py raw.csv > profiles/r02_ncu_full_*.txt"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
U = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def val(r, name):
    if name not in ix:
        return None
    try:
        return float(r[ix[name]].replace(",", "")) * U.get(units[ix[name]], 1.0)
    except ValueError:
        return None


STALLS = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print(f"# {os.path.basename(sys.argv[1])}: ncu --set full --clock-control none; HBM peak {PEAK:.0f} GB/s (MEASURED_PEAKS.json)")
for r in rows[2:]:
    t = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum") or 0, val(r, "dram__bytes_write.sum") or 0
    name = r[ix["Kernel Name"]][:90]
    grid = r[ix["Grid Size"]] if "Grid Size" in ix else "?"
    blk = r[ix["Block Size"]] if "Block Size" in ix else "?"
    print(f"\n{name}  grid {grid} block {blk}")
    if t:
        print(f"  {t:9.1f} us   DRAM read {rd / 1e6:8.1f} MB  write {wr / 1e6:8.1f} MB  -> {(rd + wr) / t / 1e3:7.0f} GB/s = {(rd + wr) / t / 1e3 / PEAK:.2f} of HBM peak")
    def pct(label, key):
        v = val(r, key)
        return f"{label} {v:.1f}" if v is not None else None
    parts = [pct("regs", "launch__registers_per_thread"), pct("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
             pct("issue active %", "smsp__issue_active.avg.pct"), pct("LSU %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
             pct("FMA %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
             pct("tensor %", "sm__pipe_tensor_subpipe_all_cycles_active.avg.pct_of_peak_sustained_active") or pct("tensor %", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"),
             pct("L2 hit %", "lts__t_sector_hit_rate.pct"), pct("L1 hit %", "l1tex__t_sector_hit_rate.pct"),
             pct("DRAM %", "dram__throughput.avg.pct_of_peak_sustained_elapsed")]
    print("  " + "  ".join(p for p in parts if p))
    st = sorted(((val(r, h) or 0.0, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in STALLS), reverse=True)[:5]
    print("  stalls per issue: " + ", ".join(f"{n} {v:.2f}" for v, n in st))
