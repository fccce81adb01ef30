"""Per-role DRAM traffic from an `ncu --set full` capture of one step (tools/profile_step.py).

    ncu -i gpurun_out/<rep>.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_traffic.py raw.csv profiles/r01_traffic.json

Writes {role: {"launches": n, "dram_bytes_per_launch": mean(read+write), "ncu_time_us_per_launch": ...}}; bench.py
reports the entry of the dominant role as roofline.traffic (per launch, like roofline.achieved)."""
import csv
import json
import sys

# kernel-name fragment -> role tag used by the library's per-launch profiler (backend_cuda.cu ProfScope tags)
ROLES = [
    ("resample2d_kernel", "resample_banded"), ("banded2d_kernel", "resample_banded"), ("banded_kernel", "resample_banded"),
    ("kpipe_kernel", "dft_last_analysis"), ("rowgemm_smallk_kernel", "dft_last_synthesis"),
    ("conv1x1_tc_kernel", "conv1x1"), ("wgrad_tc_kernel", "conv1x1_wgrad"), ("mid2_kernel", "dft_mid"), ("cmm_kernel", "mode_contraction"),
    ("proj_bwd_kernel", "project_bwd"), ("proj_fwd_kernel", "project_fwd"), ("lift_bwd_kernel", "lift_bwd"), ("lift_fwd_kernel", "lift_fwd"),
    ("gelu_bwd_bias_kernel", "gelu_bwd"), ("gelu_bwd_kernel", "gelu_bwd"), ("norm_act_bwd_kernel", "instnorm_gelu_bwd"),
    ("norm_act_fwd_kernel", "instnorm_gelu_fwd"), ("plane_stats_kernel", "instnorm_stats"), ("channel_sum_kernel", "bias_grad"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        role = next((tag for frag, tag in ROLES if frag in name), None)
        if role is None:
            continue

        def val(m):
            return float(r[ix[m]].replace(",", "")) * UNIT.get(units[ix[m]], 1.0)

        e = out.setdefault(role, {"launches": 0, "bytes": 0.0, "us": 0.0})
        e["launches"] += 1
        e["bytes"] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        e["us"] += val("gpu__time_duration.sum")
    res = {k: {"launches": v["launches"], "dram_bytes_per_launch": v["bytes"] / v["launches"], "ncu_time_us_per_launch": v["us"] / v["launches"]}
           for k, v in sorted(out.items())}
    json.dump({"source": src, "note": "ncu --set full --clock-control none, one step of tools/profile_step.py (cold-cache, serialised)", "roles": res},
              open(dst, "w"), indent=1)
    for k, v in res.items():
        print(f"{k:22s} n={v['launches']:3d}  {v['dram_bytes_per_launch'] / 1e6:10.1f} MB/launch  {v['ncu_time_us_per_launch']:9.1f} us")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
