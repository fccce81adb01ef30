"""Per-role DRAM traffic from an ncu capture of one step (tools/profile_step.py).  Two input formats:

    # (a) metric capture written as a CSV log (one row per launch and metric) -- small enough for every launch of a step
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/traffic.csv python tools/profile_step.py
    # (b) the raw page of a --set full report:  ncu -i <rep>.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_traffic.py <csv> profiles/r01_traffic.json

Writes {role: {"launches": n, "dram_bytes_per_launch": mean(read+write), "ncu_time_us_per_launch": ...}}; bench.py
reports the entry of the dominant role as roofline.traffic (per launch, like roofline.achieved)."""
import csv
import json
import os
import sys

# kernel-name fragment -> role tag used by the library's per-launch profiler (backend_cuda.cu ProfScope tags)
ROLES = [
    ("resample2d_kernel", "resample_banded"), ("banded2d_kernel", "resample_banded"), ("banded_kernel", "resample_banded"),
    ("kpipe_kernel", "dft_last_analysis"), ("rowgemm_smallk_kernel", "dft_last_synthesis"),
    ("conv1x1_tc_kernel", "conv1x1"), ("wgrad_tc_kernel", "conv1x1_wgrad"), ("mid2_kernel", "dft_mid"), ("mid_tc_kernel", "dft_mid"), ("cmm_kernel", "mode_contraction"),
    ("cmm_tc_kernel", "mode_contraction"), ("cmm_tc4_kernel", "mode_contraction"),
    ("proj_bwd_tcp_kernel", "project_bwd"), ("proj_fwd_tc_kernel", "project_fwd"),
    ("cmm2_kernel", "mode_contraction"), ("norm_fwd_cluster_kernel", "instnorm_gelu_fwd"), ("norm_bwd_cluster_kernel", "instnorm_gelu_bwd"),
    ("lp_partial_kernel", "lp_loss"), ("lp_finish_kernel", "lp_loss"), ("lp_bwd_kernel", "lp_loss_bwd"),
    ("proj_bwd_kernel", "project_bwd"), ("proj_fwd_kernel", "project_fwd"), ("lift_bwd_kernel", "lift_bwd"), ("lift_fwd_kernel", "lift_fwd"),
    ("gelu_bwd_bias_kernel", "gelu_bwd"), ("gelu_bwd_kernel", "gelu_bwd"), ("norm_act_bwd_kernel", "instnorm_gelu_bwd"),
    ("norm_act_fwd_kernel", "instnorm_gelu_fwd"), ("plane_stats_kernel", "instnorm_stats"), ("channel_sum_kernel", "bias_grad"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def launches(src):
    """yield (kernel name, {metric: value in bytes / us}) per launch for either input format"""
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    if "Metric Name" in ix:                      # (a) long format
        cur, acc = None, None
        for r in rows[1:]:
            if len(r) <= ix["Metric Value"]:
                continue
            if r[ix["ID"]] != cur:
                if acc is not None:
                    yield acc
                cur, acc = r[ix["ID"]], (r[ix["Kernel Name"]], {})
            acc[1][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * UNIT.get(r[ix["Metric Unit"]], 1.0)
        if acc is not None:
            yield acc
    else:                                        # (b) raw page: header row, unit row, one row per launch
        units = rows[1]
        for r in rows[2:]:
            yield r[ix["Kernel Name"]], {m: float(r[ix[m]].replace(",", "")) * UNIT.get(units[ix[m]], 1.0)
                                         for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}


def main(src, dst):
    out = {}
    for name, m in launches(src):
        role = next((tag for frag, tag in ROLES if frag in name), None)
        if role is None:
            continue
        e = out.setdefault(role, {"launches": 0, "bytes": 0.0, "us": 0.0})
        e["launches"] += 1
        e["bytes"] += m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
        e["us"] += m["gpu__time_duration.sum"]
    res = {k: {"launches": v["launches"], "dram_bytes_per_launch": v["bytes"] / v["launches"], "ncu_time_us_per_launch": v["us"] / v["launches"]}
           for k, v in sorted(out.items())}
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench   # the capture is tied to the device code it was taken on (bench.py drops it when the sources changed)

    json.dump({"source": os.path.basename(src), "src_sha16": bench.source_sha16(), "note": "ncu --clock-control none, one step of tools/profile_step.py (cold-cache, serialised launches)", "roles": res},
              open(dst, "w"), indent=1)
    for k, v in res.items():
        print(f"{k:22s} n={v['launches']:3d}  {v['dram_bytes_per_launch'] / 1e6:10.1f} MB/launch  {v['ncu_time_us_per_launch']:9.1f} us")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
