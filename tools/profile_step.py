"""One profiled training step for ncu: everything before the step is outside the cudaProfiler range.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [--workload darcy] [--batch 32]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from uno_b200.losses import LpLoss  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="darcy")
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()
cls, cargs, ckw, xshape, tshape, def_b, _ = bench.WORKLOADS[args.workload]
B = args.batch or def_b
model = bench.build_model(args.workload)
step = bench.make_step(model, LpLoss(size_average=False), B, tshape, ar_steps=bench.AR_STEPS.get(args.workload, 0))
torch.manual_seed(1)
x = torch.randn(B, *xshape, device="cuda")
y = torch.randn(B, *tshape, device="cuda")
for _ in range(args.warmup):
    step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
