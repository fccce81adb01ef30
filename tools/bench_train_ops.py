"""Timing of the training-step ops next to the hot path on the Darcy model's parameter set:
the fused multi-tensor Adam (one launch per 24 tensors) against the reference's per-tensor update written with the same
torch ops Adam.py:23-52 issues, and the fused relative-L2 loss against the torch expression of utilities3.py:86-100.

    python tools/bench_train_ops.py > gpurun_out/train_ops.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import math  # noqa: E402

import torch  # noqa: E402

import bench  # noqa: E402
from uno_b200.losses import LpLoss  # noqa: E402
from uno_b200.optim import Adam  # noqa: E402


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def reference_style_adam(params, states, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    # the op sequence of Adam.py:27-52, one tensor at a time
    bc1, bc2 = 1 - b1**step, 1 - b2**step
    for p, st in zip(params, states):
        g = p.grad
        st["m"].mul_(b1).add_(g, alpha=1 - b1)
        st["v"].mul_(b2).addcmul_(g, g.conj(), value=1 - b2)
        denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
        p.data.addcdiv_(st["m"], denom, value=-lr / bc1)


model = bench.build_model("darcy")
params = [p for p in model.parameters()]
nfloat = sum(p.numel() * (2 if p.is_complex() else 1) for p in params)
for p in params:
    p.grad = torch.randn_like(p)
opt = Adam(params, lr=1e-3)
ms_fused = timed(lambda: opt.step())
states = [dict(m=torch.zeros_like(p), v=torch.zeros_like(p)) for p in params]
ms_ref = timed(lambda: reference_style_adam(params, states, 5))

B, S = 32, 421
x = torch.randn(B, S * S, device="cuda", requires_grad=True)
y = torch.randn(B, S * S, device="cuda")
loss = LpLoss(size_average=False)


def fused_loss():
    x.grad = None
    loss(x, y).backward()


def torch_loss():
    x.grad = None
    d = torch.norm(x - y, 2, 1)
    torch.sum(d / torch.norm(y, 2, 1)).backward()


ms_lf, ms_lt = timed(fused_loss), timed(torch_loss)
print(json.dumps({
    "adam": {"tensors": len(params), "floats": nfloat, "fused_ms": ms_fused, "reference_style_torch_ms": ms_ref,
             "fused_GBps": 4.0 * nfloat * 6 / (ms_fused * 1e-3) / 1e9},
    "lp_loss_fwd_bwd": {"shape": [B, S * S], "fused_ms": ms_lf, "torch_ms": ms_lt},
}))
