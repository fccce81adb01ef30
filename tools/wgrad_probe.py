"""Timing probes of the tcgen05 weight-gradient kernel (tc_wgrad.cuh): the backward of one 1x1 channel mix with parts of the
kernel switched off through the switch wgrad_debug (bit 1 no MMA, 2 no operand stores, 8 no global loads, 16 no proxy fence,
32 no final flush; results are garbage, only the time means something).  Any non-zero value selects the probe instantiation
wgrad_tc_kernel<true>; 0 is the shipped kernel.  Times come from the library's per-launch CUDA events (uno_profile_*); what a
part costs on the critical path is the difference to mode 0.  Written for the open question in DESIGN.md section 8: 1.8 us per
32-pixel chunk per CTA against 0.4 - 0.7 us in the analysis kernel.

    python tools/wgrad_probe.py            # prints one line per shape
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uno_b200 import config  # noqa: E402

from uno_b200 import _lib  # noqa: E402
from uno_b200 import integral_operators as IO  # noqa: E402

lib = _lib.get()
lib.uno_profile_report.restype = C.c_size_t


def wgrad_ms(layer, x, gy, d, n=6):
    def step():
        layer.zero_grad(set_to_none=True)
        layer(x, *d).backward(gy)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    lib.uno_profile_enable(1)
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    size = lib.uno_profile_report(None, 0)
    buf = C.create_string_buffer(size + 16)
    lib.uno_profile_report(buf, size + 16)
    lib.uno_profile_enable(0)
    prof = json.loads(buf.value.decode())
    w = prof["conv1x1_wgrad"]
    return w["ms"] / w["launches"], w["bytes"] / w["launches"]


MODES = (0, 64, 1, 2, 8, 16, 32, 11, 43)    # 64: the probe instantiation with nothing switched off
SHAPES = {"Darcy conv5 pointwise (128 -> 32 at 240^2)": (32, 128, 32, 240), "Darcy conv0 pointwise (32 -> 64 at 240^2)": (32, 32, 64, 240),
          "level 120^2 (64 -> 128)": (32, 64, 128, 120)}
for name, (B, Ci, Co, S) in SHAPES.items():
    torch.manual_seed(0)
    layer = IO.pointwise_op_2D(Ci, Co, S, S).cuda()      # identity-size resample: a pure channel mix
    x = torch.randn(B, Ci, S, S, device="cuda", requires_grad=True)
    gy = torch.randn(B, Co, S, S, device="cuda")
    out = []
    nbytes = 0
    for mode in MODES:
        config.set("wgrad_debug", mode)
        ms, nbytes = wgrad_ms(layer, x, gy, (S, S))
        out.append(f"{mode}:{ms:.3f}")
    config.set("wgrad_debug", 0)
    chunks_per_cta = B * ((S * S + 31) // 32) / 148
    print(f"{name}: {nbytes / 1e6:.0f} MB per launch, {chunks_per_cta:.0f} chunks per CTA | weight-gradient kernel ms by debug mode |", "  ".join(out), flush=True)
