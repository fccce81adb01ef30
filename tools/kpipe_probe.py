"""Timing probes of the last-axis analysis kernel (tc_kpipe.cuh): the forward of one spectral convolution with parts of the
kernel switched off through the switch kpipe_debug (bit 1 no MMA, 2 no operand stores, 4 no B copy, 8 no global loads, 16 no
proxy fence; results are garbage, only the time means something).  The kernel's own
time comes from the library's per-launch CUDA events (uno_profile_*); what a part costs on the critical path is the
difference to mode 0."""
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


def analysis_ms(layer, x, d, n=8):
    with torch.no_grad():
        for _ in range(2):
            layer(x, *d)
        torch.cuda.synchronize()
        lib.uno_profile_enable(1)
        for _ in range(n):
            layer(x, *d)
        torch.cuda.synchronize()
        size = lib.uno_profile_report(None, 0)
        buf = C.create_string_buffer(size + 16)
        lib.uno_profile_report(buf, size + 16)
        lib.uno_profile_enable(0)
    prof = json.loads(buf.value.decode())
    return prof["dft_last_analysis"]["ms"] / n, sum(v["ms"] for v in prof.values()) / n


MODES = (0, 1, 2, 8, 15, 31, 47, 143, 191)    # 31 = 15 + no proxy fence, 47 = 15 + plain arrival for commit, 143 = 15 + per-warp arrival, 191 = all
for name, (B, Ci, Co, S, D, m) in {"481 (4-byte rows)": (32, 32, 64, 481, 240, 18), "240 (16-byte rows)": (32, 64, 128, 240, 120, 8),
                                    "512": (16, 32, 32, 512, 512, 20)}.items():
    torch.manual_seed(0)
    layer = IO.SpectralConv2d_Uno(Ci, Co, D, D, m, m).cuda()
    x = torch.randn(B, Ci, S, S, device="cuda")
    out = []
    for mode in MODES:
        config.set("kpipe_debug", mode)
        a, tot = analysis_ms(layer, x, (D, D))
        out.append(f"{mode}:{a:.3f}")
    config.set("kpipe_debug", 0)
    print(f"{name}: x {x.numel() * 4 / 1e6:.0f} MB, all kernels of the forward {tot:.3f} ms | analysis kernel ms by debug mode |", "  ".join(out), flush=True)
