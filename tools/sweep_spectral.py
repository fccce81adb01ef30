"""SURVEY.md §8(d) config 5 and the per-level roofline: the spectral convolution alone, forward and backward timed
separately with CUDA events, against the algorithmic bytes of §8(d) and against the stock torch-CUDA path
(cuFFT full-spectrum transforms + a complex einsum — what the reference's module does on a GPU) on the same B200.

    python tools/sweep_spectral.py --mode sweep  [--out gpurun_out/sweep.json]   # S x m x C grid, 512 MiB inputs
    python tools/sweep_spectral.py --mode levels [--workload darcy]              # every block of the model, its own shapes

Inputs are larger than L2 in sweep mode (512 MiB); in levels mode a 256 MiB buffer is rewritten between iterations.
The baseline is a plain torch restatement kept inside this file (rfft2 / einsum / irfft2 with the semantics of
integral_operators.py:181-207): it is a timing bar only — parity is proven in tests/ against oracle/.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from uno_b200 import integral_operators as IO  # noqa: E402


def stock_torch_layer(x, w1, w2, d1, d2, m1, m2):
    xh = torch.fft.rfft2(x, norm="forward")
    yh = torch.zeros(x.shape[0], w1.shape[1], d1, d2 // 2 + 1, dtype=torch.cfloat, device=x.device)
    yh[:, :, :m1, :m2] = torch.einsum("bixy,ioxy->boxy", xh[:, :, :m1, :m2], w1)
    yh[:, :, -m1:, :m2] = torch.einsum("bixy,ioxy->boxy", xh[:, :, -m1:, :m2], w2)
    return torch.fft.irfft2(yh, s=(d1, d2), norm="forward")


def conv_bytes(B, Ci, Co, n_in, n_out, M, corners=2):
    """§8(d): fwd = 4B(Ci·Nin + Co·Nout) + 8·nW·Ci·Co·M ; bwd = 4B(Co·Nout + Ci·Nin) + 8·B·Ci·nW·M + 16·nW·Ci·Co·M."""
    fwd = 4 * B * (Ci * n_in + Co * n_out) + 8 * corners * Ci * Co * M
    bwd = 4 * B * (Co * n_out + Ci * n_in) + 8 * B * Ci * corners * M + 16 * corners * Ci * Co * M
    return fwd, bwd


class Timer:
    def __init__(self, flush):
        self.scratch = torch.empty(64 << 20, device="cuda") if flush else None

    def ms(self, prep, fn, iters):
        """median over `iters` of fn() alone; prep() runs untimed before each (rebuilds the autograd graph)."""
        out = []
        for _ in range(iters):
            state = prep()
            if self.scratch is not None:
                self.scratch.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(state)
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1))
        out.sort()
        return out[len(out) // 2]


def time_layer(timer, fwd, params, x, g, iters):
    """fwd(x) -> y ; returns (fwd ms, bwd ms). Forward runs with autograd recording (it saves what backward needs)."""
    for _ in range(2):
        y = fwd(x)
        torch.autograd.grad(y, [x] + params, g)
    t_f = timer.ms(lambda: None, lambda _: fwd(x), iters)
    t_b = timer.ms(lambda: fwd(x), lambda y: torch.autograd.grad(y, [x] + params, g), iters)
    return t_f, t_b


def measure(timer, B, Ci, Co, hin, win, d1, d2, m1, m2, iters, peak):
    torch.manual_seed(0)
    layer = IO.SpectralConv2d_Uno(Ci, Co, d1, d2, m1, m2).cuda()
    torch.manual_seed(1)
    x = torch.randn(B, Ci, hin, win, device="cuda", requires_grad=True)
    g = torch.randn(B, Co, d1, d2, device="cuda")
    params = [layer.weights1, layer.weights2]
    ours = time_layer(timer, lambda t: layer(t, d1, d2), params, x, g, iters)
    stock = time_layer(timer, lambda t: stock_torch_layer(t, layer.weights1, layer.weights2, d1, d2, m1, m2), params, x, g, iters)
    with torch.no_grad():
        ya, yb = layer(x, d1, d2), stock_torch_layer(x, layer.weights1, layer.weights2, d1, d2, m1, m2)
        err = float((ya - yb).abs().max() / yb.abs().max())
    bf, bb = conv_bytes(B, Ci, Co, hin * win, d1 * d2, m1 * m2)
    row = {"B": B, "Ci": Ci, "Co": Co, "in": [hin, win], "out": [d1, d2], "modes": [m1, m2],
           "fwd_ms": round(ours[0], 4), "bwd_ms": round(ours[1], 4),
           "fwd_gbs": round(bf / ours[0] / 1e6, 1), "bwd_gbs": round(bb / ours[1] / 1e6, 1),
           "fwd_frac": round(bf / ours[0] / 1e6 / peak, 3), "bwd_frac": round(bb / ours[1] / 1e6 / peak, 3),
           "stock_fwd_ms": round(stock[0], 4), "stock_bwd_ms": round(stock[1], 4),
           "speedup_fwd": round(stock[0] / ours[0], 2), "speedup_bwd": round(stock[1] / ours[1], 2),
           "max_rel_diff_vs_stock": err}
    del layer, x, g
    torch.cuda.empty_cache()
    return row


def model_levels(workload, B):
    """(Ci, Co, in grid, out grid, modes) of every 2-D operator block, recorded from one forward of the real model."""
    model = bench.build_model(workload)
    seen = []

    def hook(mod, args, out):
        seen.append((mod.conv.in_channels, mod.conv.out_channels, tuple(args[0].shape[2:]), tuple(out.shape[2:]),
                     (mod.conv.modes1, mod.conv.modes2)))

    hs = [m.register_forward_hook(hook) for m in model.modules() if isinstance(m, IO.OperatorBlock_2D)]
    xshape = bench.WORKLOADS[workload][3]
    with torch.no_grad():
        model(torch.randn(2, *xshape, device="cuda"))
    for h in hs:
        h.remove()
    del model
    torch.cuda.empty_cache()
    return seen


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="sweep", choices=["sweep", "levels"])
    ap.add_argument("--workload", default="darcy")
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--sizes", default="64,128,256,512")
    ap.add_argument("--modes", default="12,20,32")
    ap.add_argument("--channels", default="32,64,128")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    peak, how = bench.peaks()
    rows = []
    if args.mode == "sweep":
        timer = Timer(flush=False)
        for S in map(int, args.sizes.split(",")):
            for m in map(int, args.modes.split(",")):
                for C in map(int, args.channels.split(",")):
                    B = (1 << 27) // (C * S * S)
                    rows.append(measure(timer, B, C, C, S, S, S, S, m, m, args.iters, peak))
                    print(json.dumps(rows[-1]), flush=True)
    else:
        timer = Timer(flush=True)
        B = bench.WORKLOADS[args.workload][5]
        for Ci, Co, gin, gout, modes in model_levels(args.workload, B):
            rows.append(measure(timer, B, Ci, Co, gin[0], gin[1], gout[0], gout[1], modes[0], modes[1], args.iters, peak))
            print(json.dumps(rows[-1]), flush=True)
    doc = {"mode": args.mode, "workload": args.workload if args.mode == "levels" else None, "hbm_peak_gbs": peak, "peak_source": how,
           "bytes": "SURVEY.md 8(d) algorithmic bytes per SpectralConv call", "timing": f"median of {args.iters}, CUDA events, "
           + ("512 MiB inputs (> L2)" if args.mode == "sweep" else "256 MiB scratch rewritten between iterations"),
           "stock": "torch.fft.rfft2 / einsum / irfft2 on the same GPU (cuFFT + cuBLAS)", "rows": rows}
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(doc, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
