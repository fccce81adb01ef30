"""Is the reference's own GPU path consistent with its CPU path?  The sweep (tools/sweep_spectral.py) found ~0.1 relative
difference between this library and the stock torch-CUDA layer at S = 128 and 256 only.  This script runs the SAME stock layer
(rfft2 / einsum / irfft2, integral_operators.py:181-207) on the CPU (MKL) and on the GPU (cuFFT) and compares both with the
CUDA kernels: the layer hands irfft2 a half-spectrum whose k2 = 0 column is not Hermitian along k1, for which a C2R
transform's result is implementation-defined."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from sweep_spectral import stock_torch_layer  # noqa: E402
from uno_b200 import integral_operators as IO  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


CASES = [(2, 4, S, 12) for S in (64, 96, 128, 256, 481, 512)] + [(256, 32, 128, 12), (64, 32, 256, 12), (16, 32, 128, 12), (64, 8, 128, 12)]
for B, C, S, m in CASES:
    if True:
        torch.manual_seed(0)
        layer = IO.SpectralConv2d_Uno(C, C, S, S, m, m).cuda()
        x = torch.randn(B, C, S, S, device="cuda")
        w1, w2 = layer.weights1.detach(), layer.weights2.detach()
        with torch.no_grad():
            ours = layer(x, S, S).cpu()
            gpu = stock_torch_layer(x, w1, w2, S, S, m, m).cpu()
            cpu = stock_torch_layer(x.cpu(), w1.cpu(), w2.cpu(), S, S, m, m)
            # the same spectrum with the k2 = 0 column made Hermitian along k1 first: every C2R implementation must agree on it
            xh = torch.fft.rfft2(x, norm="forward")
            yh = torch.zeros(B, C, S, S // 2 + 1, dtype=torch.cfloat, device="cuda")
            yh[:, :, :m, :m] = torch.einsum("bixy,ioxy->boxy", xh[:, :, :m, :m], w1)
            yh[:, :, -m:, :m] = torch.einsum("bixy,ioxy->boxy", xh[:, :, -m:, :m], w2)
            col = yh[:, :, :, 0]
            yh[:, :, :, 0] = 0.5 * (col + torch.conj(torch.roll(torch.flip(col, dims=[-1]), 1, dims=-1)))
            herm_gpu = torch.fft.irfft2(yh, s=(S, S), norm="forward").cpu()
            herm_cpu = torch.fft.irfft2(yh.cpu(), s=(S, S), norm="forward")
        print(f"B={B:4d} C={C:3d} S={S:4d} m={m}: stock GPU vs stock CPU {rel(gpu, cpu):.2e} | ours vs stock CPU {rel(ours, cpu):.2e} | ours vs stock GPU {rel(ours, gpu):.2e}"
              f" | Hermitian-symmetrised input: GPU vs CPU {rel(herm_gpu, herm_cpu):.2e}", flush=True)
