"""Relative Lp loss of the reference training loops (utilities3.py:75-103, ``LpLoss.rel``), same constructor and
call signature.  For p = 2 (the only value the reference uses) forward and backward are CUDA kernels behind the C ABI
(``uno_lp_loss_fwd`` / ``uno_lp_loss_bwd``): one pass over x and y each way instead of six elementwise / reduction
launches.  Inputs must be CUDA float32 tensors; other p fall back to the same torch expression upstream evaluates."""
import ctypes as C

import torch

from . import _capi
from ._lib import get as _get_lib


class _RelL2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, reduction):
        lib = _get_lib()
        for t in (x, y):
            if not t.is_cuda or t.dtype != torch.float32:
                raise RuntimeError(f"uno_b200: LpLoss expects CUDA float32 tensors (got {t.dtype} on {t.device}); there is no CPU path")
        B = x.shape[0]
        x2, y2 = x.reshape(B, -1).contiguous(), y.reshape(B, -1).contiguous()
        if x2.shape != y2.shape:
            raise RuntimeError(f"uno_b200: LpLoss shapes differ: {tuple(x.shape)} vs {tuple(y.shape)}")
        N = x2.shape[1]
        loss = torch.empty(B if reduction == 0 else 1, dtype=torch.float32, device=x.device)
        norms = torch.empty((B, 2), dtype=torch.float32, device=x.device)
        ws = torch.empty(2 * B, dtype=torch.float64, device=x.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        _capi.check(lib, lib.uno_lp_loss_fwd(p(x2), p(y2), B, N, reduction, p(loss), p(norms), p(ws), ws.numel() * 8, st))
        ctx.save_for_backward(x2, y2, norms)
        ctx.reduction, ctx.x_shape = reduction, x.shape
        return loss if reduction == 0 else loss[0]

    @staticmethod
    def backward(ctx, gl):
        lib = _get_lib()
        x2, y2, norms = ctx.saved_tensors
        B, N = x2.shape
        gl = gl.reshape(-1).contiguous().float()
        gx = torch.empty_like(x2)
        st = C.c_void_p(torch.cuda.current_stream(x2.device).cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        _capi.check(lib, lib.uno_lp_loss_bwd(p(x2), p(y2), p(norms), p(gl), B, N, ctx.reduction, p(gx), st))
        return gx.reshape(ctx.x_shape), None, None


class LpLoss:
    def __init__(self, d=2, p=2, size_average=True, reduction=True):
        assert d > 0 and p > 0
        self.d, self.p, self.reduction, self.size_average = d, p, reduction, size_average

    def rel(self, x, y):
        n = x.size()[0]
        if self.p == 2 and not y.requires_grad:
            mode = 0 if not self.reduction else (2 if self.size_average else 1)
            return _RelL2Fn.apply(x, y, mode)
        diff = torch.norm(x.reshape(n, -1) - y.reshape(n, -1), self.p, 1)
        ynorm = torch.norm(y.reshape(n, -1), self.p, 1)
        if self.reduction:
            return torch.mean(diff / ynorm) if self.size_average else torch.sum(diff / ynorm)
        return diff / ynorm

    def __call__(self, x, y):
        return self.rel(x, y)
