"""
TEST INFRASTRUCTURE ONLY -- functional torch (CPU) port of the reference hot path.

Purpose
  * fp64-capable, autograd-differentiable restatement used by the GPU parity tests for backward
    passes (``tests/``), at sizes where the numpy oracle has no backward;
  * the timed ``cpu_baseline`` / ``--impl reference`` leg of ``bench.py`` (kind "port"): it issues the
    same library calls the reference issues on CPU (MKL rfftn/irfftn, complex einsum, 1x1 conv,
    anti-aliased bicubic interpolate), so its timing stands in for the reference, which is Python
    source under /root/reference and cannot travel to the GPU box.

Only ``tests/``, ``__graft_entry__.smoke()`` and those two ``bench.py`` legs may import it.  The product
(``uno_b200/``) never does.  Pinned against the real reference by ``tests/golden/make_golden.py``.

One dimension-generic implementation (the reference has three hand-unrolled copies); citations are
to /root/reference/integral_operators.py.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

_SUB = "xyz"


def _corners(modes: Sequence[int]):
    d = len(modes)
    last = slice(0, modes[-1])
    if d == 1:
        return [(last,)]
    lo = [slice(0, m) for m in modes[:-1]]
    hi = [slice(-m, None) for m in modes[:-1]]
    if d == 2:
        return [(lo[0], last), (hi[0], last)]
    return [(lo[0], lo[1], last), (hi[0], lo[1], last), (lo[0], hi[1], last), (hi[0], hi[1], last)]


def spectral_conv(x: torch.Tensor, weights: Sequence[torch.Tensor], out_dims, modes) -> torch.Tensor:
    """:47-72, :181-207, :385-427 -- rfftn(norm=forward), per-corner einsum into a zero spectrum of the
    OUTPUT grid, irfftn(s=out, norm=forward)."""
    d = len(out_dims)
    axes = list(range(-d, 0))
    cdtype = weights[0].dtype
    x_ft = torch.fft.rfftn(x, dim=axes, norm="forward")
    B, Co = x.shape[0], weights[0].shape[1]
    spec_shape = (B, Co) + tuple(int(v) for v in out_dims[:-1]) + (int(out_dims[-1]) // 2 + 1,)
    out_ft = torch.zeros(spec_shape, dtype=cdtype, device=x.device)
    s = _SUB[:d]
    expr = f"bi{s},io{s}->bo{s}"
    for sl, w in zip(_corners(modes), weights):
        idx = (slice(None), slice(None)) + sl
        out_ft[idx] = torch.einsum(expr, x_ft[idx], w)
    return torch.fft.irfftn(out_ft, s=[int(v) for v in out_dims], dim=axes, norm="forward")


def pointwise_op_2d(x, conv_w, conv_b, out_dims):
    """:224-243"""
    z = F.conv2d(x, conv_w, conv_b)
    return F.interpolate(z, size=(int(out_dims[0]), int(out_dims[1])), mode="bicubic", align_corners=True, antialias=True)


def pointwise_op_3d(x, conv_w, conv_b, out_dims):
    """:438-468"""
    d1, d2, d3 = (int(v) for v in out_dims)
    z = F.conv3d(x, conv_w, conv_b)
    ft = torch.fft.rfftn(z, dim=[-3, -2, -1])
    ft_u = torch.zeros_like(ft)
    h1, h2, h3 = d1 // 2, d2 // 2, d3 // 2
    for s1 in (slice(0, h1), slice(-h1, None)):
        for s2 in (slice(0, h2), slice(-h2, None)):
            ft_u[:, :, s1, s2, :h3] = ft[:, :, s1, s2, :h3]
    out = torch.fft.irfftn(ft_u, s=(d1, d2, d3))
    return F.interpolate(out, size=(d1, d2, d3), mode="trilinear", align_corners=True)


def operator_block(x, weights, conv_w, conv_b, out_dims, modes, gamma=None, beta=None, non_lin=True):
    """:272-284 / :501-513"""
    d = len(out_dims)
    a = spectral_conv(x, weights, out_dims, modes)
    b = pointwise_op_2d(x, conv_w, conv_b, out_dims) if d == 2 else pointwise_op_3d(x, conv_w, conv_b, out_dims)
    out = a + b
    if gamma is not None:
        out = F.instance_norm(out, weight=gamma, bias=beta, eps=1e-5)
    if non_lin:
        out = F.gelu(out)
    return out


# ---------------------------------------------------------------------------------------------
# nn.Module wrappers with the reference's parameter names, so reference state_dicts load 1:1 and
# ``uno_b200.models`` can be instantiated over this port (``ops=oracle.uno_torch_port``) in tests.
# ---------------------------------------------------------------------------------------------
class _SpectralConvND(nn.Module):
    def __init__(self, in_codim, out_codim, dims, modes, dtype=torch.cfloat):
        super().__init__()
        in_codim, out_codim = int(in_codim), int(out_codim)
        self.in_channels, self.out_channels = in_codim, out_codim
        self.dims = list(dims)
        self.modes = list(modes)
        scale = (1 / (2 * in_codim)) ** 0.5
        for i in range(2 ** (len(dims) - 1)):
            w = scale * torch.randn(in_codim, out_codim, *self.modes, dtype=dtype)
            setattr(self, f"weights{i + 1}", nn.Parameter(w))

    def forward(self, x, *dims):
        if dims and dims[0] is not None:
            self.dims = list(dims)
        ws = [getattr(self, f"weights{i + 1}") for i in range(2 ** (len(self.dims) - 1))]
        return spectral_conv(x, ws, self.dims, self.modes)


class SpectralConv2d_Uno(_SpectralConvND):
    def __init__(self, in_codim, out_codim, dim1, dim2, modes1=None, modes2=None):
        if modes1 is None:
            modes1, modes2 = dim1 // 2 - 1, dim2 // 2
        super().__init__(in_codim, out_codim, (dim1, dim2), (modes1, modes2))


class SpectralConv3d_Uno(_SpectralConvND):
    def __init__(self, in_codim, out_codim, dim1, dim2, dim3, modes1=None, modes2=None, modes3=None):
        if modes1 is None:
            modes1, modes2, modes3 = dim1, dim2, dim3 // 2 + 1
        super().__init__(in_codim, out_codim, (dim1, dim2, dim3), (modes1, modes2, modes3))


class _PointwiseND(nn.Module):
    def __init__(self, in_codim, out_codim, dims):
        super().__init__()
        conv = {2: nn.Conv2d, 3: nn.Conv3d}[len(dims)]
        self.conv = conv(int(in_codim), int(out_codim), 1)
        self.dims = [int(v) for v in dims]

    def forward(self, x, *dims):
        od = list(dims) if dims and dims[0] is not None else self.dims
        fn = pointwise_op_2d if len(od) == 2 else pointwise_op_3d
        return fn(x, self.conv.weight, self.conv.bias, od)


class pointwise_op_2D(_PointwiseND):
    def __init__(self, in_codim, out_codim, dim1, dim2):
        super().__init__(in_codim, out_codim, (dim1, dim2))


class pointwise_op_3D(_PointwiseND):
    def __init__(self, in_codim, out_codim, dim1, dim2, dim3):
        super().__init__(in_codim, out_codim, (dim1, dim2, dim3))


class _OperatorBlockND(nn.Module):
    def _finish(self, x1, x2):
        out = x1 + x2
        if self.normalize:
            out = self.normalize_layer(out)
        if self.non_lin:
            out = F.gelu(out)
        return out

    def forward(self, x, *dims):
        return self._finish(self.conv(x, *dims), self.w(x, *dims))


class OperatorBlock_2D(_OperatorBlockND):
    def __init__(self, in_codim, out_codim, dim1, dim2, modes1, modes2, Normalize=False, Non_Lin=True):
        super().__init__()
        self.conv = SpectralConv2d_Uno(in_codim, out_codim, dim1, dim2, modes1, modes2)
        self.w = pointwise_op_2D(in_codim, out_codim, dim1, dim2)
        self.normalize, self.non_lin = Normalize, Non_Lin
        if Normalize:
            self.normalize_layer = nn.InstanceNorm2d(int(out_codim), affine=True)


class OperatorBlock_3D(_OperatorBlockND):
    def __init__(self, in_codim, out_codim, dim1, dim2, dim3, modes1, modes2, modes3, Normalize=False, Non_Lin=True):
        super().__init__()
        self.conv = SpectralConv3d_Uno(in_codim, out_codim, dim1, dim2, dim3, modes1, modes2, modes3)
        self.w = pointwise_op_3D(in_codim, out_codim, dim1, dim2, dim3)
        self.normalize, self.non_lin = Normalize, Non_Lin
        if Normalize:
            self.normalize_layer = nn.InstanceNorm3d(int(out_codim), affine=True)


class SpectralConv1d_Uno(_SpectralConvND):
    def __init__(self, in_codim, out_codim, dim1, modes1=None):
        if modes1 is None:
            modes1 = dim1 // 2
        super().__init__(in_codim, out_codim, (dim1,), (modes1,))


# ---- model glue (callers of the hot path) ----------------------------------
def lift(a, grid, w_a, b_a, w_b, b_b, pad_lo, pad_hi):
    """darcy_flow_uno2d.py:96-107 / navier_stokes_uno2d.py:191-201 / navier_stokes_uno3d.py:497-511:
    cat(a, grid) -> Linear -> gelu -> Linear -> gelu -> channels-first -> F.pad (zeros)."""
    x = torch.cat((a, grid.expand(a.shape[0], *grid.shape)), dim=-1)
    h = F.gelu(F.linear(F.gelu(F.linear(x, w_a, b_a)), w_b, b_b))
    nd = h.dim() - 2
    h = h.permute(0, nd + 1, *range(1, nd + 1))
    pads = []
    for lo, hi in reversed(list(zip(pad_lo, pad_hi))):
        pads += [int(lo), int(hi)]
    return F.pad(h, pads) if any(pads) else h.contiguous()


def project(srcs, w1, b1, w2, b2, crop_lo, crop_hi):
    """darcy_flow_uno2d.py:121-131 / navier_stokes_uno2d.py:215-225 / navier_stokes_uno3d.py:551-575:
    cat(srcs, dim=1) -> crop -> channels-last -> Linear -> gelu -> Linear."""
    c = torch.cat(list(srcs), dim=1) if len(srcs) > 1 else srcs[0]
    idx = [slice(None), slice(None)]
    for a, (lo, hi) in enumerate(zip(crop_lo, crop_hi)):
        n = c.shape[2 + a]
        idx.append(slice(int(lo), n - int(hi)))
    c = c[tuple(idx)]
    nd = c.dim() - 2
    c = c.permute(0, *range(2, nd + 2), 1)
    return F.linear(F.gelu(F.linear(c, w1, b1)), w2, b2)


# ---- training-step ops --------------------------------------------------------------
class LpLoss:
    """utilities3.py:75-103: relative Lp error per sample, summed / averaged over the batch."""

    def __init__(self, d=2, p=2, size_average=True, reduction=True):
        assert d > 0 and p > 0
        self.d, self.p, self.reduction, self.size_average = d, p, reduction, size_average

    def rel(self, x, y):
        b = x.size()[0]
        ratio = torch.norm(x.reshape(b, -1) - y.reshape(b, -1), self.p, 1) / torch.norm(y.reshape(b, -1), self.p, 1)
        if not self.reduction:
            return ratio
        return ratio.mean() if self.size_average else ratio.sum()

    __call__ = rel


def adam_step(param, grad, state, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
    """One update of Adam.py:23-52 on a single tensor, out of place, any dtype (complex: second moment = |g|^2 kept as
    a complex number with zero imaginary part).  state = dict(exp_avg, exp_avg_sq[, max_exp_avg_sq]); returns the new
    parameter and updates `state` in place."""
    b1, b2 = betas
    g = grad + weight_decay * param if weight_decay != 0 else grad
    state["exp_avg"] = b1 * state["exp_avg"] + (1 - b1) * g
    state["exp_avg_sq"] = b2 * state["exp_avg_sq"] + (1 - b2) * (g * g.conj())
    second = state["exp_avg_sq"]
    if amsgrad:
        state["max_exp_avg_sq"] = torch.maximum(state["max_exp_avg_sq"], second)
        second = state["max_exp_avg_sq"]
    denom = second.sqrt() / (1 - b2**step) ** 0.5 + eps
    return param - (lr / (1 - b1**step)) * state["exp_avg"] / denom
