"""Drop-in replacement for the reference's ``integral_operators`` module.

Same class names, constructor / forward signatures, public attributes, parameter names, shapes,
dtypes and init RNG order as /root/reference/integral_operators.py, so that the reference's model
files (``from integral_operators import *``) and checkpoints (state_dict keys)
work unchanged -- but every forward/backward runs the hand-written sm_100a kernels behind the C ABI
(``include/uno_b200.h``) instead of torch.fft / einsum / cuDNN / ATen interpolate.

Inputs must be CUDA float32 tensors.  There is no CPU fallback: a CPU tensor raises RuntimeError.
"""
from __future__ import annotations

import numpy as np  # noqa: F401  (re-exported: the reference module exposes np/torch/nn/F via `import *`)
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401

from . import functional as _fn

__all__ = [
    "SpectralConv1d_Uno", "SpectralConv2d_Uno", "SpectralConv3d_Uno",
    "pointwise_op_1D", "pointwise_op_2D", "pointwise_op_3D",
    "OperatorBlock_1D", "OperatorBlock_2D", "OperatorBlock_3D",
    "torch", "np", "nn", "F",
]


def _spectral_param(scale, in_codim, out_codim, *modes):
    # same RNG consumption as the reference: scale * torch.randn(..., dtype=cfloat)
    return nn.Parameter(scale * torch.randn(in_codim, out_codim, *modes, dtype=torch.cfloat))


class SpectralConv1d_Uno(nn.Module):
    """1-D Fourier integral operator with grid resampling (integral_operators.py:7-72)."""

    def __init__(self, in_codim, out_codim, dim1, modes1=None):
        super().__init__()
        in_codim, out_codim = int(in_codim), int(out_codim)
        self.in_channels, self.out_channels = in_codim, out_codim
        self.dim1 = dim1
        self.modes1 = modes1 if modes1 is not None else dim1 // 2
        self.scale = (1 / (2 * in_codim)) ** (1.0 / 2.0)
        self.weights1 = _spectral_param(self.scale, in_codim, out_codim, self.modes1)

    def compl_mul1d(self, input, weights):
        return torch.einsum("bix,iox->box", input, weights)

    def forward(self, x, dim1=None):
        if dim1 is not None:
            self.dim1 = dim1  # sticky, as in the reference (:52-53)
        return _fn.spectral_conv(x, [self.weights1], (self.dim1,), (self.modes1,))


class SpectralConv2d_Uno(nn.Module):
    """2-D Fourier integral operator with grid resampling (integral_operators.py:127-207)."""

    def __init__(self, in_codim, out_codim, dim1, dim2, modes1=None, modes2=None):
        super().__init__()
        in_codim, out_codim = int(in_codim), int(out_codim)
        self.in_channels, self.out_channels = in_codim, out_codim
        self.dim1, self.dim2 = dim1, dim2
        if modes1 is not None:
            self.modes1, self.modes2 = modes1, modes2
        else:
            self.modes1, self.modes2 = dim1 // 2 - 1, dim2 // 2
        self.scale = (1 / (2 * in_codim)) ** (1.0 / 2.0)
        self.weights1 = _spectral_param(self.scale, in_codim, out_codim, self.modes1, self.modes2)
        self.weights2 = _spectral_param(self.scale, in_codim, out_codim, self.modes1, self.modes2)

    def compl_mul2d(self, input, weights):
        return torch.einsum("bixy,ioxy->boxy", input, weights)

    def forward(self, x, dim1=None, dim2=None):
        if dim1 is not None:
            self.dim1, self.dim2 = dim1, dim2  # sticky (:182-184)
        return _fn.spectral_conv(x, [self.weights1, self.weights2], (self.dim1, self.dim2), (self.modes1, self.modes2))


class SpectralConv3d_Uno(nn.Module):
    """3-D Fourier integral operator, four corner blocks (integral_operators.py:287-427)."""

    def __init__(self, in_codim, out_codim, dim1, dim2, dim3, modes1=None, modes2=None, modes3=None):
        super().__init__()
        in_codim, out_codim = int(in_codim), int(out_codim)
        self.in_channels, self.out_channels = in_codim, out_codim
        self.dim1, self.dim2, self.dim3 = dim1, dim2, dim3
        if modes1 is not None:
            self.modes1, self.modes2, self.modes3 = modes1, modes2, modes3
        else:
            self.modes1, self.modes2, self.modes3 = dim1, dim2, dim3 // 2 + 1
        self.scale = (1 / (2 * in_codim)) ** (1.0 / 2.0)
        m = (self.modes1, self.modes2, self.modes3)
        self.weights1 = _spectral_param(self.scale, in_codim, out_codim, *m)
        self.weights2 = _spectral_param(self.scale, in_codim, out_codim, *m)
        self.weights3 = _spectral_param(self.scale, in_codim, out_codim, *m)
        self.weights4 = _spectral_param(self.scale, in_codim, out_codim, *m)

    def compl_mul3d(self, input, weights):
        return torch.einsum("bixyz,ioxyz->boxyz", input, weights)

    def forward(self, x, dim1=None, dim2=None, dim3=None):
        if dim1 is not None:
            self.dim1, self.dim2, self.dim3 = dim1, dim2, dim3  # sticky (:391-394)
        ws = [self.weights1, self.weights2, self.weights3, self.weights4]
        return _fn.spectral_conv(x, ws, (self.dim1, self.dim2, self.dim3), (self.modes1, self.modes2, self.modes3))


class pointwise_op_1D(nn.Module):
    """integral_operators.py:75-93.  The reference forward asks F.interpolate for mode='linear' with
    antialias=True, which torch >= 1.11 rejects with ValueError; that behaviour is preserved."""

    def __init__(self, in_codim, out_codim, dim1):
        super().__init__()
        self.conv = nn.Conv1d(int(in_codim), int(out_codim), 1)
        self.dim1 = int(dim1)

    def forward(self, x, dim1=None):
        raise ValueError(
            "Anti-alias option is restricted to bilinear and bicubic modes and requires a 4-D tensor as input "
            "(pointwise_op_1D: the reference raises the same error on torch >= 1.11)"
        )


class pointwise_op_2D(nn.Module):
    """Conv2d(k=1) + bicubic anti-aliased resample, align_corners=True (integral_operators.py:210-243)."""

    def __init__(self, in_codim, out_codim, dim1, dim2):
        super().__init__()
        self.conv = nn.Conv2d(int(in_codim), int(out_codim), 1)
        self.dim1, self.dim2 = int(dim1), int(dim2)

    def forward(self, x, dim1=None, dim2=None):
        if dim1 is None:
            dim1, dim2 = self.dim1, self.dim2
        return _fn.pointwise_op(x, self.conv.weight, self.conv.bias, (dim1, dim2))


class pointwise_op_3D(nn.Module):
    """Conv3d(k=1) + rfftn / corner copy / irfftn(s=out) resample (integral_operators.py:430-468)."""

    def __init__(self, in_codim, out_codim, dim1, dim2, dim3):
        super().__init__()
        self.conv = nn.Conv3d(int(in_codim), int(out_codim), 1)
        self.dim1, self.dim2, self.dim3 = int(dim1), int(dim2), int(dim3)

    def forward(self, x, dim1=None, dim2=None, dim3=None):
        if dim1 is None:
            dim1, dim2, dim3 = self.dim1, self.dim2, self.dim3
        return _fn.pointwise_op(x, self.conv.weight, self.conv.bias, (dim1, dim2, dim3))


class OperatorBlock_1D(nn.Module):
    """integral_operators.py:96-124 (forward raises through pointwise_op_1D, as the reference does)."""

    def __init__(self, in_codim, out_codim, dim1, modes1, Normalize=True, Non_Lin=True):
        super().__init__()
        self.conv = SpectralConv1d_Uno(in_codim, out_codim, dim1, modes1)
        self.w = pointwise_op_1D(in_codim, out_codim, dim1)
        self.normalize = Normalize
        self.non_lin = Non_Lin
        if Normalize:
            self.normalize_layer = torch.nn.InstanceNorm1d(int(out_codim), affine=True)

    def forward(self, x, dim1=None):
        x1_out = self.conv(x, dim1)
        x2_out = self.w(x, dim1)  # raises ValueError, see pointwise_op_1D
        return x1_out + x2_out


class _OperatorBlockND(nn.Module):
    def _run(self, x, dims, fanout=1):
        conv = self.conv
        nd = len(dims)
        names = ("dim1", "dim2", "dim3")[:nd]
        if dims[0] is not None:
            for n, v in zip(names, dims):
                setattr(conv, n, v)  # the reference's inner conv call mutates conv.dim* (sticky)
            out_dims = tuple(dims)
        else:
            out_dims = tuple(getattr(conv, n) for n in names)
            # the pointwise branch falls back to ITS OWN stored dims in the reference; a mismatch
            # makes the reference fail at the add, so we require agreement
            w_dims = tuple(getattr(self.w, n) for n in names)
            if tuple(int(v) for v in out_dims) != w_dims:
                raise RuntimeError(
                    f"The size of tensor a {tuple(out_dims)} must match the size of tensor b {w_dims} "
                    "(conv and w disagree on the default output grid)"
                )
        modes = tuple(getattr(conv, n) for n in ("modes1", "modes2", "modes3")[:nd])
        weights = [getattr(conv, f"weights{i + 1}") for i in range(2 ** (nd - 1))]
        gamma = beta = None
        eps = 1e-5
        if self.normalize:
            gamma, beta, eps = self.normalize_layer.weight, self.normalize_layer.bias, self.normalize_layer.eps
        return _fn.operator_block(x, weights, self.w.conv.weight, self.w.conv.bias, out_dims, modes, gamma, beta, self.non_lin, eps,
                                  fanout=fanout)


class OperatorBlock_2D(_OperatorBlockND):
    """gelu?(InstanceNorm2d?(conv(x) + w(x))) in one fused call (integral_operators.py:246-284)."""

    def __init__(self, in_codim, out_codim, dim1, dim2, modes1, modes2, Normalize=False, Non_Lin=True):
        super().__init__()
        self.conv = SpectralConv2d_Uno(in_codim, out_codim, dim1, dim2, modes1, modes2)
        self.w = pointwise_op_2D(in_codim, out_codim, dim1, dim2)
        self.normalize = Normalize
        self.non_lin = Non_Lin
        if Normalize:
            self.normalize_layer = torch.nn.InstanceNorm2d(int(out_codim), affine=True)

    def forward(self, x, dim1=None, dim2=None, fanout=1):
        """``fanout=2`` (extension, not in the reference's signature): return the output twice, as two aliases of the same
        memory, for a tensor that feeds two consumers (a skip connection).  The block's backward then receives the two
        upstream gradients separately and adds them inside its first kernel instead of autograd's separate add pass."""
        return self._run(x, (dim1, dim2), fanout)


class OperatorBlock_3D(_OperatorBlockND):
    """gelu?(InstanceNorm3d?(conv(x) + w(x))) in one fused call (integral_operators.py:471-513)."""

    def __init__(self, in_codim, out_codim, dim1, dim2, dim3, modes1, modes2, modes3, Normalize=False, Non_Lin=True):
        super().__init__()
        self.conv = SpectralConv3d_Uno(in_codim, out_codim, dim1, dim2, dim3, modes1, modes2, modes3)
        self.w = pointwise_op_3D(in_codim, out_codim, dim1, dim2, dim3)
        self.normalize = Normalize
        self.non_lin = Non_Lin
        if Normalize:
            self.normalize_layer = torch.nn.InstanceNorm3d(int(out_codim), affine=True)

    def forward(self, x, dim1=None, dim2=None, dim3=None, fanout=1):
        return self._run(x, (dim1, dim2, dim3), fanout)
