"""U-shaped neural operator architectures over the B200 operator blocks.

Host-side mirror of the reference model files (the callers of the hot path):
same class names, constructor arguments, sub-module names (hence state_dict keys, Appendix D),
parameter-creation order (hence seed-for-seed identical initial weights) and forward semantics as

    UNO_9                 /root/reference/darcy_flow_uno2d.py:27-141
    UNO, UNO_P            /root/reference/navier_stokes_uno2d.py:145-238, :24-138
    Uno3D_T10             /root/reference/navier_stokes_uno3d.py:412-603

written table-driven instead of unrolled.  The reference's own model files also run unchanged on top
of ``uno_b200.integral_operators``; these classes exist so that the benchmark
and the tests do not need /root/reference at run time.  ``ops`` selects the module that provides
``OperatorBlock_2D/3D`` (default: the CUDA drop-in; the tests pass the CPU oracle port).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def _default_ops():
    from . import integral_operators

    return integral_operators


def _glue_of(ops):
    """The module providing the fused model glue ``lift`` / ``project`` for an operator module: the CUDA
    kernels behind the C ABI (``uno_b200.functional``) for the drop-in, or the operator module itself when
    it brings its own (the CPU oracle port used by the tests)."""
    if hasattr(ops, "lift") and hasattr(ops, "project"):
        return ops
    from . import functional

    return functional


def _twice(fused, fn, *args):
    """A tensor with two consumers (a skip connection).  On the CUDA blocks the producer returns two aliases of its output
    (``fanout=2``) and its backward adds the two upstream gradients inside its first kernel; any other operator module (the CPU
    oracle port of the tests) just hands the tensor out twice and autograd adds the gradients."""
    if fused:
        return fn(*args, fanout=2)
    y = fn(*args)
    return y, y


class _GridCache:
    """The reference rebuilds its coordinate features with numpy on the host in every forward
    (darcy_flow_uno2d.py:135-141).  Same values, built once per (shape, device)."""

    def __init__(self):
        self._cache = {}

    def get(self, key, device, builder):
        k = (key, str(device))
        g = self._cache.get(k)
        if g is None:
            g = builder().to(device)
            self._cache[k] = g
        return g


def _lin(a, b, n):
    return torch.tensor(np.linspace(a, b, n), dtype=torch.float)


class UNO_9(nn.Module):
    """Darcy-flow U-NO, 5 operator blocks (darcy_flow_uno2d.py:27-141)."""

    def __init__(self, in_width, width, pad=5, factor=1, ops=None):
        super().__init__()
        ops = ops or _default_ops()
        Blk = ops.OperatorBlock_2D
        self.in_width, self.width, self.padding = in_width, width, pad
        w, f = width, factor
        self.fc_n1 = nn.Linear(in_width, w // 2)
        self.fc0 = nn.Linear(w // 2, w)
        self.conv0 = Blk(w, 2 * f * w, 40, 40, 18, 18)
        self.conv1 = Blk(2 * f * w, 4 * f * w, 20, 20, 8, 8, Normalize=True)
        self.conv2 = Blk(4 * f * w, 4 * f * w, 20, 20, 8, 8)
        self.conv4 = Blk(4 * f * w, 2 * f * w, 40, 40, 8, 8, Normalize=True)
        self.conv5 = Blk(4 * f * w, w, 85, 85, 18, 18)
        self.fc1 = nn.Linear(2 * w, w)
        self.fc2 = nn.Linear(w, 1)
        self._grids = _GridCache()
        self._glue = _glue_of(ops)
        self._fused = self._glue is not ops

    def get_grid(self, shape, device):
        b, sx, sy = shape[0], shape[1], shape[2]

        def build():
            gx = _lin(0, 1, sx).reshape(1, sx, 1, 1).repeat([1, 1, sy, 1])
            gy = _lin(0, 1, sy).reshape(1, 1, sy, 1).repeat([1, sx, 1, 1])
            return torch.cat((gx, gy), dim=-1)

        return self._grids.get((sx, sy), device, build).expand(b, -1, -1, -1)

    def forward(self, x):
        # lift + permute + pad (darcy_flow_uno2d.py:96-107) and cat + crop + permute + projection (:121-131) are one
        # kernel each; the padding grows the grid to the right / bottom only
        grid = self.get_grid((1,) + tuple(x.shape[1:]), x.device)[0]
        grow = math.ceil(x.shape[2] / 85) * self.padding
        # h and c0 each feed two consumers: their producers fan them out (see _twice)
        h, h_skip = _twice(self._fused, self._glue.lift, x, grid, self.fc_n1.weight, self.fc_n1.bias, self.fc0.weight, self.fc0.bias,
                           (0, 0), (grow, grow))
        D1, D2 = h.shape[-2], h.shape[-1]
        c0, c0_skip = _twice(self._fused, self.conv0, h, D1 // 2, D2 // 2)
        c1 = self.conv1(c0, D1 // 4, D2 // 4)
        c2 = self.conv2(c1, D1 // 4, D2 // 4)
        c4 = torch.cat([self.conv4(c2, D1 // 2, D2 // 2), c0_skip], dim=1)
        c5 = self.conv5(c4, D1, D2)
        return self._glue.project([c5, h_skip], self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, (0, 0), (grow, grow))


class _NS2DBase(nn.Module):
    def get_grid(self, shape, device):
        b, sx, sy = shape[0], shape[1], shape[2]

        def build():
            gx = _lin(0, 2 * np.pi, sx).reshape(1, sx, 1, 1).repeat([1, 1, sy, 1])
            gy = _lin(0, 2 * np.pi, sy).reshape(1, 1, sy, 1).repeat([1, sx, 1, 1])
            return torch.cat((torch.sin(gx), torch.sin(gy), torch.cos(gx), torch.cos(gy)), dim=-1)

        return self._grids.get((sx, sy), device, build).expand(b, -1, -1, -1)

    def _lift(self, x):
        # navier_stokes_uno2d.py:191-201 -- F.pad on all four sides
        grid = self.get_grid((1,) + tuple(x.shape[1:]), x.device)[0]
        p = self.padding
        return _twice(self._fused, self._glue.lift, x, grid, self.fc.weight, self.fc.bias, self.fc0.weight, self.fc0.bias, (p, p), (p, p))

    def _project(self, srcs):
        # navier_stokes_uno2d.py:215-225 -- the reference crops only the trailing edge ([..., :-p, :-p])
        p = self.padding
        return self._glue.project(srcs, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, (0, 0), (p, p))


class UNO(_NS2DBase):
    """Navier-Stokes 2-D U-NO, 7 blocks, 3/4 domain scaling (navier_stokes_uno2d.py:145-238)."""

    def __init__(self, in_width, width, pad=0, factor=3 / 4, ops=None):
        super().__init__()
        ops = ops or _default_ops()
        Blk = ops.OperatorBlock_2D
        self.in_width, self.width, self.factor, self.padding = in_width, width, factor, pad
        w, f = width, factor
        self.fc = nn.Linear(in_width, w // 2)
        self.fc0 = nn.Linear(w // 2, w)
        self.L0 = Blk(w, 2 * f * w, 48, 48, 22, 22)
        self.L1 = Blk(2 * f * w, 4 * f * w, 32, 32, 14, 14)
        self.L2 = Blk(4 * f * w, 8 * f * w, 16, 16, 6, 6)
        self.L3 = Blk(8 * f * w, 8 * f * w, 16, 16, 6, 6)
        self.L4 = Blk(8 * f * w, 4 * f * w, 32, 32, 6, 6)
        self.L5 = Blk(8 * f * w, 2 * f * w, 48, 48, 14, 14)
        self.L6 = Blk(4 * f * w, w, 64, 64, 22, 22)
        self.fc1 = nn.Linear(2 * w, 4 * w)
        self.fc2 = nn.Linear(4 * w, 1)
        self._grids = _GridCache()
        self._glue = _glue_of(ops)
        self._fused = self._glue is not ops

    def forward(self, x):
        h, h_skip = self._lift(x)
        D1, D2 = h.shape[-2], h.shape[-1]
        f = self.factor
        c0, c0_skip = _twice(self._fused, self.L0, h, int(D1 * f), int(D2 * f))
        c1, c1_skip = _twice(self._fused, self.L1, c0, D1 // 2, D2 // 2)
        c2 = self.L2(c1, D1 // 4, D2 // 4)
        c3 = self.L3(c2, D1 // 4, D2 // 4)
        c4 = torch.cat([self.L4(c3, D1 // 2, D2 // 2), c1_skip], dim=1)
        c5 = torch.cat([self.L5(c4, int(D1 * f), int(D2 * f)), c0_skip], dim=1)
        return self._project([self.L6(c5, D1, D2), h_skip])


class UNO_P(_NS2DBase):
    """Navier-Stokes 2-D U-NO with factor-2 domain scaling and a lifted-input skip into the
    projection (navier_stokes_uno2d.py:24-138)."""

    def __init__(self, in_width, width, pad=0, factor=1, ops=None):
        super().__init__()
        ops = ops or _default_ops()
        Blk = ops.OperatorBlock_2D
        self.in_width, self.width, self.factor, self.padding = in_width, width, factor, pad
        w, f = width, factor
        self.fc = nn.Linear(in_width, w // 2)
        self.fc0 = nn.Linear(w // 2, w)
        self.L0 = Blk(w, 2 * f * w, 32, 32, 14, 14)
        self.L1 = Blk(2 * f * w, 4 * f * w, 16, 16, 6, 6)
        self.L2 = Blk(4 * f * w, 8 * f * w, 8, 8, 3, 3)
        self.L3 = Blk(8 * f * w, 8 * f * w, 8, 8, 3, 3)
        self.L4 = Blk(8 * f * w, 4 * f * w, 16, 16, 3, 3)
        self.L5 = Blk(8 * f * w, 2 * f * w, 32, 32, 6, 6)
        self.L6 = Blk(4 * f * w, w, 64, 64, 14, 14)
        self.fc1 = nn.Linear(2 * w, 3 * w)
        self.fc2 = nn.Linear(3 * w + w // 2, 1)
        self._grids = _GridCache()

    def forward(self, x):
        x = torch.cat((x, self.get_grid(x.shape, x.device)), dim=-1)
        x_fc = F.gelu(self.fc(x))
        p = self.padding
        h = F.pad(F.gelu(self.fc0(x_fc)).permute(0, 3, 1, 2), [p, p, p, p])
        D1, D2 = h.shape[-2], h.shape[-1]
        c0 = self.L0(h, D1 // 2, D2 // 2)
        c1 = self.L1(c0, D1 // 4, D2 // 4)
        c2 = self.L2(c1, D1 // 8, D2 // 8)
        c3 = self.L3(c2, D1 // 8, D2 // 8)
        c4 = torch.cat([self.L4(c3, D1 // 4, D2 // 4), c1], dim=1)
        c5 = torch.cat([self.L5(c4, D1 // 2, D2 // 2), c0], dim=1)
        c6 = torch.cat([self.L6(c5, D1, D2), h], dim=1)
        if p != 0:
            c6 = c6[..., p:-p, p:-p]
        t = F.gelu(self.fc1(c6.permute(0, 2, 3, 1)))
        return self.fc2(torch.cat([t, x_fc], dim=3))


class Uno3D_T10(nn.Module):
    """Navier-Stokes space-time U-NO, 7 blocks, time axis padded by int(pad*0.1*T)
    (navier_stokes_uno3d.py:412-603)."""

    def __init__(self, in_width, width, pad=2, factor=1, pad_both=False, ops=None):
        super().__init__()
        ops = ops or _default_ops()
        Blk = ops.OperatorBlock_3D
        self.in_width, self.width, self.pad, self.pad_both = in_width, width, pad, pad_both
        w, f = width, factor
        self.fc = nn.Linear(in_width, in_width * 2)
        self.fc0 = nn.Linear(in_width * 2, w)
        self.conv0 = Blk(w, 2 * f * w, 48, 48, 10, 22, 22, 5, Normalize=True)
        self.conv1 = Blk(2 * f * w, 4 * f * w, 32, 32, 10, 14, 14, 5)
        self.conv2 = Blk(4 * f * w, 8 * f * w, 16, 16, 10, 6, 6, 5)
        self.conv3 = Blk(8 * f * w, 16 * f * w, 16, 16, 10, 6, 6, 5, Normalize=True)
        self.conv6 = Blk(16 * f * w, 4 * f * w, 32, 32, 10, 6, 6, 5)
        self.conv7 = Blk(8 * f * w, 2 * f * w, 48, 48, 10, 14, 14, 5, Normalize=True)
        self.conv8 = Blk(4 * f * w, 2 * w, 64, 64, 10, 22, 22, 5)
        self.fc1 = nn.Linear(3 * w, 4 * w)
        self.fc2 = nn.Linear(4 * w, 1)
        self._grids = _GridCache()
        self._glue = _glue_of(ops)
        self._fused = self._glue is not ops

    def get_grid(self, shape, device):
        b, sx, sy, sz = shape[0], shape[1], shape[2], shape[3]

        def build():
            gx = _lin(0, 2 * np.pi, sx).reshape(1, sx, 1, 1, 1).repeat([1, 1, sy, sz, 1])
            gy = _lin(0, 2 * np.pi, sy).reshape(1, 1, sy, 1, 1).repeat([1, sx, 1, sz, 1])
            gz = _lin(0, 1, sz).reshape(1, 1, 1, sz, 1).repeat([1, sx, sy, 1, 1])
            return torch.cat((torch.sin(gx), torch.sin(gy), torch.cos(gx), torch.cos(gy), gz), dim=-1)

        return self._grids.get((sx, sy, sz), device, build).expand(b, -1, -1, -1, -1)

    @staticmethod
    def _skip(src, like):
        # trilinear with align_corners=True onto an identical grid is the identity (source index == target index,
        # interpolation weight exactly 0), which is the case for every skip of the shipped configurations
        if tuple(src.shape[2:]) == tuple(like.shape[2:]):
            return src
        return F.interpolate(src, size=tuple(like.shape[2:]), mode="trilinear", align_corners=True)

    def forward(self, x):
        # navier_stokes_uno3d.py:497-511: lift, channels-first, pad the time axis (trailing edge, or both)
        grid = self.get_grid((1,) + tuple(x.shape[1:]), x.device)[0]
        self.padding = int(self.pad * 0.1 * x.shape[3])
        lo = self.padding if self.pad_both else 0
        h, h_skip = _twice(self._fused, self._glue.lift, x, grid, self.fc.weight, self.fc.bias, self.fc0.weight, self.fc0.bias,
                           (0, 0, lo), (0, 0, self.padding))
        D1, D2, D3 = h.shape[-3], h.shape[-2], h.shape[-1]
        c0, c0_skip = _twice(self._fused, self.conv0, h, int(3 * D1 / 4), int(3 * D2 / 4), D3)
        c1, c1_skip = _twice(self._fused, self.conv1, c0, D1 // 2, D2 // 2, D3)
        c2 = self.conv2(c1, D1 // 4, D2 // 4, int(1.0 * D3))
        c3 = self.conv3(c2, D1 // 4, D2 // 4, int(1.0 * D3))
        c6 = self.conv6(c3, D1 // 2, D2 // 2, int(1.0 * D3))
        c6 = torch.cat([c6, self._skip(c1_skip, c6)], dim=1)
        c7 = self.conv7(c6, int(3 * D1 / 4), int(3 * D2 / 4), D3)
        c7 = torch.cat([c7, self._skip(c0_skip, c7)], dim=1)
        c8 = self.conv8(c7, D1, D2, D3)
        # :551-575: cat with the lifted input, crop the time padding, project
        return self._glue.project([c8, self._skip(h_skip, c8)], self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                                  (0, 0, lo), (0, 0, self.padding))
