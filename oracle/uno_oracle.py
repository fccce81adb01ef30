"""
TEST INFRASTRUCTURE ONLY -- numpy fp64 restatement of the U-NO integral-operator hot path.

This file is the parity *checker*.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``uno_b200/`` (the product) imports or calls anything in ``oracle/``.

Pinned against the real reference (``/root/reference/integral_operators.py`` imported in the build
container) by ``tests/golden/make_golden.py``; the resulting fixtures live in ``tests/golden/*.npz``
and are re-checked by ``tests/test_oracle.py`` on every run.  The reference itself has no tests or
golden vectors, so reference-generated fixtures are the only pin there is.

Every function cites the reference lines it restates (paths relative to /root/reference).
All arithmetic is float64 / complex128 unless stated.
"""
from __future__ import annotations

import itertools
import math
from typing import List, Sequence, Tuple

import numpy as np

try:  # exact erf for GELU
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover - scipy is part of the image
    _erf = np.vectorize(math.erf)


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------
def corner_slices(modes: Sequence[int]) -> List[Tuple[slice, ...]]:
    """Spectrum corner blocks in the reference's WRITE ORDER.

    1-D: one block ``[:m1]``                       (integral_operators.py:66-68)
    2-D: ``[:m1,:m2]`` then ``[-m1:,:m2]``         (integral_operators.py:198-203)
    3-D: (lo,lo) (hi,lo) (lo,hi) (hi,hi), ``[:m3]`` (integral_operators.py:410-421)
    """
    d = len(modes)
    last = slice(0, modes[-1])
    if d == 1:
        return [(last,)]
    lo = [slice(0, m) for m in modes[:-1]]
    hi = [slice(-m, None) for m in modes[:-1]]
    if d == 2:
        return [(lo[0], last), (hi[0], last)]
    if d == 3:
        return [
            (lo[0], lo[1], last),
            (hi[0], lo[1], last),
            (lo[0], hi[1], last),
            (hi[0], hi[1], last),
        ]
    raise ValueError("ndim must be 1, 2 or 3")


def _check_modes(in_dims, out_dims, modes):
    """Error behaviour of the reference: a mode count larger than the input or the
    output (half-)spectrum makes the einsum / slice-assign fail."""
    d = len(modes)
    for a in range(d):
        if a == d - 1:
            lim_in, lim_out = in_dims[a] // 2 + 1, out_dims[a] // 2 + 1
        else:
            lim_in, lim_out = in_dims[a], out_dims[a]
        if modes[a] > lim_in:
            raise ValueError(f"modes[{a}]={modes[a]} exceeds the input spectrum ({lim_in})")
        if modes[a] > lim_out:
            raise ValueError(f"modes[{a}]={modes[a]} exceeds the output spectrum ({lim_out})")


# ---------------------------------------------------------------------------------------------
# SpectralConv{1,2,3}d_Uno.forward
# ---------------------------------------------------------------------------------------------
def spectral_conv_fwd(x, weights, out_dims, modes, return_xhat=False):
    """integral_operators.py:47-72 (1-D), :181-207 (2-D), :385-427 (3-D).

    x        [B, Ci, *in_dims] real
    weights  list of 1/2/4 complex arrays [Ci, Co, *modes]  (weights1..4)
    returns  [B, Co, *out_dims] real (fp64)
    """
    x = np.asarray(x, dtype=np.float64)
    d = len(out_dims)
    in_dims = x.shape[2:]
    _check_modes(in_dims, out_dims, modes)
    axes = tuple(range(-d, 0))
    B, Co = x.shape[0], weights[0].shape[1]
    x_ft = np.fft.rfftn(x, axes=axes, norm="forward")  # :56 / :187 / :398
    out_ft = np.zeros((B, Co) + tuple(out_dims[:-1]) + (out_dims[-1] // 2 + 1,), np.complex128)
    sub = "xyz"[:d]
    expr = f"bi{sub},io{sub}->bo{sub}"  # compl_mul{1,2,3}d  :45 / :179 / :383
    for sl, w in zip(corner_slices(modes), weights):
        idx = (slice(None), slice(None)) + sl
        out_ft[idx] = np.einsum(expr, x_ft[idx], np.asarray(w, np.complex128))
    y = np.fft.irfftn(out_ft, s=tuple(out_dims), axes=axes, norm="forward")  # :71 / :206 / :424
    if return_xhat:
        return y, x_ft
    return y


def _axis_mode_index(n: int, m: int, last: bool) -> np.ndarray:
    """frequency indices kept along one axis: [0..m) on the half axis, [0..m) + [n-m..n) otherwise."""
    if last:
        return np.arange(m)
    return np.concatenate([np.arange(m), np.arange(n - m, n)])


def spectral_conv_fwd_dft(x, weights, out_dims, modes):
    """Same operator written as truncated DFT -> per-mode contraction -> truncated inverse DFT
    (the form the CUDA kernels implement).  Independent of any FFT
    library; O(N * modes) so small cases only."""
    x = np.asarray(x, dtype=np.float64)
    d = len(out_dims)
    in_dims = x.shape[2:]
    _check_modes(in_dims, out_dims, modes)
    B, Ci = x.shape[:2]
    Co = weights[0].shape[1]
    # forward truncated DFT, axis by axis (last axis first), 1/N scaling per axis
    cur = x.astype(np.complex128)
    for a in range(d - 1, -1, -1):
        n = in_dims[a]
        k = _axis_mode_index(n, modes[a], a == d - 1)
        F = np.exp(-2j * np.pi * np.outer(k, np.arange(n)) / n) / n  # [K, n]
        cur = np.moveaxis(np.tensordot(F, cur, axes=([1], [2 + a])), 0, 2 + a)
    xhat = cur  # [B,Ci,(2)m1,(2)m2,m3]
    # contraction per corner; each corner lives in a (lo|hi) half along every non-last axis
    yhat = np.zeros((B, Co) + xhat.shape[2:], np.complex128)
    sub = "xyz"[:d]
    expr = f"bi{sub},io{sub}->bo{sub}"
    halves = list(itertools.product(*[(0, 1)] * (d - 1)))  # (h1,h2): 0=lo 1=hi
    # reference order: w1 (lo,lo) w2 (hi,lo) w3 (lo,hi) w4 (hi,hi) -> first axis varies fastest
    halves.sort(key=lambda h: tuple(reversed(h)))
    for h, w in zip(halves, weights):
        idx = [slice(None), slice(None)]
        for a in range(d - 1):
            idx.append(slice(h[a] * modes[a], (h[a] + 1) * modes[a]))
        idx.append(slice(0, modes[-1]))
        idx = tuple(idx)
        yhat[idx] = np.einsum(expr, xhat[idx], np.asarray(w, np.complex128))
    # scatter to the output spectrum in write order (later corners overwrite), then inverse
    out_ft = np.zeros((B, Co) + tuple(out_dims[:-1]) + (out_dims[-1] // 2 + 1,), np.complex128)
    for h, sl in zip(halves, corner_slices(modes)):
        src = [slice(None), slice(None)]
        for a in range(d - 1):
            src.append(slice(h[a] * modes[a], (h[a] + 1) * modes[a]))
        src.append(slice(0, modes[-1]))
        out_ft[(slice(None), slice(None)) + sl] = yhat[tuple(src)]
    cur = out_ft
    for a in range(d - 1):  # unnormalised inverse C2C on the leading axes
        n = out_dims[a]
        E = np.exp(2j * np.pi * np.outer(np.arange(n), np.arange(n)) / n)
        cur = np.moveaxis(np.tensordot(E, cur, axes=([1], [2 + a])), 0, 2 + a)
    # C2R on the last axis: imaginary part of DC (and Nyquist) is dropped, interior bins doubled
    n = out_dims[-1]
    kk = np.arange(n // 2 + 1)
    c = np.full(n // 2 + 1, 2.0)
    c[0] = 1.0
    if n % 2 == 0:
        c[-1] = 1.0
    E = np.exp(2j * np.pi * np.outer(kk, np.arange(n)) / n) * c[:, None]  # [K, n]
    y = np.real(np.tensordot(cur, E, axes=([cur.ndim - 1], [0])))
    return y


# ---------------------------------------------------------------------------------------------
# backward of SpectralConv (autograd of the reference)
# ---------------------------------------------------------------------------------------------
def spectral_conv_bwd(x, weights, out_dims, modes, gy):
    """Returns (gx, [gw...]) with torch's convention for complex leaves:
    ``gw = dL/dRe(w) + i dL/dIm(w)`` (what ``weights1.grad`` holds after ``backward()``)."""
    x = np.asarray(x, dtype=np.float64)
    gy = np.asarray(gy, dtype=np.float64)
    d = len(out_dims)
    in_dims = x.shape[2:]
    axes = tuple(range(-d, 0))
    n_in = float(np.prod(in_dims))
    x_ft = np.fft.rfftn(x, axes=axes, norm="forward")
    # adjoint of the unnormalised C2R: c(k_last) * unnormalised forward transform of gy
    n = out_dims[-1]
    c = np.full(n // 2 + 1, 2.0)
    c[0] = 1.0
    if n % 2 == 0:
        c[-1] = 1.0
    g_ft = np.fft.rfftn(gy, axes=axes) * c
    slices = corner_slices(modes)
    # last-writer-wins on the output spectrum: an entry overwritten by a later corner has no gradient
    owner = -np.ones(g_ft.shape[2:], dtype=np.int64)
    for ci, sl in enumerate(slices):
        owner[sl] = ci
    sub = "xyz"[:d]
    gws = []
    gx_ft = np.zeros_like(x_ft)
    for ci, (sl, w) in enumerate(zip(slices, weights)):
        idx = (slice(None), slice(None)) + sl
        g = g_ft[idx] * (owner[sl] == ci)
        w = np.asarray(w, np.complex128)
        gws.append(np.einsum(f"bi{sub},bo{sub}->io{sub}", np.conj(x_ft[idx]), g))
        gx_ft[idx] += np.einsum(f"bo{sub},io{sub}->bi{sub}", g, np.conj(w))
    # adjoint of rfftn(norm="forward"): (1/N) Re sum_k gx_ft[k] e^{+...}; no interior doubling
    full_shape = x.shape[:2] + tuple(in_dims)
    full = np.zeros(full_shape, np.complex128)
    full[..., : in_dims[-1] // 2 + 1] = gx_ft
    gx = np.real(np.fft.ifftn(full, axes=axes)) * (n_in / n_in)
    return gx, gws


# ---------------------------------------------------------------------------------------------
# pointwise_op_2D : Conv2d(k=1) + bicubic anti-aliased resample, align_corners=True
# ---------------------------------------------------------------------------------------------
def _cubic_aa(t: np.float32) -> np.float32:
    """Keys cubic, a = -0.5 (ATen BicubicFilterFunctor / aa_filter), evaluated in fp32."""
    f = np.float32
    a = f(-0.5)
    t = f(abs(t))
    if t < f(1.0):
        return f(f(f(f(f(f(a + f(2.0)) * t) - f(a + f(3.0))) * t) * t) + f(1.0))
    if t < f(2.0):
        return f(f(f(f(f(a * t) - f(f(5.0) * a)) * t) + f(f(8.0) * a)) * t - f(f(4.0) * a))
    return f(0.0)


def bicubic_aa_matrix(n_in: int, n_out: int) -> np.ndarray:
    """Dense [n_out, n_in] matrix of ``F.interpolate(mode='bicubic', align_corners=True,
    antialias=True)`` along one axis (integral_operators.py:240-242).

    Follows ATen's fp32 arithmetic (``_compute_indices_min_size_weights_aa``): scale =
    (in-1)/(out-1), support = 2*scale when down-sampling else 2, weights normalised to sum 1.
    Returned as float64 holding the fp32 weight values."""
    f = np.float32
    R = np.zeros((n_out, n_in), np.float64)
    if n_in == n_out:
        # ATen short-circuits same-size interpolation to a copy
        np.fill_diagonal(R, 1.0)
        return R
    scale = f(n_in - 1) / f(n_out - 1) if n_out > 1 else f(0.0)
    interp = f(4.0)  # bicubic interp_size
    support = f(interp * f(0.5) * scale) if scale >= f(1.0) else f(interp * f(0.5))
    invscale = f(f(1.0) / scale) if scale >= f(1.0) else f(1.0)
    for i in range(n_out):
        center = f(scale * f(f(i) + f(0.5)))
        xmin = max(int(f(center - support + f(0.5))), 0)
        xsize = min(int(f(center + support + f(0.5))), n_in) - xmin
        ws = np.zeros(max(xsize, 0), np.float32)
        total = f(0.0)
        for j in range(xsize):
            w = _cubic_aa(f(f(f(j + xmin) - center + f(0.5)) * invscale))
            ws[j] = w
            total = f(total + w)
        if total != f(0.0):
            ws = (ws / total).astype(np.float32)
        R[i, xmin : xmin + xsize] = ws
    return R


def conv1x1(x, weight, bias):
    """nn.Conv{1,2,3}d(kernel_size=1): weight [Co,Ci,1..], bias [Co] (integral_operators.py:82,:220,:433)."""
    x = np.asarray(x, np.float64)
    w = np.asarray(weight, np.float64).reshape(weight.shape[0], weight.shape[1])
    y = np.tensordot(w, x, axes=([1], [1]))  # [Co, B, ...]
    y = np.moveaxis(y, 0, 1)
    if bias is not None:
        y = y + np.asarray(bias, np.float64).reshape((1, -1) + (1,) * (x.ndim - 2))
    return y


def pointwise_op_2d_fwd(x, conv_w, conv_b, out_dims):
    """integral_operators.py:224-243."""
    z = conv1x1(x, conv_w, conv_b)
    R1 = bicubic_aa_matrix(z.shape[2], out_dims[0])
    R2 = bicubic_aa_matrix(z.shape[3], out_dims[1])
    return np.einsum("ph,bchw,qw->bcpq", R1, z, R2)


def pointwise_op_3d_fwd(x, conv_w, conv_b, out_dims):
    """integral_operators.py:438-468: Conv3d(k=1) -> rfftn -> copy 4 low corners
    (d//2 per side, [:d3//2] on the half axis) -> irfftn(s=out) (crop / zero-pad at the END of every
    axis, backward norm) -> same-size trilinear interpolate (identity)."""
    z = conv1x1(x, conv_w, conv_b)
    d1, d2, d3 = out_dims
    ft = np.fft.rfftn(z, axes=(-3, -2, -1))
    ft_u = np.zeros_like(ft)
    h1, h2, h3 = d1 // 2, d2 // 2, d3 // 2
    lo1, hi1 = slice(0, h1), slice(-h1, None)
    lo2, hi2 = slice(0, h2), slice(-h2, None)
    l3 = slice(0, h3)
    for s1, s2 in ((lo1, lo2), (hi1, lo2), (lo1, hi2), (hi1, hi2)):
        ft_u[:, :, s1, s2, l3] = ft[:, :, s1, s2, l3]
    return np.fft.irfftn(ft_u, s=(d1, d2, d3), axes=(-3, -2, -1))


def fourier_resample_matrix(n_in: int, n_out: int) -> np.ndarray:
    """NOT the reference: the band-limited Fourier resample that the library offers as an opt-in replacement for
    pointwise_op_3D's quirky spectral resample (UNO_B200_POINTWISE3D_FIXED=1).  Real [n_out, n_in]:
    R[j, h] = (1/n_in) * sum_{|k| <= K} exp(2 pi i k (j/n_out - h/n_in)),  K = (min(n_in, n_out) - 1) // 2."""
    K = (min(n_in, n_out) - 1) // 2
    j = np.arange(n_out, dtype=np.float64)[:, None] / n_out
    h = np.arange(n_in, dtype=np.float64)[None, :] / n_in
    R = np.ones((n_out, n_in))
    for k in range(1, K + 1):
        R += 2.0 * np.cos(2.0 * np.pi * k * (j - h))
    return R / n_in


def pointwise_op_3d_fixed_fwd(x, conv_w, conv_b, out_dims):
    """Conv3d(k=1) followed by the separable band-limited Fourier resample above (the opt-in mode; not the reference)."""
    z = conv1x1(x, conv_w, conv_b)
    R = [fourier_resample_matrix(z.shape[2 + a], out_dims[a]) for a in range(3)]
    return np.einsum("pd,qe,rf,bcdef->bcpqr", R[0], R[1], R[2], z)


# ---------------------------------------------------------------------------------------------
# InstanceNorm / GELU / OperatorBlock
# ---------------------------------------------------------------------------------------------
def instance_norm(x, gamma, beta, eps=1e-5):
    """nn.InstanceNorm{1,2,3}d(affine=True, track_running_stats=False): biased variance per (b,c)
    plane (integral_operators.py:110,:270,:499)."""
    x = np.asarray(x, np.float64)
    ax = tuple(range(2, x.ndim))
    mu = x.mean(axis=ax, keepdims=True)
    var = x.var(axis=ax, keepdims=True)
    shp = (1, -1) + (1,) * (x.ndim - 2)
    return (x - mu) / np.sqrt(var + eps) * np.asarray(gamma, np.float64).reshape(shp) + np.asarray(
        beta, np.float64
    ).reshape(shp)


def gelu(x):
    """F.gelu default (exact erf form) (integral_operators.py:123,:283,:512)."""
    x = np.asarray(x, np.float64)
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def operator_block_fwd(x, weights, conv_w, conv_b, out_dims, modes, norm=None, non_lin=True):
    """OperatorBlock_{2,3}D.forward: gelu(IN?(conv(x) + w(x)))  (integral_operators.py:272-284, :501-513).
    ``norm`` is ``None`` or ``(gamma, beta)``."""
    d = len(out_dims)
    x1 = spectral_conv_fwd(x, weights, out_dims, modes)
    if d == 2:
        x2 = pointwise_op_2d_fwd(x, conv_w, conv_b, out_dims)
    elif d == 3:
        x2 = pointwise_op_3d_fwd(x, conv_w, conv_b, out_dims)
    else:
        raise ValueError("OperatorBlock_1D raises in torch >= 1.11 (SURVEY.md B.1)")
    out = x1 + x2
    if norm is not None:
        out = instance_norm(out, norm[0], norm[1])
    if non_lin:
        out = gelu(out)
    return out
