"""TEST INFRASTRUCTURE ONLY: numpy front-end to the host-emulation build of the C ABI."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from uno_b200 import _capi  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        sys.path.insert(0, HERE)
        import build as _build  # tests/hostemu/build.py

        _lib = _capi.bind(_build.build())
        assert _lib.uno_backend_name() == b"host-emulation"
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


def _ws(nbytes):
    return np.empty(max(nbytes, 8) // 4 + 64, np.float32)


def spectral_fwd(x, weights, out_dims, modes, want_xhat=True):
    L = lib()
    x = _f32(x)
    ws_ = [_c64(w) for w in weights]
    B, Ci = x.shape[:2]
    Co = ws_[0].shape[1]
    d = _capi.conv_desc(B, Ci, Co, x.shape[2:], out_dims, modes)
    _capi.check(L, L.uno_spectral_conv_check(C.byref(d)))
    y = np.empty((B, Co) + tuple(out_dims), np.float32)
    xhat = np.empty(L.uno_spectral_conv_xhat_elems(C.byref(d)), np.complex64) if want_xhat else None
    ws = _ws(L.uno_spectral_conv_workspace_bytes(C.byref(d)))
    wp = _capi.ptr_array([w.ctypes.data for w in ws_])
    _capi.check(L, L.uno_spectral_conv_fwd(C.byref(d), _p(x), wp, _p(y), _p(xhat), _p(ws), ws.nbytes, None))
    return (y, xhat) if want_xhat else y


def spectral_bwd(x_shape, weights, out_dims, modes, gy, xhat):
    L = lib()
    ws_ = [_c64(w) for w in weights]
    B, Ci = x_shape[:2]
    Co = ws_[0].shape[1]
    d = _capi.conv_desc(B, Ci, Co, x_shape[2:], out_dims, modes)
    gy = _f32(gy)
    gx = np.empty(x_shape, np.float32)
    gws = [np.empty_like(w) for w in ws_]
    ws = _ws(L.uno_spectral_conv_workspace_bytes(C.byref(d)))
    wp = _capi.ptr_array([w.ctypes.data for w in ws_])
    gwp = _capi.ptr_array([w.ctypes.data for w in gws])
    _capi.check(L, L.uno_spectral_conv_bwd(C.byref(d), _p(gy), _p(xhat), wp, _p(gx), gwp, 0, _p(ws), ws.nbytes, None))
    return gx, gws


def pointwise_fwd(x, conv_w, conv_b, out_dims, want_saved=True):
    L = lib()
    x = _f32(x)
    cw = _f32(conv_w).reshape(conv_w.shape[0], conv_w.shape[1])
    cb = _f32(conv_b)
    B, Ci = x.shape[:2]
    Co = cw.shape[0]
    d = _capi.conv_desc(B, Ci, Co, x.shape[2:], out_dims)
    z = np.empty((B, Co) + tuple(out_dims), np.float32)
    n = L.uno_pointwise_saved_elems(C.byref(d))
    saved = np.empty(n, np.float32) if (want_saved and n) else None
    ws = _ws(L.uno_pointwise_workspace_bytes(C.byref(d)))
    _capi.check(L, L.uno_pointwise_fwd(C.byref(d), _p(x), _p(cw), _p(cb), _p(z), _p(saved), _p(ws), ws.nbytes, None))
    return z, saved


def pointwise_bwd(x, conv_w, out_dims, gz, saved):
    L = lib()
    x = _f32(x)
    cw = _f32(conv_w).reshape(conv_w.shape[0], conv_w.shape[1])
    B, Ci = x.shape[:2]
    Co = cw.shape[0]
    d = _capi.conv_desc(B, Ci, Co, x.shape[2:], out_dims)
    gz = _f32(gz)
    gx = np.empty_like(x)
    gw = np.empty_like(cw)
    gb = np.empty(Co, np.float32)
    ws = _ws(L.uno_pointwise_workspace_bytes(C.byref(d)))
    _capi.check(L, L.uno_pointwise_bwd(C.byref(d), _p(gz), _p(x), _p(saved), _p(cw), _p(gx), _p(gw), _p(gb), _p(ws), ws.nbytes, None))
    return gx, gw, gb


def block_fwd(x, weights, conv_w, conv_b, out_dims, modes, gamma=None, beta=None, non_lin=True, train=True):
    L = lib()
    x = _f32(x)
    ws_ = [_c64(w) for w in weights]
    cw = _f32(conv_w).reshape(conv_w.shape[0], conv_w.shape[1])
    cb = _f32(conv_b)
    B, Ci = x.shape[:2]
    Co = cw.shape[0]
    normalize = gamma is not None
    g = _f32(gamma) if normalize else None
    bt = _f32(beta) if normalize else None
    cd = _capi.conv_desc(B, Ci, Co, x.shape[2:], out_dims, modes)
    bd = _capi.block_desc(cd, normalize, non_lin)
    y = np.empty((B, Co) + tuple(out_dims), np.float32)
    need_pre = train and (normalize or non_lin)
    pre = np.empty_like(y) if need_pre else None
    stats = np.empty((B * Co, 2), np.float32) if (need_pre and normalize) else None
    xhat = np.empty(L.uno_spectral_conv_xhat_elems(C.byref(cd)), np.complex64) if train else None
    n = L.uno_pointwise_saved_elems(C.byref(cd))
    saved = np.empty(n, np.float32) if (train and n) else None
    ws = _ws(L.uno_operator_block_workspace_bytes(C.byref(bd)))
    wp = _capi.ptr_array([w.ctypes.data for w in ws_])
    _capi.check(
        L,
        L.uno_operator_block_fwd(C.byref(bd), _p(x), wp, _p(cw), _p(cb), _p(g), _p(bt), _p(y), _p(xhat), _p(saved), _p(pre), _p(stats), _p(ws), ws.nbytes, None),
    )
    return y, dict(xhat=xhat, saved=saved, pre=pre, stats=stats)


def block_bwd(x, weights, conv_w, out_dims, modes, gy, ctx, gamma=None, beta=None, non_lin=True, gy2=None, gy_bs=0, gy2_bs=0):
    """gy (+ gy2): upstream gradient(s); gy_bs / gy2_bs: batch strides in floats when gy / gy2 are channel slices of wider
    arrays (pass the wide array's sliced VIEW: the pointer of its first element is taken), 0 = contiguous."""
    L = lib()
    x = _f32(x)
    ws_ = [_c64(w) for w in weights]
    cw = _f32(conv_w).reshape(conv_w.shape[0], conv_w.shape[1])
    B, Ci = x.shape[:2]
    Co = cw.shape[0]
    normalize = gamma is not None
    g = _f32(gamma) if normalize else None
    bt = _f32(beta) if normalize else None
    cd = _capi.conv_desc(B, Ci, Co, x.shape[2:], out_dims, modes)
    bd = _capi.block_desc(cd, normalize, non_lin)
    if gy_bs == 0:
        gy = _f32(gy)
    if gy2 is not None and gy2_bs == 0:
        gy2 = _f32(gy2)
    gx = np.empty_like(x)
    gws = [np.empty_like(w) for w in ws_]
    gcw = np.empty_like(cw)
    gcb = np.empty(Co, np.float32)
    gg = np.empty(Co, np.float32) if normalize else None
    gb = np.empty(Co, np.float32) if normalize else None
    ws = _ws(L.uno_operator_block_workspace_bytes(C.byref(bd)))
    wp = _capi.ptr_array([w.ctypes.data for w in ws_])
    gwp = _capi.ptr_array([w.ctypes.data for w in gws])
    _capi.check(
        L,
        L.uno_operator_block_bwd2(
            C.byref(bd), _p(gy), gy_bs, _p(gy2), gy2_bs, _p(x), _p(ctx["xhat"]), _p(ctx["saved"]), _p(ctx["pre"]), _p(ctx["stats"]), wp,
            _p(cw), _p(g), _p(bt), _p(gx), gwp, _p(gcw), _p(gcb), _p(gg), _p(gb), _p(ws), ws.nbytes, None,
        ),
    )
    return gx, gws, gcw, gcb, gg, gb


# ---- model glue --------------------------------------------------------------------------------------
def lift_fwd(a, grid, w_a, b_a, w_b, b_b, pad_lo, pad_hi):
    L = lib()
    a, grid, w_a, b_a, w_b, b_b = (_f32(t) for t in (a, grid, w_a, b_a, w_b, b_b))
    dims = a.shape[1:-1]
    d = _capi.lift_desc(a.shape[0], dims, pad_lo, pad_hi, a.shape[-1], grid.shape[-1], w_a.shape[0], w_b.shape[0])
    _capi.check(L, L.uno_lift_check(C.byref(d)))
    out_dims = tuple(n + lo + hi for n, lo, hi in zip(dims, pad_lo, pad_hi))
    h = np.full((a.shape[0], w_b.shape[0]) + out_dims, np.nan, np.float32)
    _capi.check(L, L.uno_lift_fwd(C.byref(d), _p(a), _p(grid), _p(w_a), _p(b_a), _p(w_b), _p(b_b), _p(h), None))
    return h


def lift_bwd(gh, a, grid, w_a, b_a, w_b, b_b, pad_lo, pad_hi, want_ga=True, gh2=None):
    L = lib()
    gh, a, grid, w_a, b_a, w_b, b_b = (_f32(t) for t in (gh, a, grid, w_a, b_a, w_b, b_b))
    gh2 = _f32(gh2) if gh2 is not None else None
    d = _capi.lift_desc(a.shape[0], a.shape[1:-1], pad_lo, pad_hi, a.shape[-1], grid.shape[-1], w_a.shape[0], w_b.shape[0])
    ga = np.full_like(a, np.nan) if want_ga else None
    outs = [np.full_like(t, np.nan) for t in (w_a, b_a, w_b, b_b)]
    _capi.check(L, L.uno_lift_bwd2(C.byref(d), _p(gh), _p(gh2), _p(a), _p(grid), _p(w_a), _p(b_a), _p(w_b), _p(b_b), _p(ga),
                                   *[_p(o) for o in outs], None))
    return (ga, *outs)


def project_fwd(srcs, w1, b1, w2, b2, crop_lo, crop_hi, want_pre=False):
    L = lib()
    srcs = [_f32(s) for s in srcs]
    w1, b1, w2, b2 = (_f32(t) for t in (w1, b1, w2, b2))
    full = srcs[0].shape[2:]
    dims = tuple(n - lo - hi for n, lo, hi in zip(full, crop_lo, crop_hi))
    d = _capi.project_desc(srcs[0].shape[0], dims, crop_lo, crop_hi, [s.shape[1] for s in srcs], w1.shape[0], w2.shape[0])
    _capi.check(L, L.uno_project_check(C.byref(d)))
    out = np.full((srcs[0].shape[0],) + dims + (w2.shape[0],), np.nan, np.float32)
    sp = _capi.ptr_array([s.ctypes.data for s in srcs])
    pre = np.full((w1.shape[0], out.size // w2.shape[0]), np.nan, np.float32) if want_pre else None
    _capi.check(L, L.uno_project_fwd(C.byref(d), sp, _p(w1), _p(b1), _p(w2), _p(b2), _p(out), _p(pre), None))
    return (out, pre) if want_pre else out


def project_bwd(gout, srcs, w1, b1, w2, crop_lo, crop_hi, pre=None):
    L = lib()
    srcs = [_f32(s) for s in srcs]
    gout, w1, b1, w2 = (_f32(t) for t in (gout, w1, b1, w2))
    full = srcs[0].shape[2:]
    dims = tuple(n - lo - hi for n, lo, hi in zip(full, crop_lo, crop_hi))
    d = _capi.project_desc(srcs[0].shape[0], dims, crop_lo, crop_hi, [s.shape[1] for s in srcs], w1.shape[0], w2.shape[0])
    gs = [np.full_like(s, np.nan) for s in srcs]
    gw1, gb1, gw2 = np.full_like(w1, np.nan), np.full_like(b1, np.nan), np.full_like(w2, np.nan)
    gb2 = np.full(w2.shape[0], np.nan, np.float32)
    sp = _capi.ptr_array([s.ctypes.data for s in srcs])
    gp = _capi.ptr_array([g.ctypes.data for g in gs])
    pre = _f32(pre) if pre is not None else None
    _capi.check(L, L.uno_project_bwd(C.byref(d), _p(gout), sp, _p(pre), _p(w1), _p(b1), _p(w2), gp, _p(gw1), _p(gb1), _p(gw2), _p(gb2), None))
    return gs, gw1, gb1, gw2, gb2


# ---- training-step ops -------------------------------------------------------------------------------------
def adam_step(params, grads, exp_avgs, exp_avg_sqs, max_sqs, step, lr, betas, eps, weight_decay, amsgrad):
    """In-place update of numpy arrays (float32 or complex64) through uno_adam_step."""
    L = lib()
    n = len(params)
    arr = (_capi.AdamTensor * n)()
    for i in range(n):
        cx = np.iscomplexobj(params[i])
        arr[i].param = params[i].ctypes.data
        arr[i].grad = grads[i].ctypes.data
        arr[i].exp_avg = exp_avgs[i].ctypes.data
        arr[i].exp_avg_sq = exp_avg_sqs[i].ctypes.data
        arr[i].max_exp_avg_sq = max_sqs[i].ctypes.data if max_sqs is not None else None
        arr[i].numel = params[i].size * (2 if cx else 1)
        arr[i].is_complex = int(cx)
    h = _capi.AdamHyper(lr, betas[0], betas[1], eps, weight_decay, int(amsgrad), step)
    _capi.check(L, L.uno_adam_step(arr, n, C.byref(h), None))


def lp_loss(x, y, reduction, gl):
    L = lib()
    x, y, gl = _f32(x), _f32(y), _f32(gl).reshape(-1)
    B = x.shape[0]
    N = x.size // B
    loss = np.full(B if reduction == 0 else 1, np.nan, np.float32)
    norms = np.full((B, 2), np.nan, np.float32)
    ws = np.empty(2 * B, np.float64)
    _capi.check(L, L.uno_lp_loss_fwd(_p(x), _p(y), B, N, reduction, _p(loss), _p(norms), _p(ws), ws.nbytes, None))
    gx = np.full_like(x, np.nan)
    _capi.check(L, L.uno_lp_loss_bwd(_p(x), _p(y), _p(norms), _p(gl), B, N, reduction, _p(gx), None))
    return loss, gx
