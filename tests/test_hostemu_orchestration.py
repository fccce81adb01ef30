"""CPU: the orchestration behind the C ABI (uno_api.cpp: plan matrices, strides, corner maps,
adjoints, epilogue selection) run on the host-emulation backend, against the golden fixtures and the
oracle.  The CUDA kernels themselves are covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest

from cases import BLOCK_CASES, POINTWISE_CASES, SPECTRAL_CASES
from conftest import BWD_TOL, FWD_TOL, ROOT, rel_err
from oracle import uno_oracle as orc

sys.path.insert(0, os.path.join(ROOT, "tests", "hostemu"))
import emu  # noqa: E402


@pytest.mark.parametrize("name", list(SPECTRAL_CASES))
def test_spectral(name, golden):
    B, Ci, Co, idim, odim, modes = SPECTRAL_CASES[name]
    g = golden("spectral")
    ws = [g[f"{name}.w{i + 1}"] for i in range(2 ** (len(idim) - 1))]
    y, xhat = emu.spectral_fwd(g[f"{name}.x"], ws, odim, modes)
    assert rel_err(y, g[f"{name}.y"]) < FWD_TOL
    gx, gws = emu.spectral_bwd(g[f"{name}.x"].shape, ws, odim, modes, g[f"{name}.gy"], xhat)
    assert rel_err(gx, g[f"{name}.gx"]) < BWD_TOL
    for i, gw in enumerate(gws):
        assert rel_err(gw, g[f"{name}.gw{i + 1}"]) < BWD_TOL
    # the saved spectrum is the reference's x_ft on the kept modes
    _, x_ft = orc.spectral_conv_fwd(g[f"{name}.x"], ws, odim, modes, return_xhat=True)
    nd = len(idim)
    idx = [np.concatenate([np.arange(m), np.arange(n - m, n)]) for m, n in zip(modes[:-1], idim[:-1])] + [np.arange(modes[-1])]
    kept = x_ft[(slice(None), slice(None)) + np.ix_(*idx)] if nd > 1 else x_ft[..., : modes[-1]]
    assert rel_err(xhat.reshape(kept.shape), kept) < 1e-5


@pytest.mark.parametrize("name", list(POINTWISE_CASES))
def test_pointwise(name, golden):
    B, Ci, Co, idim, odim = POINTWISE_CASES[name]
    g = golden("pointwise")
    for want_saved in (True, False):
        z, saved = emu.pointwise_fwd(g[f"{name}.x"], g[f"{name}.cw"], g[f"{name}.cb"], odim, want_saved)
        assert rel_err(z, g[f"{name}.y"]) < FWD_TOL
        gx, gw, gb = emu.pointwise_bwd(g[f"{name}.x"], g[f"{name}.cw"], odim, g[f"{name}.gy"], saved)
        assert rel_err(gx, g[f"{name}.gx"]) < BWD_TOL
        assert rel_err(gw, g[f"{name}.gcw"].reshape(Co, Ci)) < BWD_TOL
        assert rel_err(gb, g[f"{name}.gcb"]) < BWD_TOL


@pytest.mark.parametrize("name", list(BLOCK_CASES))
def test_block(name, golden):
    B, Ci, Co, idim, odim, modes, norm, nl = BLOCK_CASES[name]
    g = golden("blocks")
    nd = len(idim)
    P = lambda k: g[f"{name}.param.{k}"]
    G = lambda k: g[f"{name}.grad.{k}"]
    ws = [P(f"conv.weights{i + 1}") for i in range(2 ** (nd - 1))]
    ga, be = (P("normalize_layer.weight"), P("normalize_layer.bias")) if norm else (None, None)
    y, ctx = emu.block_fwd(g[f"{name}.x"], ws, P("w.conv.weight"), P("w.conv.bias"), odim, modes, ga, be, nl)
    assert rel_err(y, g[f"{name}.y"]) < FWD_TOL
    y_inf, _ = emu.block_fwd(g[f"{name}.x"], ws, P("w.conv.weight"), P("w.conv.bias"), odim, modes, ga, be, nl, train=False)
    assert rel_err(y_inf, y) < 1e-6
    gx, gws, gcw, gcb, gg, gb = emu.block_bwd(g[f"{name}.x"], ws, P("w.conv.weight"), odim, modes, g[f"{name}.gy"], ctx, ga, be, nl)
    assert rel_err(gx, g[f"{name}.gx"]) < BWD_TOL
    for i, gw in enumerate(gws):
        assert rel_err(gw, G(f"conv.weights{i + 1}")) < BWD_TOL
    assert rel_err(gcw, G("w.conv.weight").reshape(Co, Ci)) < BWD_TOL
    scale = float(np.abs(gcw).max())
    assert np.abs(gcb - G("w.conv.bias")).max() < BWD_TOL * max(float(np.abs(G("w.conv.bias")).max()), scale)
    if norm:
        assert rel_err(gg, G("normalize_layer.weight")) < BWD_TOL
        assert rel_err(gb, G("normalize_layer.bias")) < BWD_TOL


@pytest.mark.parametrize("name", list(BLOCK_CASES))
def test_block_backward_with_two_strided_gradient_sources(name, golden):
    """uno_operator_block_bwd2: the upstream gradient as the SUM of two tensors that are channel slices of wider ones (what
    autograd hands a block whose output feeds a skip concatenation and another block) must give the gradients of the plain call."""
    B, Ci, Co, idim, odim, modes, norm, nl = BLOCK_CASES[name]
    g = golden("blocks")
    nd = len(idim)
    P = lambda k: g[f"{name}.param.{k}"]
    ws = [P(f"conv.weights{i + 1}") for i in range(2 ** (nd - 1))]
    ga, be = (P("normalize_layer.weight"), P("normalize_layer.bias")) if norm else (None, None)
    _, ctx = emu.block_fwd(g[f"{name}.x"], ws, P("w.conv.weight"), P("w.conv.bias"), odim, modes, ga, be, nl)
    gy = g[f"{name}.gy"].astype(np.float32)
    want = emu.block_bwd(g[f"{name}.x"], ws, P("w.conv.weight"), odim, modes, gy, ctx, ga, be, nl)
    rng = np.random.default_rng(5)
    part = rng.standard_normal(gy.shape).astype(np.float32)
    wide_a = rng.standard_normal((B, Co + 3) + tuple(odim)).astype(np.float32)      # gradient of cat([y, skip(3 ch)], dim=1)
    wide_b = rng.standard_normal((B, 2 + Co) + tuple(odim)).astype(np.float32)      # gradient of cat([other(2 ch), y], dim=1)
    wide_a[:, :Co] = part
    wide_b[:, 2:] = gy - part
    n_out = int(np.prod(odim))
    got = emu.block_bwd(g[f"{name}.x"], ws, P("w.conv.weight"), odim, modes, wide_a[:, :Co], ctx, ga, be, nl,
                        gy2=wide_b[:, 2:], gy_bs=(Co + 3) * n_out, gy2_bs=(2 + Co) * n_out)
    def flat(t):
        gx, gws, gcw, gcb, gg, gb = t
        return [gx, *gws, gcw, gcb] + ([gg, gb] if norm else [])
    for a, b in zip(flat(got), flat(want)):
        a, b = np.asarray(a), np.asarray(b)
        assert np.abs(a - b).max() <= 2e-6 * max(float(np.abs(b).max()), 1e-6) + 1e-7
    # one strided source alone
    got1 = emu.block_bwd(g[f"{name}.x"], ws, P("w.conv.weight"), odim, modes, wide_a[:, :Co], ctx, ga, be, nl, gy_bs=(Co + 3) * n_out)
    want1 = emu.block_bwd(g[f"{name}.x"], ws, P("w.conv.weight"), odim, modes, part, ctx, ga, be, nl)
    for a, b in zip(flat(got1), flat(want1)):
        assert np.abs(np.asarray(a) - np.asarray(b)).max() <= 2e-6 * max(float(np.abs(np.asarray(b)).max()), 1e-6) + 1e-7


def test_error_codes():
    x = np.zeros((1, 2, 8, 8), np.float32)
    with pytest.raises(RuntimeError, match="exceeds the input spectrum"):
        emu.spectral_fwd(x, [np.zeros((2, 2, 3, 6), np.complex64)] * 2, (8, 8), (3, 6))
    with pytest.raises(RuntimeError, match="exceeds the output spectrum"):
        emu.spectral_fwd(x, [np.zeros((2, 2, 5, 3), np.complex64)] * 2, (4, 8), (5, 3))
    with pytest.raises(RuntimeError, match="pointwise_op_1D"):
        emu.pointwise_fwd(np.zeros((1, 2, 8), np.float32), np.zeros((2, 2), np.float32), np.zeros(2, np.float32), (4,))


def test_profile_levels_report_labels_and_algorithmic_bytes():
    """bench.py's per-U-level roofline: one scope per fused spectral-conv call shape and direction, charged with the
    algorithmic bytes / contraction flops of SURVEY.md 8(d)."""
    import ctypes as C
    import json

    import emu

    L = emu.lib()
    rng = np.random.default_rng(0)
    B, Ci, Co, idim, odim, modes = 2, 3, 5, (12, 10), (8, 14), (3, 4)
    x = rng.standard_normal((B, Ci) + idim).astype(np.float32)
    ws = [(rng.standard_normal((Ci, Co) + modes) + 1j * rng.standard_normal((Ci, Co) + modes)).astype(np.complex64) for _ in range(2)]
    L.uno_profile_enable(1)
    try:
        for _ in range(2):
            y, xhat = emu.spectral_fwd(x, ws, odim, modes)
        emu.spectral_bwd(x.shape, ws, odim, modes, np.ones_like(y), xhat)
        n = L.uno_profile_report_levels(None, 0)
        buf = C.create_string_buffer(n + 16)
        L.uno_profile_report_levels(buf, n + 16)
    finally:
        L.uno_profile_enable(0)
    rep = json.loads(buf.value.decode())
    fwd = rep["spectral fwd B=2 3->5 [12,10]->[8,14] modes=[3,4]"]
    bwd = rep["spectral bwd B=2 3->5 [12,10]->[8,14] modes=[3,4]"]
    n_in, n_out, M, nW = 12 * 10, 8 * 14, 3 * 4, 2
    fwd_bytes = 4 * B * (Ci * n_in + Co * n_out) + 8 * nW * Ci * Co * M
    bwd_bytes = 4 * B * (Co * n_out + Ci * n_in) + 8 * B * Ci * nW * M + 16 * nW * Ci * Co * M
    assert fwd["calls"] == 2 and fwd["bytes"] == 2 * fwd_bytes and fwd["flops"] == 2 * 8 * B * Ci * Co * nW * M
    assert bwd["calls"] == 1 and bwd["bytes"] == bwd_bytes and bwd["flops"] == 2 * 8 * B * Ci * Co * nW * M
    # profiling off: nothing is recorded
    emu.spectral_fwd(x, ws, odim, modes)
    L.uno_profile_enable(1)
    n = L.uno_profile_report_levels(None, 0)
    L.uno_profile_enable(0)
    assert n == 2


def test_bench_spectral_levels_from_report():
    """bench.py turns uno_profile_report_levels' JSON into the per-U-level roofline entries of its output line."""
    import ctypes as C
    import importlib.util
    import json

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rep = {
        "spectral fwd B=32 32->64 [481,481]->[240,240] modes=[18,18]": {"calls": 2, "launches": 10, "ms": 2.0, "bytes": 2.0e9, "flops": 4.0e9},
        "spectral bwd B=32 32->64 [481,481]->[240,240] modes=[18,18]": {"calls": 2, "launches": 12, "ms": 0.0, "bytes": 1.0, "flops": 1.0},
    }
    payload = json.dumps(rep).encode()

    class FakeLib:
        @staticmethod
        def uno_profile_report_levels(buf, cap):
            if buf is not None:
                C.memmove(buf, payload, min(len(payload), cap - 1))
            return len(payload)

    out = bench.spectral_levels(FakeLib, 2, 6500.0)
    assert len(out) == 1                                   # the entry without a time is dropped
    e = out[0]
    assert e["level"].startswith("spectral fwd") and e["calls_per_step"] == 1 and e["launches_per_call"] == 5
    assert abs(e["ms_per_call"] - 1.0) < 1e-12 and abs(e["GBps"] - 1000.0) < 1e-6 and abs(e["hbm_frac"] - 1000.0 / 6500.0) < 1e-9
    assert abs(e["contraction_TFLOPs"] - 2.0) < 1e-9 and 0 < e["tensor_frac"] < 1

    class Broken:
        @staticmethod
        def uno_profile_report_levels(buf, cap):
            raise OSError("no such symbol")

    assert "error" in bench.spectral_levels(Broken, 2, 6500.0)


@pytest.mark.parametrize("idim,odim", [((8, 6, 9), (6, 6, 7)), ((6, 5, 7), (9, 8, 10)), ((7, 7, 7), (7, 7, 7)), ((8, 8, 8), (4, 12, 8))])
def test_pointwise3d_fixed_mode(idim, odim):
    """switch pointwise3d_fixed=1 (opt-in, not the reference): pointwise_op_3D with the band-limited Fourier resample --
    against the separable Dirichlet-kernel oracle, with the properties the quirky operator lacks (a constant stays the same
    constant; a trigonometric polynomial inside the kept band is reproduced exactly on the new grid) and the adjoint backward."""
    rng = np.random.default_rng(3)
    B, Ci, Co = 2, 3, 2
    x = rng.standard_normal((B, Ci) + idim).astype(np.float32)
    cw = rng.standard_normal((Co, Ci)).astype(np.float32)
    cb = rng.standard_normal(Co).astype(np.float32)
    assert emu.lib().uno_config_set(b"pointwise3d_fixed", 1) == 0
    try:
        _pointwise3d_fixed_checks(rng, x, cw, cb, idim, odim, B, Ci, Co)
    finally:
        assert emu.lib().uno_config_set(b"pointwise3d_fixed", 0) == 0
    # the default (reference) behaviour is untouched once the switch is off
    z_ref, _ = emu.pointwise_fwd(x, cw, cb, odim, True)
    assert rel_err(z_ref, orc.pointwise_op_3d_fwd(x, cw, cb, odim)) < FWD_TOL


def _pointwise3d_fixed_checks(rng, x, cw, cb, idim, odim, B, Ci, Co):
    z, saved = emu.pointwise_fwd(x, cw, cb, odim, True)
    z_or = orc.pointwise_op_3d_fixed_fwd(x, cw, cb, odim)
    assert rel_err(z, z_or) < FWD_TOL, rel_err(z, z_or)
    # constants: conv of a constant field is a constant per channel, and the resample keeps it (gain 1)
    xc = np.ones_like(x)
    zc, _ = emu.pointwise_fwd(xc, cw, cb, odim, True)
    want = (cw.sum(axis=1) + cb).reshape(1, Co, 1, 1, 1)
    assert np.abs(zc - want).max() < 2e-5 * np.abs(want).max()
    # a band-limited plane wave is reproduced on the new grid
    K = [(min(a, b) - 1) // 2 for a, b in zip(idim, odim)]
    k = [min(1, K[0]), -min(1, K[1]), min(2, K[2])]
    def wave(dims):
        g = np.meshgrid(*[np.arange(n) / n for n in dims], indexing="ij")
        return np.cos(2 * np.pi * (k[0] * g[0] + k[1] * g[1] + k[2] * g[2]) + 0.3)
    xw = np.broadcast_to(wave(idim), (1, 1) + idim).astype(np.float32)
    zw, _ = emu.pointwise_fwd(xw, np.ones((1, 1), np.float32), np.zeros(1, np.float32), odim, True)
    assert np.abs(zw[0, 0] - wave(odim)).max() < 2e-5
    # backward = adjoint: <R(Wx + b), g> differentiated by hand
    gz = rng.standard_normal(z.shape).astype(np.float32)
    gx, gcw, gcb = emu.pointwise_bwd(x, cw, odim, gz, saved)
    R = [orc.fourier_resample_matrix(idim[a], odim[a]) for a in range(3)]
    gt = np.einsum("pd,qe,rf,bcpqr->bcdef", R[0], R[1], R[2], gz.astype(np.float64))      # R^T g on the conv output grid
    gx_or = np.einsum("oc,bodef->bcdef", cw.astype(np.float64), gt)
    gcw_or = np.einsum("bodef,bcdef->oc", gt, x.astype(np.float64))
    gcb_or = gt.sum(axis=(0, 2, 3, 4))
    assert rel_err(gx, gx_or) < BWD_TOL and rel_err(gcw.reshape(Co, Ci), gcw_or) < BWD_TOL and rel_err(gcb, gcb_or) < BWD_TOL
