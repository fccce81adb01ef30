"""CPU: the C-ABI library loads and exports every symbol of include/uno_b200.h, and the host-side
planning code (twiddle / mode-map matrices, resample bands) matches numpy restatements."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

from cases import RESAMPLE_PAIRS
from conftest import ROOT
from oracle import uno_oracle as orc
from uno_b200 import _capi

sys.path.insert(0, os.path.join(ROOT, "tests", "hostemu"))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "uno_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(uno_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_prototypes_agree():
    assert set(_declared_symbols()) == set(_capi.SYMBOLS)


def test_product_library_exports_every_symbol():
    """No compute call: only that the nvcc-built library loads (without a GPU) and resolves the ABI."""
    from uno_b200 import build

    lib = _capi.bind(build.build())
    for name in _declared_symbols():
        assert hasattr(lib, name)
    assert lib.uno_backend_name() == b"cuda-sm100a"
    assert lib.uno_version() >= 100


@pytest.fixture(scope="module")
def lib():
    import emu

    return emu.lib()


def _call(fn, shape, *args):
    out = np.zeros(shape, np.float32)
    assert fn(*args, out.ctypes.data_as(C.c_void_p)) == 0
    return out


@pytest.mark.parametrize("n,m", [(64, 20), (481, 18), (223, 8), (13, 5), (83, 5), (16, 9)])
def test_last_axis_matrices(lib, n, m):
    scale = 1.0 / (n * 7)
    A = _call(lib.uno_plan_dft_last_analysis, (n, 2 * m), n, m, scale)
    k = np.arange(m)
    F = np.exp(-2j * np.pi * np.outer(np.arange(n), k) / n) * scale
    assert np.abs(A[:, 0::2] - F.real).max() < 1e-7 * scale * 2 + 1e-12
    assert np.abs(A[:, 1::2] - F.imag).max() < 1e-7 * scale * 2 + 1e-12
    # analysis == rfft on the kept bins
    x = np.random.default_rng(0).standard_normal(n)
    X = (x @ A.astype(np.float64)).reshape(m, 2)
    ref = np.fft.rfft(x)[:m] * scale
    assert np.abs(X[:, 0] + 1j * X[:, 1] - ref).max() < 1e-6 * np.abs(ref).max()
    # synthesis == irfft(n) of the zero-padded spectrum (unnormalised), incl. dropped Im of DC/Nyquist
    Smat = _call(lib.uno_plan_dft_last_synthesis, (2 * m, n), n, m, 1.0, 1)
    spec = np.random.default_rng(1).standard_normal((m, 2))
    full = np.zeros(n // 2 + 1, complex)
    full[:m] = spec[:, 0] + 1j * spec[:, 1]
    ref = np.fft.irfft(full, n=n) * n
    got = spec.reshape(-1) @ Smat.astype(np.float64)
    assert np.abs(got - ref).max() < 2e-6 * np.abs(ref).max()
    # adjoint variant (no hermitian doubling) is the transpose of the analysis matrix up to scale
    St = _call(lib.uno_plan_dft_last_synthesis, (2 * m, n), n, m, scale, 0)
    assert np.abs(St - A.T).max() < 1e-12


@pytest.mark.parametrize("n,m", [(64, 20), (481, 18), (16, 6), (8, 6), (10, 6)])
def test_mid_axis_matrices(lib, n, m):
    A = _call(lib.uno_plan_dft_mid_analysis, (2 * m, n, 2), n, m)
    Ac = A[..., 0] + 1j * A[..., 1]
    k = np.concatenate([np.arange(m), np.arange(n - m, n)])
    assert np.abs(Ac - np.exp(-2j * np.pi * np.outer(k, np.arange(n)) / n)).max() < 2e-7
    Sy = _call(lib.uno_plan_dft_mid_synthesis, (n, 2 * m, 2), n, m)
    Sc = Sy[..., 0] + 1j * Sy[..., 1]
    # reference semantics: scatter lo block then hi block into a length-n spectrum (hi overwrites), ifft*n
    rng = np.random.default_rng(2)
    v = rng.standard_normal(2 * m) + 1j * rng.standard_normal(2 * m)
    spec = np.zeros(n, complex)
    spec[:m] = v[:m]
    spec[n - m:] = v[m:]
    ref = np.fft.ifft(spec) * n
    assert np.abs(Sc @ v - ref).max() < 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("n_in,n_out", [(16, 12), (8, 16), (12, 12), (16, 8), (64, 48), (5, 9)])
def test_spectral_resample_matrix(lib, n_in, n_out):
    L = _call(lib.uno_plan_sr_mid, (n_out, n_in, 2), n_in, n_out)
    Lc = L[..., 0] + 1j * L[..., 1]
    # restatement of one leading axis of pointwise_op_3D (integral_operators.py:448-463)
    h = n_out // 2
    x = np.random.default_rng(3).standard_normal(n_in) + 1j * np.random.default_rng(4).standard_normal(n_in)
    ft = np.fft.fft(x)
    ft_u = np.zeros_like(ft)
    ft_u[:h] = ft[:h]
    ft_u[-h:] = ft[-h:] if h else ft
    spec = np.zeros(n_out, complex)
    c = min(n_in, n_out)
    spec[:c] = ft_u[:c]
    ref = np.fft.ifft(spec) * n_out
    assert np.abs(Lc @ x - ref).max() < 2e-5 * max(np.abs(ref).max(), 1.0)
    assert lib.uno_plan_sr_last_modes(n_in, n_out) == max(0, min(n_out // 2, n_in // 2 + 1))


@pytest.mark.parametrize("pair", RESAMPLE_PAIRS + ((8, 4), (4, 8), (9, 30), (30, 9), (7, 7), (256, 64)))
def test_bicubic_bands(lib, pair):
    a, b = pair
    R = _call(lib.uno_plan_bicubic_aa, (b, a), a, b, 0)
    assert np.abs(R - orc.bicubic_aa_matrix(a, b)).max() < 1e-7      # same fp32 recipe in C++ and numpy
    assert np.abs(R.sum(1) - 1).max() < 1e-6                          # rows sum to one -> bias passes through
    Rt = _call(lib.uno_plan_bicubic_aa, (a, b), a, b, 1)
    assert np.array_equal(Rt, R.T)


@pytest.mark.parametrize("pair", RESAMPLE_PAIRS + ((7, 12), (12, 7), (30, 9), (9, 30), (5, 40), (300, 20)))
@pytest.mark.parametrize("transpose", [0, 1])
def test_band_groups_reproduce_the_band(pair, transpose, lib):
    """The register-blocked image the fused resample kernel consumes expands to exactly the banded matrix."""
    L = lib
    n_in, n_out = pair
    shape = (n_in, n_out) if transpose else (n_out, n_in)
    dense = np.zeros(shape, np.float32)
    assert L.uno_plan_bicubic_aa(n_in, n_out, transpose, dense.ctypes.data_as(C.c_void_p)) == 0
    grouped = np.zeros(shape, np.float32)
    gw = (C.c_int * 2)()
    assert L.uno_plan_band_groups(n_in, n_out, transpose, grouped.ctypes.data_as(C.c_void_p), gw) == 0
    if gw[0] == 0:      # band too wide for the register-blocked kernel: the generic kernel takes it
        assert max(n_in, n_out) / min(n_in, n_out) > 2.5 or min(shape) < 8
        return
    assert (gw[0], gw[1]) in ((8, 8), (4, 8), (4, 16))
    assert np.array_equal(grouped, dense)


@pytest.mark.parametrize("n_in,n_out", [(8, 6), (6, 9), (7, 7), (13, 8), (5, 16)])
def test_sr_mid_fixed_is_the_dirichlet_kernel(lib, n_in, n_out):
    from oracle import uno_oracle as orc

    L = _call(lib.uno_plan_sr_mid_fixed, (n_out, n_in, 2), n_in, n_out)
    R = orc.fourier_resample_matrix(n_in, n_out) * n_in       # the plan keeps the 1/N of the whole operator in the last-axis matrix
    assert np.abs(L[..., 0] - R).max() < 1e-5 and np.abs(L[..., 1]).max() < 1e-5
    assert lib.uno_plan_sr_last_modes_fixed(n_in, n_out) == (min(n_in, n_out) - 1) // 2 + 1
