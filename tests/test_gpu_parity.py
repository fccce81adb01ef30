"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in modules) against the golden
fixtures generated from the real reference and against the oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from cases import BLOCK_CASES, POINTWISE_CASES, RESAMPLE_PAIRS, SPECTRAL_CASES
from conftest import BWD_TOL, FWD_TOL, rel_err

pytestmark = pytest.mark.gpu


def _t(a, grad=False):
    t = torch.tensor(np.asarray(a), device="cuda")
    return t.requires_grad_(grad)


def _n(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name", list(SPECTRAL_CASES))
def test_spectral_conv_golden(name, golden, cuda_lib):
    from uno_b200 import integral_operators as ops

    B, Ci, Co, idim, odim, modes = SPECTRAL_CASES[name]
    g = golden("spectral")
    cls = {1: ops.SpectralConv1d_Uno, 2: ops.SpectralConv2d_Uno, 3: ops.SpectralConv3d_Uno}[len(idim)]
    m = cls(Ci, Co, *odim, *modes).cuda()
    nw = 2 ** (len(idim) - 1)
    with torch.no_grad():
        for i in range(nw):
            getattr(m, f"weights{i + 1}").copy_(_t(g[f"{name}.w{i + 1}"]))
    x = _t(g[f"{name}.x"], grad=True)
    y = m(x)
    assert y.shape == g[f"{name}.y"].shape
    assert rel_err(_n(y), g[f"{name}.y"]) < FWD_TOL
    y.backward(_t(g[f"{name}.gy"]))
    assert rel_err(_n(x.grad), g[f"{name}.gx"]) < BWD_TOL
    for i in range(nw):
        assert rel_err(_n(getattr(m, f"weights{i + 1}").grad), g[f"{name}.gw{i + 1}"]) < BWD_TOL
    # inference path (no saved spectrum) gives the same result
    with torch.no_grad():
        y2 = m(x.detach())
    assert torch.equal(y2, y.detach())


@pytest.mark.parametrize("name", list(POINTWISE_CASES))
def test_pointwise_golden(name, golden, cuda_lib):
    from uno_b200 import integral_operators as ops

    B, Ci, Co, idim, odim = POINTWISE_CASES[name]
    g = golden("pointwise")
    cls = {2: ops.pointwise_op_2D, 3: ops.pointwise_op_3D}[len(idim)]
    m = cls(Ci, Co, *odim).cuda()
    with torch.no_grad():
        m.conv.weight.copy_(_t(g[f"{name}.cw"]))
        m.conv.bias.copy_(_t(g[f"{name}.cb"]))
    x = _t(g[f"{name}.x"], grad=True)
    y = m(x)
    assert rel_err(_n(y), g[f"{name}.y"]) < FWD_TOL
    y.backward(_t(g[f"{name}.gy"]))
    assert rel_err(_n(x.grad), g[f"{name}.gx"]) < BWD_TOL
    assert rel_err(_n(m.conv.weight.grad), g[f"{name}.gcw"]) < BWD_TOL
    assert rel_err(_n(m.conv.bias.grad), g[f"{name}.gcb"]) < BWD_TOL


@pytest.mark.parametrize("name", list(BLOCK_CASES))
def test_operator_block_golden(name, golden, cuda_lib):
    from uno_b200 import integral_operators as ops

    B, Ci, Co, idim, odim, modes, norm, nl = BLOCK_CASES[name]
    g = golden("blocks")
    cls = {2: ops.OperatorBlock_2D, 3: ops.OperatorBlock_3D}[len(idim)]
    m = cls(Ci, Co, *odim, *modes, Normalize=norm, Non_Lin=nl).cuda()
    sd = {k[len(name) + 7:]: torch.tensor(g[k]) for k in g.files if k.startswith(f"{name}.param.")}
    m.load_state_dict(sd)
    x = _t(g[f"{name}.x"], grad=True)
    y = m(x, *odim)
    assert rel_err(_n(y), g[f"{name}.y"]) < FWD_TOL
    y.backward(_t(g[f"{name}.gy"]))
    assert rel_err(_n(x.grad), g[f"{name}.gx"]) < BWD_TOL
    scale = max(float(np.abs(g[k]).max()) for k in g.files if k.startswith(f"{name}.grad."))
    for k, p in m.named_parameters():
        ref = g[f"{name}.grad.{k}"]
        # the conv bias in front of an InstanceNorm has an exactly-zero true gradient: compare absolutely
        den = scale if (norm and k == "w.conv.bias") else max(float(np.abs(ref).max()), 1e-3 * scale)
        err = float(np.abs(_n(p.grad) - ref).max()) / den
        assert err < BWD_TOL, (k, err)
    with torch.no_grad():
        y2 = m(x.detach(), *odim)
    assert rel_err(_n(y2), _n(y)) < 1e-6


def test_config1_golden(golden, cuda_lib):
    """BASELINE.json configs[0]: SpectralConv2d 32->32, batch 2, 64x64, 20 modes, seed 0."""
    from uno_b200 import integral_operators as ops

    g = golden("config1")
    torch.manual_seed(0)
    m = ops.SpectralConv2d_Uno(32, 32, 64, 64, 20, 20)
    x = torch.randn(2, 32, 64, 64)
    assert abs(float(m.weights1.detach().abs().double().sum()) - float(g["w1_abs_sum"])) < 1e-6 * float(g["w1_abs_sum"])
    assert abs(float(x.abs().double().sum()) - float(g["x_abs_sum"])) < 1e-6 * float(g["x_abs_sum"])
    m = m.cuda()
    xc = x.cuda().requires_grad_(True)
    y = m(xc)
    assert abs(float(y.double().sum()) - float(g["y_sum"])) < 2e-3          # 365.5583
    assert np.abs(_n(y[0, 0, 0, :3]) - g["y_head"]).max() < 2e-6
    assert rel_err(_n(y[:, ::4, ::4, ::4]), g["y_sub"]) < FWD_TOL
    torch.manual_seed(1)
    gy = torch.randn(2, 32, 64, 64)
    y.backward(gy.cuda())
    assert rel_err(_n(xc.grad[:, ::4, ::4, ::4]), g["gx_sub"]) < BWD_TOL
    assert rel_err(_n(m.weights1.grad[::4, ::4, ::2, ::2]), g["gw1_sub"]) < BWD_TOL
    assert rel_err(_n(m.weights2.grad[::4, ::4, ::2, ::2]), g["gw2_sub"]) < BWD_TOL


@pytest.mark.parametrize("pair", RESAMPLE_PAIRS + ((30, 9), (9, 30), (100, 36)))
def test_resample_matches_aten_bicubic_aa(pair, cuda_lib):
    """pointwise_op_2D with an identity channel mix isolates the fused resample kernel (register-blocked or generic):
    forward against ATen's own anti-aliased bicubic (the call the reference makes, integral_operators.py:240-242),
    backward against its autograd, at the real U-NO level sizes (rectangular: rows n_in -> n_out, columns the reverse
    pair so both band types meet in one launch)."""
    import torch.nn.functional as F

    from uno_b200 import functional as Fn

    n_in, n_out = pair
    C = 3
    torch.manual_seed(n_in * 1000 + n_out)
    x = torch.randn(2, C, n_in, n_out, device="cuda", requires_grad=True)
    w = torch.eye(C, device="cuda").reshape(C, C, 1, 1).requires_grad_(True)
    b = torch.zeros(C, device="cuda", requires_grad=True)
    z = Fn.pointwise_op(x, w, b, (n_out, n_in))
    xr = x.detach().clone().requires_grad_(True)
    zr = F.interpolate(xr, size=(n_out, n_in), mode="bicubic", align_corners=True, antialias=True)
    assert float((z - zr).abs().max() / zr.abs().max()) < FWD_TOL
    g = torch.randn_like(z)
    z.backward(g)
    zr.backward(g)
    assert float((x.grad - xr.grad).abs().max() / xr.grad.abs().max()) < BWD_TOL


@pytest.mark.parametrize("case", [
    # (B, Ci, Co, in_dims, out_dims, modes): shapes that take the TILED mode-contraction kernel (M, N >= 16) with ragged
    # 32 x 32 tiles, a mode count that is not a multiple of the 4 modes a CTA owns, k tails, and a strided outer mode axis
    (32, 24, 40, (20, 18), (16, 14), (5, 3)),
    (17, 48, 33, (12, 12), (12, 12), (6, 6)),
    (20, 16, 19, (10, 9, 8), (8, 9, 8), (3, 2, 3)),
])
def test_spectral_conv_tiled_contraction_vs_oracle(case, cuda_lib):
    from oracle import uno_oracle as orc
    from uno_b200 import functional as Fn

    B, Ci, Co, idim, odim, modes = case
    rng = np.random.default_rng(B * 100 + Ci)
    x = rng.standard_normal((B, Ci) + idim).astype(np.float32)
    nw = 2 ** (len(idim) - 1)
    ws = [((rng.standard_normal((Ci, Co) + modes) + 1j * rng.standard_normal((Ci, Co) + modes)) / np.sqrt(2 * Ci)).astype(np.complex64)
          for _ in range(nw)]
    gy = rng.standard_normal((B, Co) + odim).astype(np.float32)
    y_ref = orc.spectral_conv_fwd(x, ws, odim, modes)
    gx_ref, gw_ref = orc.spectral_conv_bwd(x, ws, odim, modes, gy)
    xt = _t(x, grad=True)
    wt = [_t(w, grad=True) for w in ws]
    y = Fn.spectral_conv(xt, wt, odim, modes)
    assert rel_err(_n(y), y_ref) < FWD_TOL
    y.backward(_t(gy))
    assert rel_err(_n(xt.grad), gx_ref) < BWD_TOL
    for w, r in zip(wt, gw_ref):
        assert rel_err(_n(w.grad), r) < BWD_TOL
