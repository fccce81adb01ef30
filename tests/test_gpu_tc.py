"""GPU: the tcgen05 (3xTF32) kernels against the SIMT fp32 kernels and the oracle, shape by shape."""
import numpy as np
import pytest
import torch

from conftest import BWD_TOL, FWD_TOL, rel_err
from oracle import uno_oracle as orc

pytestmark = pytest.mark.gpu


def _run(fn, disable_tc):
    from uno_b200 import config

    with config.switches(tc=0 if disable_tc else 1):
        out = fn()
        torch.cuda.synchronize()
        return out


# (B, Ci, Co, in, out, modes): output widths cover one/two N tiles, padded tiles, odd leading dims,
# K = 2*modes2 with and without 16-byte aligned rows, and ragged row-tile tails
SHAPES = [
    (2, 3, 4, (40, 64), (30, 240), (6, 18)),
    (1, 2, 3, (64, 64), (17, 481), (5, 18)),
    (2, 2, 2, (32, 32), (50, 120), (8, 8)),
    (1, 3, 2, (32, 32), (33, 31), (4, 5)),
    (2, 2, 3, (32, 48), (9, 16), (3, 7)),
    (1, 2, 2, (64, 64), (300, 64), (20, 22)),
    (1, 1, 2, (96, 96), (130, 446), (9, 32)),
    # odd row pitch -> row-parity tiles with the column-shifted twiddle image: several 256-row blocks with a ragged
    # last one, two column tiles, and the float2 path of both parities
    (3, 2, 3, (40, 40), (45, 301), (6, 20)),
    (2, 2, 5, (32, 32), (131, 77), (7, 9)),
    # long contraction (K = input width > 64): K-pipelined kernel, 16-byte aligned and unaligned rows,
    # ragged last chunk, several row tiles per CTA
    (1, 2, 2, (20, 481), (10, 240), (5, 18)),
    (2, 2, 3, (12, 240), (12, 120), (4, 8)),
    (1, 2, 2, (9, 223), (9, 111), (3, 33)),
    (1, 3, 2, (8, 130), (8, 64), (3, 5)),
    (4, 8, 2, (1200, 100), (16, 16), (4, 6)),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_rowgemm_tc_matches_simt_and_oracle(shape, cuda_lib):
    from uno_b200 import integral_operators as ops

    B, Ci, Co, idim, odim, modes = shape
    torch.manual_seed(0)
    m = ops.SpectralConv2d_Uno(Ci, Co, *odim, *modes).cuda()
    x = torch.randn(B, Ci, *idim, device="cuda")
    gy = torch.randn(B, Co, *odim, device="cuda")

    def run():
        xx = x.clone().requires_grad_(True)
        m.zero_grad(set_to_none=True)
        y = m(xx)
        y.backward(gy)
        return y.detach().cpu().numpy(), xx.grad.cpu().numpy(), torch.view_as_real(m.weights1.grad).cpu().numpy()

    y_tc, gx_tc, gw_tc = _run(run, disable_tc=False)
    y_si, gx_si, gw_si = _run(run, disable_tc=True)
    ws = [m.weights1.detach().cpu().numpy(), m.weights2.detach().cpu().numpy()]
    y_or = orc.spectral_conv_fwd(x.cpu().numpy(), ws, odim, modes)
    gx_or, gw_or = orc.spectral_conv_bwd(x.cpu().numpy(), ws, odim, modes, gy.cpu().numpy())
    assert rel_err(y_si, y_or) < FWD_TOL
    assert rel_err(y_tc, y_or) < FWD_TOL, rel_err(y_tc, y_or)
    assert rel_err(gx_tc, gx_or) < BWD_TOL, rel_err(gx_tc, gx_or)
    assert rel_err(gx_si, gx_or) < BWD_TOL
    assert rel_err(gw_tc[..., 0] + 1j * gw_tc[..., 1], gw_or[0]) < BWD_TOL


def test_block_epilogues_tc(cuda_lib):
    """The fused accumulate / GELU epilogues of the tensor-core synthesis kernel."""
    from oracle import uno_torch_port as port
    from uno_b200 import integral_operators as ops

    for norm, nl, odim in [(False, True, (24, 240)), (True, True, (24, 120)), (False, False, (10, 481)), (False, True, (12, 63)),
                          (False, True, (45, 301)), (True, True, (70, 33))]:
        torch.manual_seed(1)
        blk = ops.OperatorBlock_2D(3, 4, *odim, 5, 9, Normalize=norm, Non_Lin=nl).cuda()
        x = torch.randn(2, 3, 30, 100, device="cuda", requires_grad=True)
        y = blk(x, *odim)
        gy = torch.randn_like(y)
        y.backward(gy)
        with torch.no_grad():
            y_inf = blk(x.detach(), *odim)
        xr = x.detach().cpu().double().requires_grad_(True)
        ws = [blk.conv.weights1.detach().cpu().to(torch.cdouble), blk.conv.weights2.detach().cpu().to(torch.cdouble)]
        ga = blk.normalize_layer.weight.detach().cpu().double() if norm else None
        be = blk.normalize_layer.bias.detach().cpu().double() if norm else None
        yr = port.operator_block(xr, ws, blk.w.conv.weight.detach().cpu().double(), blk.w.conv.bias.detach().cpu().double(), odim, (5, 9), ga, be, nl)
        yr.backward(gy.cpu().double())
        assert rel_err(y.detach().cpu().numpy(), yr.detach().numpy()) < FWD_TOL
        assert rel_err(y_inf.cpu().numpy(), yr.detach().numpy()) < FWD_TOL
        assert rel_err(x.grad.cpu().numpy(), xr.grad.numpy()) < BWD_TOL


# (B, Ci, Co, H, W): identity-size pointwise op = pure 1x1 channel mix (tensor-core path needs H*W % 4 == 0, >= 128)
CONV_SHAPES = [(2, 32, 64, 24, 24), (1, 48, 96, 16, 20), (1, 128, 128, 20, 20), (2, 20, 24, 12, 12), (1, 64, 32, 120, 120), (3, 8, 8, 16, 8)]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_channel_mix_tc_matches_simt_and_fp64(shape, cuda_lib):
    from uno_b200 import integral_operators as ops

    B, Ci, Co, H, W = shape
    torch.manual_seed(0)
    m = ops.pointwise_op_2D(Ci, Co, H, W).cuda()
    x = torch.randn(B, Ci, H, W, device="cuda")
    gy = torch.randn(B, Co, H, W, device="cuda")

    def run():
        xx = x.clone().requires_grad_(True)
        m.zero_grad(set_to_none=True)
        y = m(xx)
        y.backward(gy)
        return y.detach().cpu().numpy(), xx.grad.cpu().numpy(), m.conv.weight.grad.cpu().numpy()

    y_tc, gx_tc, gw_tc = _run(run, disable_tc=False)
    y_si, gx_si, gw_si = _run(run, disable_tc=True)
    w64 = m.conv.weight.detach().cpu().double().reshape(Co, Ci)
    y_ref = torch.einsum("oc,bchw->bohw", w64, x.cpu().double()) + m.conv.bias.detach().cpu().double().view(1, -1, 1, 1)
    gx_ref = torch.einsum("oc,bohw->bchw", w64, gy.cpu().double())
    assert rel_err(y_si, y_ref.numpy()) < FWD_TOL
    assert rel_err(y_tc, y_ref.numpy()) < FWD_TOL, rel_err(y_tc, y_ref.numpy())
    assert rel_err(gx_tc, gx_ref.numpy()) < BWD_TOL, rel_err(gx_tc, gx_ref.numpy())
    gw_ref = torch.einsum("bohw,bchw->oc", gy.cpu().double(), x.cpu().double()).reshape(Co, Ci, 1, 1)
    assert rel_err(gw_si, gw_ref.numpy()) < BWD_TOL
    assert rel_err(gw_tc, gw_ref.numpy()) < BWD_TOL, rel_err(gw_tc, gw_ref.numpy())
