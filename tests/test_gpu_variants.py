"""GPU: every kernel-selection switch of the library (uno_b200/csrc/config.h) against the oracle and against the
alternative kernel it replaces -- tcgen05 leading-axis transform (tc_mid.cuh) and per-mode channel contraction (tc_cmm.cuh),
row-class and 16-loader-warp analysis (tc_kpipe.cuh), 16-warp synthesis epilogue (tc_rowgemm.cuh), big-cluster InstanceNorm,
the band-limited 3-D pointwise mode, empty batches.  All of these ran green on a B200 (profiles/r02_experimental_v44.log)
before they became defaults; both settings of each switch stay covered here so that the driver's `pytest -m gpu` runs them.
"""
import numpy as np
import pytest
import torch

from conftest import BWD_TOL, FWD_TOL, rel_err
from oracle import uno_oracle as orc

pytestmark = [pytest.mark.gpu]


def _with(fn, **switches):
    """run fn() with the library switches set (uno_config_set), restore them afterwards"""
    from uno_b200 import config

    with config.switches(**switches):
        out = fn()
        torch.cuda.synchronize()
        return out


# (B, Ci, Co, in, out, modes): leading-axis sizes with one / two / four column tiles of the transform matrix, ragged row
# tiles (B*C*modes2 not a multiple of 128), ragged last k chunk (2*H % 32 != 0), odd trailing extents; channel counts that
# give the contraction partial row tiles (2*Co % 128 != 0), several column tiles (B > 128) and a ragged k chunk (Ci % 16 != 0)
SHAPES_2D = [
    (8, 8, 8, (64, 64), (64, 64), (20, 20)),
    (4, 16, 24, (48, 40), (130, 36), (12, 9)),
    (2, 32, 64, (481, 64), (240, 40), (18, 18)),
    (3, 5, 7, (33, 31), (45, 29), (7, 5)),
    (32, 32, 32, (32, 32), (32, 32), (6, 6)),
    (4, 192, 192, (16, 16), (16, 16), (6, 6)),
    (130, 9, 70, (24, 24), (20, 20), (5, 4)),
]


@pytest.mark.parametrize("which", ["mid", "cmm", "cmm1", "both", "simt"])
@pytest.mark.parametrize("shape", SHAPES_2D)
def test_spectral2d_tc_core(shape, which, cuda_lib):
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
        return (y.detach().cpu().numpy(), xx.grad.cpu().numpy(), torch.view_as_real(m.weights1.grad).cpu().numpy(),
                torch.view_as_real(m.weights2.grad).cpu().numpy())

    # cmm_tc = 2: every shape, not only tile-filling ones; "cmm1": the one-mode-per-item kernel (tc_cmm.cuh) also where the
    # four-mode kernel (tc_cmm4.cuh) would take the shape
    sw = {"mid_tc": int(which in ("mid", "both")), "cmm_tc": 2 * int(which in ("cmm", "cmm1", "both")), "exp0": 4 if which == "cmm1" else 0}
    y_tc, gx_tc, gw1_tc, gw2_tc = _with(run, **sw)
    ws = [m.weights1.detach().cpu().numpy(), m.weights2.detach().cpu().numpy()]
    y_or = orc.spectral_conv_fwd(x.cpu().numpy(), ws, odim, modes)
    gx_or, gw_or = orc.spectral_conv_bwd(x.cpu().numpy(), ws, odim, modes, gy.cpu().numpy())
    assert rel_err(y_tc, y_or) < FWD_TOL, rel_err(y_tc, y_or)
    assert rel_err(gx_tc, gx_or) < BWD_TOL, rel_err(gx_tc, gx_or)
    assert rel_err(gw1_tc[..., 0] + 1j * gw1_tc[..., 1], gw_or[0]) < BWD_TOL
    assert rel_err(gw2_tc[..., 0] + 1j * gw2_tc[..., 1], gw_or[1]) < BWD_TOL


# last-axis analysis with rows that are not 16-byte aligned and >= 512 rows: pitches = 1, 2, 3 mod 4, ragged 512-row blocks,
# a ragged last chunk, inputs narrower and wider than one chunk ring
SHAPES_ALIGN = [
    (4, 8, 2, (40, 481), (20, 240), (5, 18)),
    (3, 6, 4, (30, 83), (30, 83), (6, 5)),
    (3, 9, 3, (33, 130), (16, 64), (4, 9)),
    (5, 7, 2, (17, 223), (17, 111), (3, 33)),
    (3, 4, 4, (129, 67), (64, 67), (8, 12)),
]


# long contraction (input width > 64): 16-byte aligned rows, 4-byte rows, ragged last chunk, several row tiles per CTA
SHAPES_LW16 = SHAPES_ALIGN + [
    (1, 2, 2, (20, 481), (10, 240), (5, 18)),
    (2, 2, 3, (12, 240), (12, 120), (4, 8)),
    (1, 3, 2, (8, 130), (8, 64), (3, 5)),
    (4, 8, 2, (1200, 100), (16, 16), (4, 6)),
    (8, 16, 2, (240, 240), (120, 120), (6, 18)),
]


@pytest.mark.parametrize("align", [0, 1])
@pytest.mark.parametrize("shape", SHAPES_LW16)
def test_analysis_16_loader_warps(shape, align, cuda_lib):
    """kpipe_lw16: the analysis kernel with 16 loader warps (all three loader paths) against 8."""
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

    a = _with(run, kpipe_lw16=1, kpipe_align=align)
    b = _with(run, kpipe_lw16=0, kpipe_align=0)
    assert rel_err(a[0], b[0]) < FWD_TOL, rel_err(a[0], b[0])
    assert rel_err(a[1], b[1]) < BWD_TOL and rel_err(a[2], b[2]) < BWD_TOL


@pytest.mark.parametrize("shape", SHAPES_ALIGN)
def test_analysis_row_classes(shape, cuda_lib):
    from uno_b200 import integral_operators as ops

    B, Ci, Co, idim, odim, modes = shape
    torch.manual_seed(0)
    m = ops.SpectralConv2d_Uno(Ci, Co, *odim, *modes).cuda()
    x = torch.randn(B, Ci, *idim, device="cuda")
    # non-finite values in the first and the last sample must not leak into the samples between them (the aligned loads of a
    # row start in the tail of the previous row and end in the head of the next one; both ends are masked)
    x[0] = float("inf")
    x[-1] = float("inf")
    gy = torch.randn(B, Co, *odim, device="cuda")

    def run():
        xx = x.clone().requires_grad_(True)
        m.zero_grad(set_to_none=True)
        y = m(xx)
        return y.detach()[1:-1].cpu().numpy()

    y_al = _with(run, kpipe_align=1)
    y_df = _with(run, kpipe_align=0)
    assert np.isfinite(y_al).all()
    assert rel_err(y_al, y_df) < FWD_TOL, rel_err(y_al, y_df)
    x[0] = torch.randn_like(x[0])
    x[-1] = torch.randn_like(x[-1])

    def run2():
        xx = x.clone().requires_grad_(True)
        m.zero_grad(set_to_none=True)
        y = m(xx)
        y.backward(gy)
        return y.detach().cpu().numpy(), xx.grad.cpu().numpy(), torch.view_as_real(m.weights1.grad).cpu().numpy()

    a = _with(run2, kpipe_align=1)
    b = _with(run2, kpipe_align=0)
    ws = [m.weights1.detach().cpu().numpy(), m.weights2.detach().cpu().numpy()]
    y_or = orc.spectral_conv_fwd(x.cpu().numpy(), ws, odim, modes)
    assert rel_err(a[0], y_or) < FWD_TOL, rel_err(a[0], y_or)
    assert rel_err(a[1], b[1]) < BWD_TOL and rel_err(a[2], b[2]) < BWD_TOL


@pytest.mark.parametrize("odim", [(24, 240), (24, 120), (10, 481), (12, 63), (45, 301), (70, 33), (300, 64), (130, 446)])
@pytest.mark.parametrize("norm,nl", [(False, True), (True, True), (False, False)])
def test_synthesis_16_warp_epilogue(odim, norm, nl, cuda_lib):
    """rowgemm_epi16: the synthesis kernel with 16 epilogue warps (store, accumulate, accumulate+GELU to a second
    tensor, in place; one and two column tiles, parity mode, ragged tiles) against the 8-warp configuration
    (both are pinned to the oracle by tests/test_gpu_tc.py)."""
    from uno_b200 import integral_operators as ops

    torch.manual_seed(1)
    blk = ops.OperatorBlock_2D(3, 4, *odim, 5, 9, Normalize=norm, Non_Lin=nl).cuda()
    conv = ops.SpectralConv2d_Uno(3, 4, *odim, 5, 9).cuda()
    x = torch.randn(2, 3, 30, 100, device="cuda")
    gy = torch.randn(2, 4, *odim, device="cuda")

    def run():
        out = []
        for m in (blk, conv):
            xx = x.clone().requires_grad_(True)
            m.zero_grad(set_to_none=True)
            y = m(xx, *odim)
            y.backward(gy)
            with torch.no_grad():
                y_inf = m(x, *odim)
            out += [y.detach().cpu().numpy(), y_inf.cpu().numpy(), xx.grad.cpu().numpy()]
        return out

    a = _with(run, rowgemm_epi16=1)
    b = _with(run, rowgemm_epi16=0)
    for u, v in zip(a, b):
        assert np.array_equal(u, v), rel_err(u, v)     # same arithmetic in the same order: bit-identical


@pytest.mark.parametrize("odim", [(400, 400), (481, 481), (300, 350)])
def test_instance_norm_big_cluster(odim, cuda_lib):
    """norm_big_cluster: planes too large for 8 x 72 KB take the cluster kernels with 200 KB per CTA (forward 4 bytes,
    backward 8 bytes per element: the sizes cover 'both fit', 'forward only' and 'backward at the limit') -- against the
    two-kernel path."""
    from uno_b200 import integral_operators as ops

    torch.manual_seed(2)
    blk = ops.OperatorBlock_2D(2, 3, *odim, 4, 4, Normalize=True).cuda()
    x = torch.randn(2, 2, 64, 64, device="cuda")
    gy = torch.randn(2, 3, *odim, device="cuda")

    def run():
        xx = x.clone().requires_grad_(True)
        blk.zero_grad(set_to_none=True)
        y = blk(xx, *odim)
        y.backward(gy)
        return [y.detach().cpu().numpy(), xx.grad.cpu().numpy(), blk.normalize_layer.weight.grad.cpu().numpy(),
                blk.normalize_layer.bias.grad.cpu().numpy()]

    a = _with(run, norm_big_cluster=1)
    b = _with(run, norm_big_cluster=0)
    assert rel_err(a[0], b[0]) < FWD_TOL, rel_err(a[0], b[0])
    for u, v in zip(a[1:], b[1:]):
        assert rel_err(u, v) < BWD_TOL, rel_err(u, v)


@pytest.mark.parametrize("idim,odim", [((16, 12, 13), (12, 12, 9)), ((12, 10, 9), (16, 14, 13)), ((24, 24, 21), (24, 24, 21))])
def test_pointwise3d_fixed_mode_gpu(idim, odim, cuda_lib):
    """pointwise3d_fixed=1 (opt-in, not the reference) on the CUDA kernels (the CPU suite checks the same orchestration on the host emulation)."""
    from uno_b200 import integral_operators as ops

    torch.manual_seed(0)
    m = ops.pointwise_op_3D(4, 6, *odim).cuda()
    x = torch.randn(2, 4, *idim, device="cuda")

    def run():
        xx = x.clone().requires_grad_(True)
        y = m(xx, *odim)
        y.sum().backward()
        return y.detach().cpu().numpy(), xx.grad.cpu().numpy()

    y, gx = _with(run, pointwise3d_fixed=1)
    cw = m.conv.weight.detach().cpu().numpy().reshape(6, 4)
    y_or = orc.pointwise_op_3d_fixed_fwd(x.cpu().numpy(), cw, m.conv.bias.detach().cpu().numpy(), odim)
    assert rel_err(y, y_or) < FWD_TOL, rel_err(y, y_or)
    R = [orc.fourier_resample_matrix(idim[a], odim[a]) for a in range(3)]
    gt = np.einsum("pd,qe,rf,bcpqr->bcdef", R[0], R[1], R[2], np.ones(y.shape))
    gx_or = np.einsum("oc,bodef->bcdef", cw.astype(np.float64), gt)
    assert rel_err(gx, gx_or) < BWD_TOL, rel_err(gx, gx_or)


def test_empty_batch_like_the_reference(cuda_lib):
    """B = 0: the reference's torch ops return empty tensors and zero parameter gradients; so do the modules and the model."""
    from uno_b200 import integral_operators as ops
    from uno_b200 import models

    torch.manual_seed(0)
    for m, x, args in [
        (ops.SpectralConv2d_Uno(3, 5, 12, 10, 4, 4).cuda(), torch.zeros(0, 3, 16, 16, device="cuda"), (12, 10)),
        (ops.pointwise_op_2D(3, 5, 12, 10).cuda(), torch.zeros(0, 3, 16, 16, device="cuda"), (12, 10)),
        (ops.OperatorBlock_2D(3, 5, 12, 10, 4, 4, Normalize=True).cuda(), torch.zeros(0, 3, 16, 16, device="cuda"), (12, 10)),
        (ops.OperatorBlock_3D(2, 4, 8, 8, 9, 3, 3, 3).cuda(), torch.zeros(0, 2, 8, 8, 9, device="cuda"), (8, 8, 9)),
    ]:
        x.requires_grad_(True)
        y = m(x, *args)
        assert y.shape == (0, y.shape[1]) + tuple(args) and y.is_cuda
        y.sum().backward()
        assert x.grad.shape == x.shape
        for p in m.parameters():
            assert p.grad is not None and float(p.grad.abs().max() if p.grad.numel() else 0) == 0.0
    net = models.UNO_9(3, 8, pad=5).cuda()
    out = net(torch.zeros(0, 85, 85, 1, device="cuda"))
    assert out.shape == (0, 85, 85, 1)
    out.sum().backward()
    assert all(p.grad is not None for p in net.parameters())


SHAPES_3D = [
    (2, 4, 6, (16, 16, 13), (12, 12, 13), (5, 5, 4)),
    (1, 8, 16, (24, 20, 21), (24, 20, 21), (8, 6, 5)),
]


@pytest.mark.parametrize("shape", SHAPES_3D)
def test_spectral3d_tc_core(shape, cuda_lib):
    from uno_b200 import integral_operators as ops

    B, Ci, Co, idim, odim, modes = shape
    torch.manual_seed(0)
    m = ops.SpectralConv3d_Uno(Ci, Co, *odim, *modes).cuda()
    x = torch.randn(B, Ci, *idim, device="cuda")
    gy = torch.randn(B, Co, *odim, device="cuda")

    def run():
        xx = x.clone().requires_grad_(True)
        m.zero_grad(set_to_none=True)
        y = m(xx)
        y.backward(gy)
        return y.detach().cpu().numpy(), xx.grad.cpu().numpy(), [torch.view_as_real(w.grad).cpu().numpy() for w in (m.weights1, m.weights2, m.weights3, m.weights4)]

    y_tc, gx_tc, gw_tc = _with(run, mid_tc=1, cmm_tc=2)
    y_si, gx_si, gw_si = _with(run, mid_tc=0, cmm_tc=0)
    # the default kernels are pinned to the oracle by tests/test_gpu_parity.py; here the two device paths are compared
    assert rel_err(y_tc, y_si) < FWD_TOL, rel_err(y_tc, y_si)
    assert rel_err(gx_tc, gx_si) < BWD_TOL, rel_err(gx_tc, gx_si)
    for a, b in zip(gw_tc, gw_si):
        assert rel_err(a, b) < BWD_TOL, rel_err(a, b)


@pytest.mark.parametrize("norm,nl,odim", [(False, True, (24, 20)), (True, True, (16, 16)), (False, False, (12, 15)), (True, False, (10, 9)),
                                         (False, True, (9, 7, 5)), (True, True, (8, 8, 6))])
def test_fanout_matches_autograd_accumulation(norm, nl, odim, cuda_lib):
    """A block output with two consumers, one of them a skip concatenation: with fanout=2 the block's backward receives the two
    upstream gradients separately -- one contiguous, one a channel slice of the concatenation's gradient (batch-strided) -- and
    adds them in its first kernel.  Must equal the plain graph, where autograd slices, copies and adds."""
    from uno_b200 import integral_operators as ops

    nd = len(odim)
    torch.manual_seed(3)
    Blk = ops.OperatorBlock_2D if nd == 2 else ops.OperatorBlock_3D
    modes = (3,) * nd
    blk = Blk(3, 4, *odim, *modes, Normalize=norm, Non_Lin=nl).cuda()
    nxt = Blk(4, 2, *odim, *modes).cuda()
    idim = tuple(n + 2 for n in odim)
    x = torch.randn(2, 3, *idim, device="cuda")
    other = torch.randn(2, 5, *odim, device="cuda")
    wcat = torch.randn(1, 9, *([1] * nd), device="cuda")

    def run(fan):
        xx = x.clone().requires_grad_(True)
        for m in (blk, nxt):
            m.zero_grad(set_to_none=True)
        if fan:
            y, y_skip = blk(xx, *odim, fanout=2)
            assert y.data_ptr() == y_skip.data_ptr()
        else:
            y = y_skip = blk(xx, *odim)
        z = nxt(y, *odim)
        c = torch.cat([other, y_skip], dim=1)
        loss = (z ** 2).sum() + (c * wcat).sum() + (c ** 2).sum() * 0.1
        loss.backward()
        return [float(loss.detach()), xx.grad.cpu().numpy()] + [
            (torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).cpu().numpy() for p in blk.parameters()]

    a, b = run(True), run(False)
    assert abs(a[0] - b[0]) < 1e-6 * abs(b[0])
    for u, v in zip(a[1:], b[1:]):
        assert np.abs(u - v).max() <= 2e-5 * max(float(np.abs(v).max()), 1e-6), rel_err(u, v)


def test_lift_fanout_matches_autograd_accumulation(cuda_lib):
    """The lifted input feeds the first block and the projection: lift(fanout=2) adds the two gradients inside lift_bwd."""
    from uno_b200 import functional as Fn

    torch.manual_seed(0)
    a = torch.randn(2, 20, 18, 1, device="cuda")
    grid = torch.randn(20, 18, 2, device="cuda")
    ps = [torch.randn(*s, device="cuda") * 0.3 for s in ((16, 3), (16,), (32, 16), (32,))]
    w1 = torch.randn(1, 32, 1, 1, device="cuda")
    w2 = torch.randn(1, 32, 1, 1, device="cuda")

    def run(fan):
        qs = [p.clone().requires_grad_(True) for p in ps]
        if fan:
            h, h2 = Fn.lift(a, grid, *qs, (0, 1), (3, 2), fanout=2)
        else:
            h = h2 = Fn.lift(a, grid, *qs, (0, 1), (3, 2))
        ((h * w1).sum() + ((h2 * w2) ** 2).sum()).backward()
        return [q.grad.cpu().numpy() for q in qs]

    for u, v in zip(run(True), run(False)):
        assert np.abs(u - v).max() <= 2e-5 * max(float(np.abs(v).max()), 1e-6)


@pytest.mark.parametrize("case", ["2d_norm", "2d_plain", "3d"])
def test_branch_overlap_matches_serial(case, cuda_lib):
    """overlap: the pointwise branch of a block on the library's side stream, the spectral branch on the caller's, joined before
    the fused epilogue (forward) / before the accumulation onto gx (backward).  Same kernels, same results as one stream."""
    from uno_b200 import integral_operators as ops

    torch.manual_seed(7)
    if case == "3d":
        blk = ops.OperatorBlock_3D(4, 6, 12, 12, 9, 4, 4, 3, Normalize=True).cuda()
        x = torch.randn(2, 4, 16, 16, 9, device="cuda")
        odim = (12, 12, 9)
    else:
        blk = ops.OperatorBlock_2D(8, 16, 60, 52, 9, 7, Normalize=(case == "2d_norm")).cuda()
        x = torch.randn(3, 8, 120, 101, device="cuda")
        odim = (60, 52)
    gy = torch.randn(x.shape[0], blk.conv.out_channels, *odim, device="cuda")

    def run():
        out = []
        for _ in range(3):          # repeated: a missing join shows up as a race between iterations
            xx = x.clone().requires_grad_(True)
            blk.zero_grad(set_to_none=True)
            y = blk(xx, *odim)
            y.backward(gy)
            out = [y.detach().cpu().numpy(), xx.grad.cpu().numpy()] + [
                (torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).cpu().numpy() for p in blk.parameters()]
        return out

    a, b = _with(run, overlap=1), _with(run, overlap=0)
    for u, v in zip(a, b):
        assert np.abs(u - v).max() <= 1e-5 * max(float(np.abs(v).max()), 1e-6), rel_err(u, v)


def test_first_call_of_a_new_shape_is_already_correct(cuda_lib):
    """Constant tables (plan matrices, tensor-core operand images) are built at the first call that needs them and read, in that
    same call, by kernels on the library's non-blocking side streams.  The upload must be complete for those streams too: with a
    plain cudaMemcpy from pageable memory the first backward of a new shape intermittently used half-written images once the
    kernels were resident.  Every iteration uses shapes no earlier test has used; the first call must equal the second."""
    from uno_b200 import integral_operators as ops

    torch.manual_seed(11)
    for i in range(6):
        idim, odim, modes = (21 + i, 19 + i, 11 + i), (17 + i, 23 - i, 13 + i), (4, 5, 3)
        blk = ops.OperatorBlock_3D(5, 7, *odim, *modes, Normalize=bool(i & 1)).cuda()
        x = torch.randn(2, 5, *idim, device="cuda")
        gy = torch.randn(2, 7, *odim, device="cuda")

        def run():
            xx = x.clone().requires_grad_(True)
            blk.zero_grad(set_to_none=True)
            y = blk(xx, *odim)
            y.backward(gy)
            return [y.detach().cpu().numpy(), xx.grad.cpu().numpy()] + [
                (torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).cpu().numpy() for p in blk.parameters()]

        first, second = run(), run()
        for u, v in zip(first, second):
            assert np.abs(u - v).max() <= 1e-5 * max(float(np.abs(v).max()), 1e-6), (i, rel_err(u, v))


# output widths whose row pitch is not a multiple of 8 floats, with at least 128 * nclass output rows so that the row-class tiles are
# taken: odd pitches (8 classes; 481 and 83 are the Darcy and NS-3D widths), pitches = 2, 6 mod 8 (4 classes), 4 mod 8 (2 classes)
SECTOR_WIDTHS = [83, 481, 301, 45, 62, 446, 54, 52, 124]


@pytest.mark.parametrize("width", SECTOR_WIDTHS)
@pytest.mark.parametrize("norm,nl", [(False, True), (False, False)])
def test_synthesis_sector_row_classes(width, norm, nl, cuda_lib):
    """tc_rowgemm.cuh row classes: a tile holds rows of one residue class and loads a column-shifted twiddle image so that every lane
    quad covers one whole 32-byte sector (shift up to 7 columns: the first threads of the first column tile own columns left of
    the row start, which must stay untouched).  Block forward (accumulate + GELU to a second tensor / plain accumulate) and
    backward (accumulate onto the pointwise gradient) against the oracle, and against the two-class parity tiles."""
    from uno_b200 import integral_operators as ops

    torch.manual_seed(5)
    odim = (44, width)
    modes = (5, min(9, width // 2))
    blk = ops.OperatorBlock_2D(3, 8, *odim, *modes, Normalize=norm, Non_Lin=nl).cuda()
    x = torch.randn(4, 3, 30, 40, device="cuda")            # 4 * 8 * 44 = 1408 output rows >= 128 * 8
    gy = torch.randn(4, 8, *odim, device="cuda")

    def run():
        xx = x.clone().requires_grad_(True)
        blk.zero_grad(set_to_none=True)
        y = blk(xx, *odim)
        y.backward(gy)
        return [y.detach().cpu().numpy(), xx.grad.cpu().numpy()]

    a = _with(run, rowgemm_parity=1)
    b = _with(run, rowgemm_parity=2)                          # 2: two-class parity tiles only (the behaviour before the row classes)
    assert rel_err(a[0], b[0]) < 1e-6 and rel_err(a[1], b[1]) < 1e-6      # same products, same order per output element
    p = {k: v.detach().cpu().numpy() for k, v in blk.state_dict().items()}
    y_or = orc.operator_block_fwd(x.cpu().numpy(), [p["conv.weights1"], p["conv.weights2"]], p["w.conv.weight"], p["w.conv.bias"],
                                  odim, modes, norm=None, non_lin=nl)
    assert rel_err(a[0], y_or) < FWD_TOL
