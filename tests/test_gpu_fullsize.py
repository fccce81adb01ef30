"""GPU: parity at the FULL widths and grids of BASELINE.json configs 2-4 and the corners of config 5 -- the CUDA path against the
reference itself on the CPU (oracle/_ref: the reference's own model files, staged by build(); the functional port when the
staged copy is absent), same weights, element-wise on the output and on EVERY parameter gradient.

    config 2  UNO_9(3,32,pad=12), 421 x 421, batch 2                 darcy_flow_main.py:95, train_darcy.py:50-54
    config 3  UNO(14,32), 64 x 64, batch 4, 3-step rollout + BPTT    ns_train_2d.py:52-67
    config 4  Uno3D_T10(6,8,pad=3), 64 x 64 x 64 (T -> 83), batch 1  ns_uno3d_main.py:103, ns_train_3d.py:51-65
    config 5  SpectralConv2d corners S=512 C=128 m=32 and S=64 C=32 m=32 against the fp32 torch restatement

Tolerances (tests/conftest.py): a whole model multiplies the per-layer bounds by its depth -- output max|d| <= 1e-4 max|y|
(the model-level bound of SURVEY 8(c)), gradients max|d| <= 4 * BWD_TOL * max|g| per parameter (8 * for the rollout, whose
backward runs through 3 model calls).  The reference's own fp32-vs-fp64 noise at these sizes is ~1e-5.
"""
import numpy as np
import pytest
import torch

import bench
from conftest import BWD_TOL, FWD_TOL, rel_err

pytestmark = pytest.mark.gpu


def _grads(model):
    return [(k, (torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).detach().cpu()) for k, p in model.named_parameters()]


def _compare(ours, ref, y_o, y_r, gtol, capsys, tag):
    ey = rel_err(y_o.detach().cpu().numpy(), y_r.detach().numpy())
    worst = ("", 0.0)
    # Each gradient is compared relative to its own largest magnitude, floored at 1e-4 of the largest gradient of the model: the
    # conv bias of a block WITH InstanceNorm has an exactly-zero gradient (the norm removes the per-plane mean); the reference
    # returns the fp32 residue of that sum (~1e-7 of the gradient scale), this library returns 0.
    gmax = max(float(b.abs().max()) for _, b in _grads(ref))
    gfloor = 1e-4 * gmax
    for (k, a), (k2, b) in zip(_grads(ours), _grads(ref)):
        assert k == k2 and a.shape == b.shape
        if k.endswith(".w.conv.bias") and float(a.abs().max()) == 0.0:
            # mathematically zero (bias in front of an InstanceNorm): the reference's value is pure rounding residue
            assert float(b.abs().max()) < 1e-5 * gmax, (k, float(b.abs().max()), gmax)
            continue
        e = float((a - b).abs().max()) / max(float(b.abs().max()), gfloor, 1e-12)
        if e > worst[1]:
            worst = (k, e)
        assert e < gtol, (k, e)
    with capsys.disabled():
        print(f"\n[fullsize {tag}] output max rel err {ey:.2e}; worst parameter gradient {worst[0]} {worst[1]:.2e} (bound {gtol:.1e})")
    assert ey < 1e-4, ey


def _pair(workload):
    ref, LpLoss, kind = bench.reference_model(workload, "cpu")
    ours = bench.build_model(workload, device="cuda")
    missing = ours.load_state_dict(ref.state_dict(), strict=True)      # same schema, same weights
    assert not missing.missing_keys and not missing.unexpected_keys
    return ref, ours, LpLoss, kind


@pytest.mark.parametrize("workload,B", [("darcy", 2), ("ns3d", 1)])
def test_full_width_model_matches_reference(workload, B, cuda_lib, capsys):
    from uno_b200.losses import LpLoss as OurLoss

    _, _, _, xshape, tshape, _, _ = bench.WORKLOADS[workload]
    ref, ours, RefLoss, kind = _pair(workload)
    torch.manual_seed(11)
    x, t = torch.randn(B, *xshape), torch.randn(B, *tshape)
    y_r = ref(x).reshape(B, *tshape)
    l_r = RefLoss(size_average=False)(y_r.reshape(B, -1), t.reshape(B, -1))
    l_r.backward()
    y_o = ours(x.cuda()).reshape(B, *tshape)
    l_o = OurLoss(size_average=False)(y_o.reshape(B, -1), t.cuda().reshape(B, -1))
    l_o.backward()
    assert abs(float(l_o.detach()) - float(l_r.detach())) < 1e-4 * abs(float(l_r.detach()))
    _compare(ours, ref, y_o, y_r, 4 * BWD_TOL, capsys, f"{workload} B={B} vs {kind}")


def test_full_width_rollout_matches_reference(cuda_lib, capsys):
    """UNO(14,32), batch 4, 3 autoregressive steps with the summed loss and one backward (ns_train_2d.py:52-67)."""
    from uno_b200.losses import LpLoss as OurLoss

    B, T = 4, 3
    ref, ours, RefLoss, kind = _pair("ns2d_ar")
    torch.manual_seed(12)
    x, t = torch.randn(B, 64, 64, 10), torch.randn(B, 64, 64, T)
    l_r = bench.make_step(ref, RefLoss(size_average=False), B, (64, 64, T), ar_steps=T)(x, t)
    l_o = bench.make_step(ours, OurLoss(size_average=False), B, (64, 64, T), ar_steps=T)(x.cuda(), t.cuda())
    assert abs(float(l_o.detach()) - float(l_r.detach())) < 1e-4 * abs(float(l_r.detach()))
    with torch.no_grad():
        y_r, y_o = ref(x), ours(x.cuda())
    _compare(ours, ref, y_o, y_r, 8 * BWD_TOL, capsys, f"ns2d_ar B={B} T={T} vs {kind}")


def _stock_layer(x, w1, w2, d1, d2, m1, m2):
    """integral_operators.py:181-207 restated with the torch calls the reference makes (CPU fp32)."""
    xh = torch.fft.rfft2(x, norm="forward")
    yh = torch.zeros(x.shape[0], w1.shape[1], d1, d2 // 2 + 1, dtype=torch.cfloat)
    yh[:, :, :m1, :m2] = torch.einsum("bixy,ioxy->boxy", xh[:, :, :m1, :m2], w1)
    yh[:, :, -m1:, :m2] = torch.einsum("bixy,ioxy->boxy", xh[:, :, -m1:, :m2], w2)
    return torch.fft.irfft2(yh, s=(d1, d2), norm="forward")


@pytest.mark.parametrize("S,C,m,B", [(512, 128, 32, 1), (64, 32, 32, 16), (256, 64, 20, 2), (128, 128, 12, 4)])
def test_sweep_corner_matches_reference_layer(S, C, m, B, cuda_lib, capsys):
    """BASELINE configs[4]: SpectralConv2d_Uno(C,C,S,S,m,m) at the corners of the kernel sweep, output, dx and both dW."""
    from uno_b200 import integral_operators as ops

    try:
        from oracle import ref_loader

        Ref = ref_loader.load("integral_operators").SpectralConv2d_Uno
    except ImportError:
        Ref = None
    torch.manual_seed(0)
    layer = ops.SpectralConv2d_Uno(C, C, S, S, m, m).cuda()
    w1, w2 = layer.weights1.detach().cpu().requires_grad_(True), layer.weights2.detach().cpu().requires_grad_(True)
    torch.manual_seed(1)
    x = torch.randn(B, C, S, S)
    g = torch.randn(B, C, S, S)
    xr = x.clone().requires_grad_(True)
    if Ref is not None:
        r = Ref(C, C, S, S, m, m)
        with torch.no_grad():
            r.weights1.copy_(w1)
            r.weights2.copy_(w2)
        y_r = r(xr)
        y_r.backward(g)
        gw_r = [r.weights1.grad, r.weights2.grad]
    else:
        y_r = _stock_layer(xr, w1, w2, S, S, m, m)
        y_r.backward(g)
        gw_r = [w1.grad, w2.grad]
    xo = x.cuda().requires_grad_(True)
    y_o = layer(xo)
    y_o.backward(g.cuda())
    ef = rel_err(y_o.detach().cpu().numpy(), y_r.detach().numpy())
    ex = rel_err(xo.grad.cpu().numpy(), xr.grad.numpy())
    ew = max(rel_err(torch.view_as_real(a.grad).cpu().numpy(), torch.view_as_real(b).numpy()) for a, b in zip((layer.weights1, layer.weights2), gw_r))
    with capsys.disabled():
        print(f"\n[sweep corner S={S} C={C} m={m} B={B}] fwd {ef:.2e} dx {ex:.2e} dW {ew:.2e}")
    assert ef < FWD_TOL and ex < BWD_TOL and ew < BWD_TOL, (ef, ex, ew)
