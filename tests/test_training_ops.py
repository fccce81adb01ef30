"""Adam (Adam.py) and LpLoss (utilities3.py) -- SURVEY.md section 8(f) row 3 -- against fixtures generated from the real
reference (tests/golden/training.npz): the oracle restatement and the C-ABI orchestration on the host emulation (CPU),
and the CUDA kernels through the public classes (GPU)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err
from oracle import uno_torch_port as port
from uno_b200 import _capi

sys.path.insert(0, os.path.join(ROOT, "tests", "hostemu"))

ADAM_CASES = {"plain": dict(lr=1e-2), "wd": dict(lr=3e-3, weight_decay=1e-2, betas=(0.8, 0.95), eps=1e-6), "amsgrad": dict(lr=1e-2, amsgrad=True)}
LOSS_MODES = {"none": 0, "sum": 1, "mean": 2}
STEP_TOL = 2e-6        # fp32 round-off of a handful of fused vs separate roundings, relative to the largest weight


def _full(kw):
    return dict(dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False), **kw)


@pytest.mark.parametrize("tag", list(ADAM_CASES))
def test_adam_oracle_and_hostemu_match_reference(tag, golden):
    import emu

    g = golden("training")
    kw = _full(ADAM_CASES[tag])
    n = int(g[f"{tag}.n"])
    # --- oracle restatement (torch, out of place)
    ps = [torch.tensor(g[f"{tag}.p0.{i}"]) for i in range(n)]
    sts = [dict(exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p), max_exp_avg_sq=torch.zeros_like(p)) for p in ps]
    # --- C ABI on the host emulation (numpy, in place)
    P = [np.array(g[f"{tag}.p0.{i}"]) for i in range(n)]
    M = [np.zeros_like(p) for p in P]
    V = [np.zeros_like(p) for p in P]
    VM = [np.zeros_like(p) for p in P] if kw["amsgrad"] else None
    for step in range(5):
        grads = [g[f"{tag}.g{step}.{i}"] for i in range(n)]
        ps = [port.adam_step(p, torch.tensor(gr), st, step + 1, **kw) for p, gr, st in zip(ps, grads, sts)]
        emu.adam_step(P, [np.ascontiguousarray(x) for x in grads], M, V, VM, step + 1, kw["lr"], kw["betas"], kw["eps"], kw["weight_decay"],
                      kw["amsgrad"])
        for i in range(n):
            ref = g[f"{tag}.p{step + 1}.{i}"]
            assert rel_err(ps[i].numpy(), ref) < STEP_TOL, (step, i)
            assert rel_err(P[i], ref) < STEP_TOL, (step, i)
    for i in range(n):
        assert rel_err(M[i], g[f"{tag}.m.{i}"]) < STEP_TOL
        assert rel_err(V[i], g[f"{tag}.v.{i}"]) < STEP_TOL
        if np.iscomplexobj(V[i]):
            assert np.all(V[i].imag == 0)       # |g|^2 semantics of Adam.py:41: a real second moment in a complex tensor


@pytest.mark.parametrize("mode", list(LOSS_MODES))
def test_lp_loss_oracle_and_hostemu_match_reference(mode, golden):
    import emu

    g = golden("training")
    x, y, gl = g["loss.x"], g["loss.y"], g[f"loss.{mode}.gl"]
    kw = dict(none=dict(reduction=False), sum=dict(size_average=False), mean=dict(size_average=True))[mode]
    xt = torch.tensor(x, requires_grad=True)
    lo = port.LpLoss(**kw)(xt, torch.tensor(y))
    (gx,) = torch.autograd.grad(lo, xt, torch.tensor(gl))
    assert rel_err(lo.detach().numpy(), g[f"loss.{mode}"]) < 1e-6
    assert rel_err(gx.numpy(), g[f"loss.{mode}.gx"]) < 1e-6
    loss, gxe = emu.lp_loss(x, y, LOSS_MODES[mode], gl)
    assert rel_err(loss.reshape(np.shape(g[f"loss.{mode}"])), g[f"loss.{mode}"]) < 1e-6
    assert rel_err(gxe.reshape(x.shape), g[f"loss.{mode}.gx"]) < 2e-6


def test_lp_loss_exactly_matched_sample_gives_zero_gradient():
    """x[b] == y[b]: torch's norm backward masks the zero norm and returns a zero gradient for that sample (the reference's
    LpLoss is torch.norm); a 0 * inf = NaN there would poison every parameter gradient."""
    import emu

    rng = np.random.default_rng(0)
    y = rng.standard_normal((3, 50)).astype(np.float32)
    x = rng.standard_normal((3, 50)).astype(np.float32)
    x[1] = y[1]
    xt = torch.tensor(x, requires_grad=True)
    lo = port.LpLoss(size_average=False)(xt, torch.tensor(y))
    lo.backward()
    assert torch.isfinite(xt.grad).all() and float(xt.grad[1].abs().max()) == 0.0
    loss, gx = emu.lp_loss(x, y, 1, np.ones(1, np.float32))
    assert np.isfinite(gx).all() and np.abs(gx.reshape(3, 50)[1]).max() == 0.0
    assert rel_err(gx.reshape(3, 50), xt.grad.numpy()) < 2e-6


def test_adam_argument_errors():
    import emu

    L = emu.lib()
    p = np.zeros(4, np.float32)
    t = (_capi.AdamTensor * 1)()
    t[0].param = t[0].grad = t[0].exp_avg = t[0].exp_avg_sq = p.ctypes.data
    t[0].numel = 4
    for bad in (dict(lr=-1.0), dict(beta1=1.0), dict(beta2=-0.1), dict(eps=-1.0), dict(weight_decay=-1.0), dict(step=0)):
        kw = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, amsgrad=0, step=1)
        kw.update(bad)
        h = _capi.AdamHyper(kw["lr"], kw["beta1"], kw["beta2"], kw["eps"], kw["weight_decay"], kw["amsgrad"], kw["step"])
        assert L.uno_adam_step(t, 1, C.byref(h), None) == 1, bad
    t[0].is_complex = 1
    h = _capi.AdamHyper(1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1)           # amsgrad on a complex tensor: upstream raises too
    assert L.uno_adam_step(t, 1, C.byref(h), None) == 1
    from uno_b200.optim import Adam

    with pytest.raises(ValueError, match="Invalid learning rate"):
        Adam([torch.nn.Parameter(torch.zeros(2))], lr=-1)
    with pytest.raises(ValueError, match="Invalid beta parameter at index 1"):
        Adam([torch.nn.Parameter(torch.zeros(2))], betas=(0.9, 1.0))


# ------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(ADAM_CASES))
def test_adam_cuda_matches_reference(tag, golden, cuda_lib):
    from uno_b200.optim import Adam

    g = golden("training")
    n = int(g[f"{tag}.n"])
    params = [torch.nn.Parameter(torch.tensor(g[f"{tag}.p0.{i}"], device="cuda")) for i in range(n)]
    opt = Adam(params, **ADAM_CASES[tag])
    for step in range(5):
        for i, p in enumerate(params):
            p.grad = torch.tensor(g[f"{tag}.g{step}.{i}"], device="cuda")
        opt.step()
        for i, p in enumerate(params):
            assert rel_err(p.detach().cpu().numpy(), g[f"{tag}.p{step + 1}.{i}"]) < STEP_TOL, (step, i)
    for i, p in enumerate(params):
        st = opt.state[p]
        assert st["step"] == 5 and st["exp_avg"].dtype == p.dtype and st["exp_avg_sq"].dtype == p.dtype
        assert rel_err(st["exp_avg"].cpu().numpy(), g[f"{tag}.m.{i}"]) < STEP_TOL
        assert rel_err(st["exp_avg_sq"].cpu().numpy(), g[f"{tag}.v.{i}"]) < STEP_TOL


@pytest.mark.gpu
def test_adam_cuda_many_tensors_and_scheduler(cuda_lib):
    """More tensors than one launch holds (24), ragged sizes, StepLR on top, a parameter without gradient."""
    from uno_b200.optim import Adam

    torch.manual_seed(3)
    shapes = [(1 + 37 * i,) for i in range(40)] + [(9000,), (3, 5, 7)]
    params = [torch.nn.Parameter(torch.randn(*s, dtype=torch.cfloat if i % 3 == 0 else torch.float)) for i, s in enumerate(shapes)]
    ref = [p.detach().clone() for p in params]
    states = [dict(exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p)) for p in params]
    cu = [torch.nn.Parameter(p.detach().cuda()) for p in params]
    opt = Adam(cu, lr=5e-3, weight_decay=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
    lr = 5e-3
    for step in range(4):
        for i, p in enumerate(cu):
            gcpu = torch.randn_like(params[i])
            if i == 7:
                p.grad = None                     # skipped by the optimiser, its step count stays behind
                continue
            p.grad = gcpu.cuda()
            k = step + 1
            ref[i] = port.adam_step(ref[i], gcpu, states[i], k, lr=lr, weight_decay=1e-3)
        opt.step()
        sched.step()
        lr = sched.get_last_lr()[0]
    for i, p in enumerate(cu):
        assert rel_err(torch.view_as_real(p.detach()).cpu().numpy() if p.is_complex() else p.detach().cpu().numpy(),
                       torch.view_as_real(ref[i]).numpy() if ref[i].is_complex() else ref[i].numpy()) < 5e-6, i
    assert 7 not in [i for i, p in enumerate(cu) if len(opt.state[p])]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", list(LOSS_MODES))
def test_lp_loss_cuda_matches_reference(mode, golden, cuda_lib):
    from uno_b200.losses import LpLoss

    g = golden("training")
    kw = dict(none=dict(reduction=False), sum=dict(size_average=False), mean=dict(size_average=True))[mode]
    x = torch.tensor(g["loss.x"], device="cuda", requires_grad=True)
    lo = LpLoss(**kw)(x, torch.tensor(g["loss.y"], device="cuda"))
    assert tuple(lo.shape) == np.shape(g[f"loss.{mode}"])
    assert rel_err(lo.detach().cpu().numpy(), g[f"loss.{mode}"]) < 1e-6
    lo.backward(torch.tensor(g[f"loss.{mode}.gl"], device="cuda"))
    assert rel_err(x.grad.cpu().numpy(), g[f"loss.{mode}.gx"]) < 2e-6


@pytest.mark.gpu
def test_lp_loss_cuda_large_and_cpu_rejected(cuda_lib):
    from uno_b200.losses import LpLoss

    torch.manual_seed(0)
    x = torch.randn(8, 421, 421, device="cuda", requires_grad=True)
    y = torch.randn(8, 421, 421, device="cuda")
    lo = LpLoss(size_average=False)(x.view(8, -1), y.view(8, -1))
    xr = x.detach().double().requires_grad_(True)
    lr_ = port.LpLoss(size_average=False)(xr.view(8, -1), y.double().view(8, -1))
    assert abs(float(lo) - float(lr_)) < 1e-6 * float(lr_)
    lo.backward()
    lr_.backward()
    assert float((x.grad.double() - xr.grad).abs().max() / xr.grad.abs().max()) < 1e-5
    with pytest.raises(RuntimeError, match="CUDA float32"):
        LpLoss()(torch.zeros(2, 3), torch.ones(2, 3))
    # an exactly matched sample: zero gradient for it, finite everywhere (torch masks the zero norm)
    x2 = torch.randn(3, 1000, device="cuda")
    y2 = torch.randn(3, 1000, device="cuda")
    x2[1] = y2[1]
    x2.requires_grad_(True)
    LpLoss(size_average=False)(x2, y2).backward()
    assert torch.isfinite(x2.grad).all() and float(x2.grad[1].abs().max()) == 0.0
