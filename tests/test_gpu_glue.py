"""GPU: the fused lift / project kernels (uno_b200/csrc/pixel_mlp.cuh) through the C ABI and the autograd shims,
against the fp64 oracle on the seeded cases, plus size-independent properties at the Darcy size."""
import numpy as np
import pytest
import torch

from cases import LIFT_CASES, PROJECT_CASES
from conftest import BWD_TOL, FWD_TOL, rel_err
from glue_util import lift_inputs, lift_oracle, project_inputs, project_oracle

pytestmark = pytest.mark.gpu


def _cu(v, grad=False):
    return torch.tensor(v, device="cuda", requires_grad=grad)


@pytest.mark.parametrize("name", list(LIFT_CASES))
def test_lift(name, cuda_lib):
    from uno_b200 import functional as Fn

    case = LIFT_CASES[name]
    _, _, lo, hi, *_ = case
    t = lift_inputs(case)
    ref = lift_oracle(case, t)
    a, wa, ba, wb, bb = (_cu(t[k], True) for k in ("a", "w_a", "b_a", "w_b", "b_b"))
    h = Fn.lift(a, _cu(t["grid"]), wa, ba, wb, bb, lo, hi)
    assert rel_err(h.detach().cpu().numpy(), ref["h"]) < FWD_TOL
    h.backward(_cu(t["gh"]))
    for got, key in ((a, "ga"), (wa, "gw_a"), (ba, "gb_a"), (wb, "gw_b"), (bb, "gb_b")):
        assert rel_err(got.grad.cpu().numpy(), ref[key]) < BWD_TOL, key


@pytest.mark.parametrize("name", list(PROJECT_CASES))
def test_project(name, cuda_lib):
    from uno_b200 import functional as Fn

    case = PROJECT_CASES[name]
    _, _, lo, hi, *_ = case
    t = project_inputs(case)
    ref = project_oracle(case, t)
    srcs = [_cu(s, True) for s in t["srcs"]]
    w1, b1, w2, b2 = (_cu(t[k], True) for k in ("w1", "b1", "w2", "b2"))
    out = Fn.project(srcs, w1, b1, w2, b2, lo, hi)
    assert rel_err(out.detach().cpu().numpy(), ref["out"]) < FWD_TOL
    out.backward(_cu(t["gout"]))
    for s, r in zip(srcs, ref["gsrcs"]):
        assert rel_err(s.grad.cpu().numpy(), r) < BWD_TOL
    for got, key in ((w1, "gw1"), (b1, "gb1"), (w2, "gw2"), (b2, "gb2")):
        assert rel_err(got.grad.cpu().numpy(), ref[key]) < BWD_TOL, key


def test_lift_no_input_grad_and_inference(cuda_lib):
    from uno_b200 import functional as Fn

    case = LIFT_CASES["darcy"]
    _, _, lo, hi, *_ = case
    t = lift_inputs(case, seed=3)
    ref = lift_oracle(case, t)
    wa, ba, wb, bb = (_cu(t[k], True) for k in ("w_a", "b_a", "w_b", "b_b"))
    h = Fn.lift(_cu(t["a"]), _cu(t["grid"]), wa, ba, wb, bb, lo, hi)      # data does not require grad (Darcy)
    h.backward(_cu(t["gh"]))
    assert rel_err(wa.grad.cpu().numpy(), ref["gw_a"]) < BWD_TOL
    with torch.no_grad():
        h2 = Fn.lift(_cu(t["a"]), _cu(t["grid"]), wa, ba, wb, bb, lo, hi)
    assert torch.equal(h2, h.detach())


def test_glue_matches_torch_at_darcy_size(cuda_lib):
    """Full BASELINE size (421^2 -> 481^2, batch 4): the fused kernels against the same chain of torch CUDA ops
    (fp32, TF32 off), forward and every gradient; also padding is exactly zero and the run is repeatable."""
    from oracle import uno_torch_port as port
    from uno_b200 import functional as Fn

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(5)
    R = lambda *s: torch.randn(*s, device="cuda", generator=g)
    B, S, P = 4, 421, 60
    a = R(B, S, S, 1)
    grid = R(S, S, 2)
    prm = [R(16, 3) / 3**0.5, 0.3 * R(16), R(32, 16) / 4, 0.3 * R(32)]
    gh = R(B, 32, S + P, S + P)
    outs = []
    for fn in (Fn.lift, port.lift):
        ps = [p.clone().requires_grad_(True) for p in prm]
        h = fn(a, grid, *ps, (0, 0), (P, P))
        h.backward(gh)
        outs.append((h.detach(), [p.grad for p in ps]))
    (h0, g0), (h1, g1) = outs
    assert float((h0 - h1).abs().max() / h1.abs().max()) < FWD_TOL
    assert float(h0[..., S:, :].abs().max()) == 0.0 and float(h0[..., :, S:].abs().max()) == 0.0
    for x, y in zip(g0, g1):
        assert float((x - y).abs().max() / y.abs().max()) < BWD_TOL
    # projection
    srcs = [R(B, 32, S + P, S + P), h0]
    prm = [R(32, 64) / 8, 0.3 * R(32), R(1, 32) / 32**0.5, 0.3 * R(1)]
    go = R(B, S, S, 1)
    outs = []
    for fn in (Fn.project, port.project):
        ss = [s.clone().requires_grad_(True) for s in srcs]
        ps = [p.clone().requires_grad_(True) for p in prm]
        o = fn(ss, *ps, (0, 0), (P, P))
        o.backward(go)
        outs.append((o.detach(), [s.grad for s in ss] + [p.grad for p in ps]))
    (o0, g0), (o1, g1) = outs
    assert float((o0 - o1).abs().max() / o1.abs().max()) < FWD_TOL
    for x, y in zip(g0, g1):
        assert float((x - y).abs().max() / y.abs().max()) < BWD_TOL
    assert float(g0[0][..., S:, :].abs().max()) == 0.0 and float(g0[1][..., :, S:].abs().max()) == 0.0


@pytest.mark.parametrize("name", ["darcy", "tc_two_chunks", "tc_wide", "tc_one_chunk"])
def test_project_backward_fp32_kernel(name, cuda_lib):
    """The fp32 projection kernels, forward and backward (switch proj_simt), against the fp64 oracle on the shapes whose
    default is the tcgen05 pair (which test_project covers), same tolerance."""
    from uno_b200 import config
    from uno_b200 import functional as Fn

    case = PROJECT_CASES[name]
    _, _, lo, hi, *_ = case
    t = project_inputs(case, seed=2)
    ref = project_oracle(case, t)
    srcs = [_cu(s, True) for s in t["srcs"]]
    w1, b1, w2, b2 = (_cu(t[k], True) for k in ("w1", "b1", "w2", "b2"))
    with config.switches(proj_simt=1):
        out = Fn.project(srcs, w1, b1, w2, b2, lo, hi)
        out.backward(_cu(t["gout"]))
        torch.cuda.synchronize()
    assert rel_err(out.detach().cpu().numpy(), ref["out"]) < FWD_TOL      # the fp32 forward too (default: tcgen05 for > 32 channels)
    for s, r in zip(srcs, ref["gsrcs"]):
        assert rel_err(s.grad.cpu().numpy(), r) < BWD_TOL
    for got, key in ((w1, "gw1"), (b1, "gb1"), (w2, "gw2"), (b2, "gb2")):
        assert rel_err(got.grad.cpu().numpy(), ref[key]) < BWD_TOL, key
