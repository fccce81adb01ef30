"""CPU: the lift / project entry points of the C ABI (descriptor checks, geometry, pointer plumbing,
pre-zeroed gradient accumulators) on the host-emulation backend, against the fp64 oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from cases import LIFT_CASES, PROJECT_CASES
from conftest import BWD_TOL, FWD_TOL, ROOT, rel_err
from glue_util import lift_inputs, lift_oracle, project_inputs, project_oracle
from uno_b200 import _capi

sys.path.insert(0, os.path.join(ROOT, "tests", "hostemu"))
import emu  # noqa: E402


@pytest.mark.parametrize("name", list(LIFT_CASES))
def test_lift(name):
    case = LIFT_CASES[name]
    _, _, lo, hi, *_ = case
    t = lift_inputs(case)
    ref = lift_oracle(case, t)
    h = emu.lift_fwd(t["a"], t["grid"], t["w_a"], t["b_a"], t["w_b"], t["b_b"], lo, hi)
    assert rel_err(h, ref["h"]) < FWD_TOL
    ga, gwa, gba, gwb, gbb = emu.lift_bwd(t["gh"], t["a"], t["grid"], t["w_a"], t["b_a"], t["w_b"], t["b_b"], lo, hi)
    for got, key in ((ga, "ga"), (gwa, "gw_a"), (gba, "gb_a"), (gwb, "gw_b"), (gbb, "gb_b")):
        assert rel_err(got, ref[key]) < BWD_TOL, key
    # uno_lift_bwd2: the upstream gradient as the sum of two tensors (h feeds the first block AND the projection)
    part = np.random.default_rng(9).standard_normal(np.shape(t["gh"])).astype(np.float32)
    two = emu.lift_bwd(part, t["a"], t["grid"], t["w_a"], t["b_a"], t["w_b"], t["b_b"], lo, hi, gh2=np.asarray(t["gh"], np.float32) - part)
    for got, key in zip(two, ("ga", "gw_a", "gb_a", "gw_b", "gb_b")):
        assert rel_err(got, ref[key]) < BWD_TOL, key


@pytest.mark.parametrize("name", list(PROJECT_CASES))
def test_project(name):
    case = PROJECT_CASES[name]
    _, _, lo, hi, *_ = case
    t = project_inputs(case)
    ref = project_oracle(case, t)
    out, pre = emu.project_fwd(t["srcs"], t["w1"], t["b1"], t["w2"], t["b2"], lo, hi, want_pre=True)
    assert rel_err(out, ref["out"]) < FWD_TOL
    assert np.isfinite(pre).all()
    for saved in (None, pre):       # backward recomputing fc1, and backward fed the saved pre-activations
        gs, gw1, gb1, gw2, gb2 = emu.project_bwd(t["gout"], t["srcs"], t["w1"], t["b1"], t["w2"], lo, hi, pre=saved)
        for g, r in zip(gs, ref["gsrcs"]):
            assert rel_err(g, r) < BWD_TOL
        for got, key in ((gw1, "gw1"), (gb1, "gb1"), (gw2, "gw2"), (gb2, "gb2")):
            assert rel_err(got, ref[key]) < BWD_TOL, key


def test_descriptor_errors():
    L = emu.lib()
    bad = [
        _capi.lift_desc(1, (4,), (0,), (0,), 1, 2, 16, 32),            # 1-D
        _capi.lift_desc(1, (4, 4), (0, 0), (0, 0), 10, 7, 16, 32),      # 17 input channels
        _capi.lift_desc(1, (4, 4), (0, 0), (0, 0), 1, 2, 33, 32),       # hidden too wide
        _capi.lift_desc(1, (4, 4), (0, -1), (0, 0), 1, 2, 16, 32),      # negative padding
        _capi.lift_desc(0, (4, 4), (0, 0), (0, 0), 1, 2, 16, 32),       # empty batch
    ]
    for d in bad:
        assert L.uno_lift_check(C.byref(d)) == 1
        assert L.uno_last_error()
    badp = [
        _capi.project_desc(1, (4, 4), (0, 0), (0, 0), (40, 40), 32, 1),  # 80 channels
        _capi.project_desc(1, (4, 4), (0, 0), (0, 0), (32,), 129, 1),
        _capi.project_desc(1, (4, 4), (0, 0), (0, 0), (32,), 32, 5),
        _capi.project_desc(1, (4, 4), (0, 0), (0, 0), (), 32, 1),
    ]
    for d in badp:
        assert L.uno_project_check(C.byref(d)) == 1
    ok = _capi.lift_desc(2, (4, 4, 3), (0, 0, 1), (0, 0, 1), 1, 5, 12, 8)
    assert L.uno_lift_check(C.byref(ok)) == 0
