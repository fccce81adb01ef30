"""CPU: host-side contract of the drop-in modules (no kernel call): names, constructor defaults,
state_dict schema, seeded init parity with the oracle port, loud failure on CPU tensors."""
import os
import sys

import pytest
import torch

from conftest import ROOT
from oracle import uno_torch_port as port


def test_dropin_module_exports():
    sys.path.insert(0, os.path.join(ROOT, "uno_b200", "dropin"))
    try:
        import integral_operators as io
    finally:
        sys.path.pop(0)
    for name in ["SpectralConv1d_Uno", "SpectralConv2d_Uno", "SpectralConv3d_Uno", "pointwise_op_1D", "pointwise_op_2D",
                 "pointwise_op_3D", "OperatorBlock_1D", "OperatorBlock_2D", "OperatorBlock_3D", "torch", "np", "nn", "F"]:
        assert hasattr(io, name), name


def test_state_dict_schema_and_seeded_init_match_port():
    from uno_b200 import integral_operators as ops

    for ctor, args, kw in [("OperatorBlock_2D", (3, 5, 8, 8, 3, 3), dict(Normalize=True)), ("OperatorBlock_3D", (2, 3, 8, 8, 6, 3, 3, 2), {}),
                           ("SpectralConv2d_Uno", (4.0, 6.0, 16, 16), {}), ("SpectralConv3d_Uno", (1, 2, 4, 4, 6), {})]:
        torch.manual_seed(7)
        a = getattr(ops, ctor)(*args, **kw)
        torch.manual_seed(7)
        b = getattr(port, ctor)(*args, **kw)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert sa[k].dtype == sb[k].dtype and sa[k].shape == sb[k].shape
            assert torch.equal(torch.view_as_real(sa[k]) if sa[k].is_complex() else sa[k], torch.view_as_real(sb[k]) if sb[k].is_complex() else sb[k]), k


def test_appendix_d_schema_uno9():
    from uno_b200 import models

    m = models.UNO_9(3, 32)
    sd = m.state_dict()
    assert sd["conv0.conv.weights1"].shape == (32, 64, 18, 18) and sd["conv0.conv.weights1"].dtype == torch.complex64
    assert sd["conv0.w.conv.weight"].shape == (64, 32, 1, 1)
    assert sd["conv1.normalize_layer.weight"].shape == (128,)
    assert sd["conv5.conv.weights2"].shape == (128, 32, 18, 18)
    assert sd["fc2.weight"].shape == (1, 32)
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params == 8_218_049      # SURVEY.md Appendix D (complex counted once)


def test_cpu_tensor_fails_loudly():
    from uno_b200 import integral_operators as ops

    with pytest.raises(RuntimeError, match="CUDA"):
        ops.OperatorBlock_2D(2, 2, 8, 8, 3, 3)(torch.randn(1, 2, 8, 8), 8, 8)
    with pytest.raises(ValueError):
        ops.pointwise_op_1D(2, 2, 8)(torch.randn(1, 2, 8))


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle / a CPU fallback."""
    import re

    for dirpath, _, files in os.walk(os.path.join(ROOT, "uno_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                # the host-emulation library is only ever mentioned in comments / docstrings, never loaded
                assert "libuno_hostemu" not in src, f


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference checkout (build container only)")
def test_model_file_shims_overlay_the_reference():
    """uno_b200/dropin in front of the reference on sys.path: `from darcy_flow_uno2d import UNO_9, UNO_11` (darcy_flow_main.py:9)
    gets our fused-glue UNO_9 and the reference's own UNO_11 (built on the drop-in blocks); same for the Navier-Stokes files."""
    import subprocess

    code = r'''
import sys, types
for n in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(n, types.ModuleType(n))
sys.dont_write_bytecode = True
from darcy_flow_uno2d import UNO_9, UNO_11
from navier_stokes_uno2d import UNO, UNO_P, UNO_S256
from navier_stokes_uno3d import Uno3D_T10, Uno3D_T40
import uno_b200.models as M, uno_b200.integral_operators as ops
assert UNO_9 is M.UNO_9 and UNO is M.UNO and UNO_P is M.UNO_P and Uno3D_T10 is M.Uno3D_T10
assert UNO_11.__module__ != "uno_b200.models" and UNO_11.__init__.__code__.co_filename.startswith("/root/reference")
m = UNO_S256(14, 8)                                        # (UNO_11 itself does not construct upstream: SURVEY.md B.4)
assert isinstance(m.L0, ops.OperatorBlock_2D)             # the reference's own model, on the CUDA drop-in blocks
import torch
torch.manual_seed(0); a = UNO_9(3, 8, pad=5)
assert [k for k in a.state_dict()][:4] == ["fc_n1.weight", "fc_n1.bias", "fc0.weight", "fc0.bias"]
print("ok")
'''
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "uno_b200", "dropin"), ROOT, "/root/reference"]),
               PYTHONDONTWRITEBYTECODE="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
