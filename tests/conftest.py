import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# fp32 tolerances of the parity contract: relative to the largest magnitude of the
# tensor being compared.  The reference's own fp32 noise floor against an fp64 restatement is
# ~4e-7 forward and ~6e-6 backward (tests/golden/make_golden.py cases).
FWD_TOL = 2e-5
BWD_TOL = 5e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "multigpu: needs two CUDA devices (run with `gpurun --gpus 2 -- python -m pytest tests -m multigpu`)")


def pytest_collection_modifyitems(config, items):
    """Two-GPU tests run only when asked for by name (-m multigpu): a one-GPU box or the CPU suite neither runs nor skips them."""
    if "multigpu" in (config.getoption("-m") or ""):
        return
    keep, drop = [], []
    for it in items:
        (drop if it.get_closest_marker("multigpu") else keep).append(it)
    if drop:
        config.hook.pytest_deselected(items=drop)
        items[:] = keep


def rel_err(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / den


@pytest.fixture(scope="session")
def golden():
    class G:
        def __init__(self):
            self._c = {}

        def __call__(self, name):
            if name not in self._c:
                self._c[name] = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
            return self._c[name]

    return G()


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if needed) and load the product CUDA library; GPU tests go through it."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from uno_b200 import _lib, build

    build.build()
    return _lib.get()
