"""CPU: the thread-by-thread numpy restatements of the opt-in tensor-core kernels' index arithmetic (tools/emu/: operand-image
layouts as the UMMA descriptors read them, gather / scatter maps, tile and tail handling, the host image builders) still agree
with a plain product.  They restate the kernels, they do not run them: the kernels' own parity tests are the opt-in GPU tests
(tests/test_gpu_experimental.py)."""
import glob
import os
import subprocess
import sys

import pytest

from conftest import ROOT

SCRIPTS = sorted(glob.glob(os.path.join(ROOT, "tools", "emu", "emu_*.py")))


@pytest.mark.parametrize("script", SCRIPTS, ids=[os.path.basename(s) for s in SCRIPTS])
def test_index_emulation(script):
    r = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
