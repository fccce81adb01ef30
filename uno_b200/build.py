"""Build recipe for the product library ``uno_b200/csrc/libuno_b200.so`` (nvcc, sm_100a only).

The library is built IN-TREE so that it travels with the repo snapshot to the GPU box; it is
git-ignored.  ``build()`` is what ``__graft_entry__.build()`` calls.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, "libuno_b200.so")

SOURCES = ["backend_cuda.cu", "uno_api.cpp", "plan.cpp", "config.cpp"]
HEADERS = ["backend.h", "plan.h", "config.h", os.path.join(ROOT, "include", "uno_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    "-shared", "-ldl",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the uno_b200 CUDA library cannot be built")


def _deps():
    out = [os.path.join(CSRC, s) for s in SOURCES]
    out += [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    out += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd, cwd=ROOT)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
