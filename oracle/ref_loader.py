"""
TEST / BASELINE INFRASTRUCTURE ONLY -- makes the UNMODIFIED reference importable where /root/reference is absent.

The reference is pure Python; it has no build system.  ``stage()`` copies its hot-path
source files, byte for byte, from ``/root/reference`` into the git-ignored ``oracle/_ref/`` (the same
place a compiled reference would put its ``.so``), so the copy travels to the GPU box with the snapshot
without entering the history.  ``__graft_entry__.build()`` calls ``stage()``; nothing is ever edited.
``load()`` imports the staged files (falling back to ``/root/reference`` itself in the build container)
with the two-line ``matplotlib`` stub the model files need.

Users: ``bench.py --impl reference`` / ``cpu_baseline`` / the ``reference_gpu`` leg (the reference's own
models timed on the host cores and through torch-CUDA on the same B200), and tests.  The product
(``uno_b200/``) never imports this.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
FILES = ["integral_operators.py", "darcy_flow_uno2d.py", "navier_stokes_uno2d.py", "navier_stokes_uno3d.py",
         "utilities3.py", "Adam.py"]
MANIFEST = "MANIFEST.json"


def _sha(path: str) -> str:
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage() -> bool:
    """Copy the reference's files into oracle/_ref/ (build container only).  Returns True when the staged copy exists."""
    if os.path.isdir(REF_SRC):
        os.makedirs(REF_DST, exist_ok=True)
        man = {}
        for f in FILES:
            src = os.path.join(REF_SRC, f)
            if not os.path.exists(src):
                continue
            dst = os.path.join(REF_DST, f)
            if not os.path.exists(dst) or _sha(dst) != _sha(src):
                shutil.copyfile(src, dst)
            man[f] = _sha(dst)
        with open(os.path.join(REF_DST, MANIFEST), "w") as fh:
            json.dump({"source": REF_SRC, "sha256": man}, fh, indent=1)
    return available()


def available() -> bool:
    return os.path.exists(os.path.join(REF_DST, "integral_operators.py")) or os.path.isdir(REF_SRC)


def verify() -> bool:
    """The staged files are the bytes the manifest recorded (nobody edited the reference)."""
    path = os.path.join(REF_DST, MANIFEST)
    if not os.path.exists(path):
        return False
    man = json.load(open(path))["sha256"]
    return all(os.path.exists(os.path.join(REF_DST, f)) and _sha(os.path.join(REF_DST, f)) == h for f, h in man.items())


def ref_dir() -> str:
    if os.path.exists(os.path.join(REF_DST, "integral_operators.py")):
        return REF_DST
    if os.path.isdir(REF_SRC):
        return REF_SRC
    raise ImportError("the reference is neither staged under oracle/_ref/ nor present at /root/reference")


_loaded = {}


def load(module: str):
    """Import one of the reference's modules (``integral_operators``, ``darcy_flow_uno2d``, ...) under its own name."""
    if module in _loaded:
        return _loaded[module]
    d = ref_dir()
    for name in ("matplotlib", "matplotlib.pyplot"):      # imported by the model files, unused on this path
        sys.modules.setdefault(name, types.ModuleType(name))
    old_flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    sys.path.insert(0, d)
    try:
        mod = importlib.import_module(module)
    finally:
        sys.path.remove(d)
        sys.dont_write_bytecode = old_flag
    got = os.path.dirname(os.path.abspath(getattr(mod, "__file__", "")))
    if got != os.path.abspath(d):
        raise ImportError(f"'{module}' resolved to {got}, not to the reference in {d}")
    _loaded[module] = mod
    return mod


# the models of BASELINE.json's configs: workload -> (module, class)
MODELS = {
    "darcy": ("darcy_flow_uno2d", "UNO_9"),
    "ns2d": ("navier_stokes_uno2d", "UNO"),
    "ns2d_ar": ("navier_stokes_uno2d", "UNO"),
    "ns3d": ("navier_stokes_uno3d", "Uno3D_T10"),
}


def model_class(workload: str):
    mod, cls = MODELS[workload]
    return getattr(load(mod), cls)


def lp_loss():
    return load("utilities3").LpLoss
