"""Overlay loader for the model-file shims: execute the reference's own model file (found further down ``sys.path``) into
the shim module, so every class and helper it defines stays importable, then swap in the classes that ``uno_b200.models``
provides with the fused lift / projection kernels.  Constructor signatures, sub-module names and state_dict keys of the
swapped classes are the reference's, so drivers and checkpoints do not notice."""
import importlib.machinery
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def overlay(module_name: str, namespace: dict, swapped) -> None:
    for entry in sys.path:
        if not entry or os.path.abspath(entry) == HERE:
            continue
        spec = importlib.machinery.PathFinder.find_spec(module_name, [entry])
        if spec is None or spec.origin is None or os.path.dirname(os.path.abspath(spec.origin)) == HERE:
            continue
        with open(spec.origin) as f:
            code = compile(f.read(), spec.origin, "exec")
        exec(code, namespace)               # the reference file itself: UNO_11, UNO_S256, Uno3D_T40, ... stay available
        break
    else:
        raise ImportError(f"uno_b200 drop-in: the reference's {module_name}.py was not found on sys.path behind the shim directory")
    from uno_b200 import models

    for name in swapped:
        namespace[name] = getattr(models, name)
