"""Loader of the product CUDA library.  There is NO fallback: if ``libuno_b200.so`` is missing or does
not export the full C ABI, importing the operators fails loudly."""
from __future__ import annotations

import os
import threading

from . import _capi

_LIB = None
_LOCK = threading.Lock()
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libuno_b200.so")


def get():
    global _LIB
    if _LIB is None:
        with _LOCK:
            if _LIB is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"uno_b200: CUDA library not built ({LIB_PATH} missing). Run `python -m uno_b200.build` "
                        "(needs nvcc); there is no CPU / PyTorch fallback for these operators."
                    )
                lib = _capi.bind(LIB_PATH)
                name = lib.uno_backend_name()
                if name != b"cuda-sm100a":
                    raise RuntimeError(f"uno_b200: unexpected backend {name!r} in {LIB_PATH}")
                _LIB = lib
    return _LIB
