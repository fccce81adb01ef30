"""Run-time kernel-selection switches of the CUDA library (include/uno_b200.h ``uno_config_*``).

The library seeds them from ``UNO_B200_<NAME>`` once, when it is first used; afterwards they change only through
these calls.  ``switches(...)`` is the context manager the variant-parity tests use."""
from __future__ import annotations

import contextlib
import ctypes as C

from . import _capi
from ._lib import get as _get_lib


def names():
    lib = _get_lib()
    out, i = [], 0
    while True:
        n = lib.uno_config_name(i)
        if n is None:
            return out
        out.append(n.decode())
        i += 1


def get(name: str) -> int:
    lib = _get_lib()
    v = C.c_int(0)
    _capi.check(lib, lib.uno_config_get(name.encode(), C.byref(v)))
    return int(v.value)


def set(name: str, value: int) -> None:  # noqa: A001 - mirrors uno_config_set
    lib = _get_lib()
    _capi.check(lib, lib.uno_config_set(name.encode(), int(value)))


@contextlib.contextmanager
def switches(**kv):
    """``with switches(mid_tc=0, cmm_tc=0): ...`` -- restores the previous values on exit."""
    old = {k: get(k) for k in kv}
    try:
        for k, v in kv.items():
            set(k, v)
        yield
    finally:
        for k, v in old.items():
            set(k, v)
