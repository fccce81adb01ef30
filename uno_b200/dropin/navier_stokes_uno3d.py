"""Shim for the reference's navier_stokes_uno3d.py: everything it defines, with Uno3D_T10 (ns_uno3d_main.py:103)
replaced by the implementation of uno_b200.models.  See INTEGRATION.md."""
from _overlay import overlay as _overlay

_overlay("navier_stokes_uno3d", globals(), ["Uno3D_T10"])
