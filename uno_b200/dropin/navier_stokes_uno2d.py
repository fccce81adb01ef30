"""Shim for the reference's navier_stokes_uno2d.py: everything it defines, with UNO and UNO_P (ns_uno2d_main.py:89)
replaced by the implementations of uno_b200.models.  See INTEGRATION.md."""
from _overlay import overlay as _overlay

_overlay("navier_stokes_uno2d", globals(), ["UNO", "UNO_P"])
