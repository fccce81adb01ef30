"""Shim for the reference's darcy_flow_uno2d.py: everything it defines, with UNO_9 (the model darcy_flow_main.py:95
trains) replaced by the fused-glue implementation of uno_b200.models.  See INTEGRATION.md."""
from _overlay import overlay as _overlay

_overlay("darcy_flow_uno2d", globals(), ["UNO_9"])
