"""Shim: put this directory in front of the reference on sys.path and its model files
(`from integral_operators import *`) run on the B200 kernels unchanged.  See INTEGRATION.md."""
from uno_b200.integral_operators import *  # noqa: F401,F403
from uno_b200.integral_operators import __all__  # noqa: F401
