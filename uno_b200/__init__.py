"""uno_b200 -- B200-native (sm_100a) implementation of the U-NO integral-operator hot path.

    from uno_b200.integral_operators import OperatorBlock_2D, SpectralConv2d_Uno, ...   # drop-in modules
    from uno_b200.models import UNO_9, UNO, Uno3D_T10                                  # U-shaped callers
    from uno_b200.parallel import GradReducer                                           # batch-shard DP

The CUDA library (uno_b200/csrc/libuno_b200.so, C ABI in include/uno_b200.h) is loaded lazily on the
first operator call and there is no CPU fallback.
"""
__version__ = "0.1.0"
