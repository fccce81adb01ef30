"""ctypes prototypes for the C ABI declared in include/uno_b200.h.

``bind(path)`` loads a shared library exporting that ABI and attaches argument / return types.
The product loads exactly one library through ``uno_b200._lib`` (the nvcc-built
``uno_b200/csrc/libuno_b200.so``); the CPU test-suite also binds the host-emulation build under
``tests/hostemu`` with the same prototypes to check the orchestration without a GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

c_float_p = C.POINTER(C.c_float)


class ConvDesc(C.Structure):
    _fields_ = [
        ("ndim", C.c_int),
        ("batch", C.c_int),
        ("in_ch", C.c_int),
        ("out_ch", C.c_int),
        ("in_dim", C.c_int * 3),
        ("out_dim", C.c_int * 3),
        ("modes", C.c_int * 3),
    ]


class BlockDesc(C.Structure):
    _fields_ = [("conv", ConvDesc), ("normalize", C.c_int), ("non_lin", C.c_int), ("eps", C.c_float)]


class PixelDesc(C.Structure):
    _fields_ = [("ndim", C.c_int), ("batch", C.c_int), ("dim", C.c_int * 3), ("pad_lo", C.c_int * 3), ("pad_hi", C.c_int * 3)]


class LiftDesc(C.Structure):
    _fields_ = [("px", PixelDesc), ("raw_ch", C.c_int), ("grid_ch", C.c_int), ("hidden", C.c_int), ("out_ch", C.c_int)]


class AdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("max_exp_avg_sq", C.c_void_p), ("numel", C.c_long), ("is_complex", C.c_int)]


class AdamHyper(C.Structure):
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("weight_decay", C.c_double),
                ("amsgrad", C.c_int), ("step", C.c_int)]


class ProjectDesc(C.Structure):
    _fields_ = [("px", PixelDesc), ("nsrc", C.c_int), ("src_ch", C.c_int * 4), ("hidden", C.c_int), ("out_ch", C.c_int)]


# every symbol include/uno_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_DESC = C.POINTER(ConvDesc)
_BDESC = C.POINTER(BlockDesc)
_PP = C.POINTER(C.c_void_p)
_LDESC = C.POINTER(LiftDesc)
_PDESC = C.POINTER(ProjectDesc)
SYMBOLS = {
    "uno_last_error": (C.c_char_p, []),
    "uno_version": (C.c_int, []),
    "uno_backend_name": (C.c_char_p, []),
    "uno_clear_plans": (None, []),
    "uno_spectral_conv_check": (C.c_int, [_DESC]),
    "uno_spectral_conv_workspace_bytes": (C.c_size_t, [_DESC]),
    "uno_spectral_conv_xhat_elems": (C.c_size_t, [_DESC]),
    "uno_spectral_conv_fwd": (C.c_int, [_DESC, _P, _PP, _P, _P, _P, C.c_size_t, _P]),
    "uno_spectral_conv_bwd": (C.c_int, [_DESC, _P, _P, _PP, _P, _PP, C.c_int, _P, C.c_size_t, _P]),
    "uno_pointwise_workspace_bytes": (C.c_size_t, [_DESC]),
    "uno_pointwise_saved_elems": (C.c_size_t, [_DESC]),
    "uno_pointwise_fwd": (C.c_int, [_DESC, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "uno_pointwise_bwd": (C.c_int, [_DESC, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "uno_operator_block_workspace_bytes": (C.c_size_t, [_BDESC]),
    "uno_operator_block_fwd": (C.c_int, [_BDESC, _P, _PP, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "uno_operator_block_bwd": (
        C.c_int,
        [_BDESC, _P, _P, _P, _P, _P, _P, _PP, _P, _P, _P, _P, _PP, _P, _P, _P, _P, _P, C.c_size_t, _P],
    ),
    "uno_operator_block_bwd2": (
        C.c_int,
        [_BDESC, _P, C.c_long, _P, C.c_long, _P, _P, _P, _P, _P, _PP, _P, _P, _P, _P, _PP, _P, _P, _P, _P, _P, C.c_size_t, _P],
    ),
    "uno_lift_check": (C.c_int, [_LDESC]),
    "uno_lift_fwd": (C.c_int, [_LDESC, _P, _P, _P, _P, _P, _P, _P, _P]),
    "uno_lift_bwd": (C.c_int, [_LDESC, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "uno_lift_bwd2": (C.c_int, [_LDESC, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "uno_project_check": (C.c_int, [_PDESC]),
    "uno_project_fwd": (C.c_int, [_PDESC, _PP, _P, _P, _P, _P, _P, _P, _P]),
    "uno_project_bwd": (C.c_int, [_PDESC, _P, _PP, _P, _P, _P, _P, _PP, _P, _P, _P, _P, _P]),
    "uno_adam_step": (C.c_int, [C.POINTER(AdamTensor), C.c_int, C.POINTER(AdamHyper), _P]),
    "uno_lp_loss_fwd": (C.c_int, [_P, _P, C.c_int, C.c_long, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "uno_lp_loss_bwd": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_long, C.c_int, _P, _P]),
    "uno_config_set": (C.c_int, [C.c_char_p, C.c_int]),
    "uno_config_get": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "uno_config_name": (C.c_char_p, [C.c_int]),
    "uno_launch_count": (C.c_long, []),
    "uno_profile_enable": (None, [C.c_int]),
    "uno_profile_report": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "uno_profile_report_levels": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "uno_plan_dft_last_analysis": (C.c_int, [C.c_int, C.c_int, C.c_double, _P]),
    "uno_plan_dft_last_synthesis": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int, _P]),
    "uno_plan_dft_mid_analysis": (C.c_int, [C.c_int, C.c_int, _P]),
    "uno_plan_dft_mid_synthesis": (C.c_int, [C.c_int, C.c_int, _P]),
    "uno_plan_sr_mid": (C.c_int, [C.c_int, C.c_int, _P]),
    "uno_plan_sr_last_modes": (C.c_int, [C.c_int, C.c_int]),
    "uno_plan_sr_mid_fixed": (C.c_int, [C.c_int, C.c_int, _P]),
    "uno_plan_sr_last_modes_fixed": (C.c_int, [C.c_int, C.c_int]),
    "uno_plan_bicubic_aa": (C.c_int, [C.c_int, C.c_int, C.c_int, _P]),
    "uno_plan_band_groups": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_int)]),
}


def bind(path: str) -> C.CDLL:
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


def conv_desc(batch: int, in_ch: int, out_ch: int, in_dims: Sequence[int], out_dims: Sequence[int], modes: Sequence[int] = ()) -> ConvDesc:
    d = ConvDesc()
    d.ndim = len(in_dims)
    d.batch, d.in_ch, d.out_ch = int(batch), int(in_ch), int(out_ch)
    for a in range(3):
        d.in_dim[a] = int(in_dims[a]) if a < len(in_dims) else 1
        d.out_dim[a] = int(out_dims[a]) if a < len(out_dims) else 1
        d.modes[a] = int(modes[a]) if a < len(modes) else 0
    return d


def block_desc(conv: ConvDesc, normalize: bool, non_lin: bool, eps: float = 1e-5) -> BlockDesc:
    b = BlockDesc()
    b.conv = conv
    b.normalize, b.non_lin, b.eps = int(bool(normalize)), int(bool(non_lin)), float(eps)
    return b


def pixel_desc(batch: int, dims: Sequence[int], pad_lo: Sequence[int], pad_hi: Sequence[int]) -> PixelDesc:
    p = PixelDesc()
    p.ndim, p.batch = len(dims), int(batch)
    for a in range(3):
        p.dim[a] = int(dims[a]) if a < len(dims) else 1
        p.pad_lo[a] = int(pad_lo[a]) if a < len(dims) else 0
        p.pad_hi[a] = int(pad_hi[a]) if a < len(dims) else 0
    return p


def lift_desc(batch, dims, pad_lo, pad_hi, raw_ch, grid_ch, hidden, out_ch) -> LiftDesc:
    d = LiftDesc()
    d.px = pixel_desc(batch, dims, pad_lo, pad_hi)
    d.raw_ch, d.grid_ch, d.hidden, d.out_ch = int(raw_ch), int(grid_ch), int(hidden), int(out_ch)
    return d


def project_desc(batch, dims, pad_lo, pad_hi, src_ch: Sequence[int], hidden, out_ch) -> ProjectDesc:
    d = ProjectDesc()
    d.px = pixel_desc(batch, dims, pad_lo, pad_hi)
    d.nsrc = len(src_ch)
    for i in range(4):
        d.src_ch[i] = int(src_ch[i]) if i < len(src_ch) else 0
    d.hidden, d.out_ch = int(hidden), int(out_ch)
    return d


def ptr_array(ptrs: Sequence[int]):
    """void*[n] from integer addresses (0 -> NULL)."""
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p if p else None
    return arr


class UnoError(RuntimeError):
    """Non-zero status from the C ABI (message from uno_last_error())."""


def check(lib: C.CDLL, rc: int) -> None:
    if rc != 0:
        msg = lib.uno_last_error()
        raise UnoError((msg.decode() if msg else "unknown error") + f" [uno status {rc}]")
