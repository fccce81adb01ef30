"""The reference's optimiser (``Adam.py``: ``Adam(params, lr, betas, eps, weight_decay, amsgrad)``) as one multi-tensor
CUDA kernel behind the C ABI (``uno_adam_step``).

Same constructor, ``param_groups`` / ``state`` layout and state keys (``step``, ``exp_avg``, ``exp_avg_sq``,
``max_exp_avg_sq``) as upstream, so ``torch.optim.lr_scheduler.StepLR`` (train_darcy.py:38) and optimiser checkpoints
work unchanged.  Semantics follow Adam.py:23-52, which is NOT torch.optim.Adam for complex weights: the second moment is
the running mean of ``g * conj(g)`` (kept, like upstream, in a complex tensor with zero imaginary part), so the real and
imaginary parts of a spectral weight share one denominator.  The reference issues ~10 elementwise launches per tensor
from a Python loop; here every group is one launch per 24 tensors.  Parameters must be CUDA float32 / complex64."""
from __future__ import annotations

import ctypes as C

import torch
from torch.optim.optimizer import Optimizer

from . import _capi
from ._lib import get as _get_lib


class Adam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        # same validation and messages as Adam.py:88-97
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if not 0.0 <= weight_decay:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault("amsgrad", False)

    @staticmethod
    def _floats(t: torch.Tensor):
        """(pointer, float count, is_complex) of a dense fp32 / complex64 CUDA tensor."""
        if not t.is_cuda or t.dtype not in (torch.float32, torch.complex64):
            raise RuntimeError(f"uno_b200.optim.Adam: expected CUDA float32 / complex64 tensors (got {t.dtype} on {t.device})")
        if not t.is_contiguous():
            raise RuntimeError("uno_b200.optim.Adam: parameters, gradients and state must be contiguous")
        cx = t.is_complex()
        return t.data_ptr(), t.numel() * (2 if cx else 1), cx

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _get_lib()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            by_step = {}
            keep = []          # tensors made contiguous for the call must outlive the launch
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                state = self.state[p]
                if len(state) == 0:                       # lazy state initialisation, as upstream
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if group["amsgrad"]:
                        state["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                if group["amsgrad"] and p.is_complex():
                    raise RuntimeError("uno_b200.optim.Adam: amsgrad is not defined for complex parameters (torch.maximum "
                                       "has no complex kernel; the reference raises here as well)")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                keep.append(g)
                pp, n, cx = self._floats(p)
                t = _capi.AdamTensor()
                t.param, t.numel, t.is_complex = pp, n, int(cx)
                t.grad = self._floats(g)[0]
                t.exp_avg = self._floats(state["exp_avg"])[0]
                t.exp_avg_sq = self._floats(state["exp_avg_sq"])[0]
                t.max_exp_avg_sq = self._floats(state["max_exp_avg_sq"])[0] if group["amsgrad"] else None
                by_step.setdefault(state["step"], []).append(t)
            for step, tensors in by_step.items():         # normally a single step value per group
                arr = (_capi.AdamTensor * len(tensors))(*tensors)
                h = _capi.AdamHyper(float(group["lr"]), float(beta1), float(beta2), float(group["eps"]), float(group["weight_decay"]),
                                    int(bool(group["amsgrad"])), int(step))
                dev = group["params"][0].device
                st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                _capi.check(lib, lib.uno_adam_step(arr, len(tensors), C.byref(h), st))
            del keep
        return loss
