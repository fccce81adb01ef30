"""Batch-sharded data parallelism for the U-NO models.

Every operator on the path is independent across the batch index except the parameter gradients, so
the only exchange is one SUM all-reduce of the gradients per step (the reference losses are SUMS over
samples -- LpLoss(size_average=False), train_darcy.py:42 -- hence SUM, not mean).  >99.7 % of the
payload is complex spectral weights; torch's DistributedDataParallel has no explicit complex support,
so gradients live in ONE flat fp32 buffer (complex parameters as (re, im) pairs, each ``param.grad`` a
view into it) that is all-reduced over NCCL / NVLink in buckets, launched from autograd hooks as soon
as a bucket's gradients are final so the transfer overlaps the rest of backward.
"""
from __future__ import annotations

import contextlib
from typing import List, Optional

import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, module: torch.nn.Module, process_group=None, bucket_mb: float = 32.0, overlap: bool = True):
        self.group = process_group
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev = self.params[0].device
        sizes = [p.numel() * (2 if p.is_complex() else 1) for p in self.params]
        # every slot starts on a 16-byte boundary (complex views need an even element offset; vector kernels like 16 B)
        slots = [(n + 3) // 4 * 4 for n in sizes]
        self.flat = torch.zeros(sum(slots), dtype=torch.float32, device=dev)
        # gradients become final in roughly reverse registration order: lay the buffer out that way
        order = list(range(len(self.params)))[::-1]
        self._bucket_of = {}
        self._slot = {}                         # param index -> (offset, floats) in the flat buffer
        self._sync = True
        self.buckets: List[List[int]] = []      # [start, end, n_params]
        off, cur_start, cur_n = 0, 0, 0
        limit = int(bucket_mb * (1 << 20) / 4)
        for i in order:
            p, n = self.params[i], sizes[i]
            if p.dtype not in (torch.float32, torch.complex64):
                raise TypeError(f"GradReducer supports float32 / complex64 parameters, got {p.dtype}")
            self._slot[i] = (off, n)
            p.grad = self._slot_view(i)
            self._bucket_of[i] = len(self.buckets)
            off += slots[i]
            cur_n += 1
            if off - cur_start >= limit:
                self.buckets.append([cur_start, off, cur_n])
                cur_start, cur_n = off, 0
        if cur_n:
            self.buckets.append([cur_start, off, cur_n])
        self._pending = [b[2] for b in self.buckets]
        self._handles = []
        self.overlap = overlap
        if overlap:
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))

    # ------------------------------------------------------------------------------------------
    def _enabled(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _slot_view(self, i) -> torch.Tensor:
        off, n = self._slot[i]
        p = self.params[i]
        view = self.flat[off : off + n]
        return torch.view_as_complex(view.view(*p.shape, 2)) if p.is_complex() else view.view(p.shape)

    def _rebind(self, i) -> None:
        """``optimizer.zero_grad()`` / ``model.zero_grad()`` default to set_to_none=True, which drops the views into the flat
        buffer; autograd then allocates a fresh ``.grad``.  Move that gradient into its slot and make ``.grad`` the view
        again, so that what is all-reduced is always what backward produced (never a stale slot)."""
        p = self.params[i]
        off, _ = self._slot[i]
        want = self.flat.data_ptr() + 4 * off
        g = p.grad
        if g is None:
            raise RuntimeError("GradReducer: a parameter has no gradient after backward (unused parameter?); "
                               "call reducer.zero_grad() before the step so that its slot holds zeros")
        if g.data_ptr() != want:
            view = self._slot_view(i)
            view.copy_(g)
            p.grad = view

    def _make_hook(self, i):
        def hook(_param):
            self._rebind(i)
            if not self._sync:
                return
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0 and self._enabled():
                s, e, _ = self.buckets[b]
                self._handles.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

        return hook

    def zero_grad(self) -> None:
        """Zero the flat buffer in place and make every param.grad a view of it again.  Use this instead of
        ``optimizer.zero_grad()``: it is one memset, and gradients then accumulate straight into the buffer that is
        all-reduced (a dropped view is detected and re-bound after backward, at the cost of a copy)."""
        self.flat.zero_()
        for i, p in enumerate(self.params):
            off, _ = self._slot[i]
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                p.grad = self._slot_view(i)
        self._pending = [b[2] for b in self.buckets]
        self._handles = []

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation: backward passes inside this context only accumulate locally; the first ``finish()``
        outside it reduces the accumulated sum ONCE (reducing after every micro-step would add already-reduced sums)."""
        old = self._sync
        self._sync = False
        try:
            yield
        finally:
            self._sync = old

    def allreduce_now(self) -> None:
        """All buckets, now, on the current stream (for steps replayed from a CUDA graph, where the hooks do not run)."""
        if not self._enabled():
            return
        hs = [dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True) for s, e, _ in self.buckets]
        for h in hs:
            h.wait()

    def finish(self) -> None:
        """Call after backward(): wait for the bucket all-reduces (or run them now if hooks are off /
        a bucket never completed, e.g. unused parameters)."""
        if not self._sync:
            return
        for i in range(len(self.params)):      # every gradient that is reduced is the one backward produced
            self._rebind(i)
        if not self._enabled():
            self._pending = [b[2] for b in self.buckets]
            return
        for b, left in enumerate(self._pending):
            if left != 0 or not self.overlap:
                s, e, _ = self.buckets[b]
                self._handles.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for h in self._handles:
            h.wait()
        self._handles = []
        self._pending = [b[2] for b in self.buckets]     # ready for the next step even if zero_grad() is not ours

    @property
    def payload_bytes(self) -> int:
        return self.flat.numel() * 4


def shard_batch(x: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """Rank r's contiguous slice [r*B/N, (r+1)*B/N) of a global batch."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    B = x.shape[0]
    if B % world:
        raise ValueError(f"global batch {B} is not divisible by world size {world}")
    per = B // world
    return x[rank * per : (rank + 1) * per]
