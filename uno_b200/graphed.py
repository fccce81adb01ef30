"""A whole training step -- zero_grad, forward (or the 10-call autoregressive rollout of ns_train_2d.py:52-67), loss, backward --
captured ONCE in a CUDA graph and replayed on static buffers (SURVEY.md 8(f) row 2, 8(b) "CUDA-graph-capturable").

The library enqueues everything on the caller's stream, never synchronises and allocates nothing of its own after the first
call of a shape (plan constants and operand images are built during the warm-up steps), so the step is capturable as is.
Replay removes the per-call host work (ctypes marshalling, allocator, autograd bookkeeping: ~20 C-ABI calls per model call),
which is what bounds the small-batch regimes -- the rollout at batch 64 / N GPUs, strong-scaled shards -- not the kernels.

Gradients live in one flat buffer (``GradReducer``): zeroing them is one memset inside the graph, and with several ranks the
bucketed all-reduce runs right after the replay on the same stream.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .parallel import GradReducer


def make_eager_step(model, loss_fn, B: int, tshape, ar_steps: int = 0, zero: Optional[Callable] = None, after: Optional[Callable] = None):
    """The reference's training step without the optimiser (train_darcy.py:50-54, ns_train_3d.py:51-65; ns_train_2d.py:52-67
    when ``ar_steps`` > 0: every prediction is fed back as the newest input frame and the per-step losses are summed)."""

    def step(x, y):
        if zero is not None:
            zero()
        else:
            model.zero_grad(set_to_none=True)
        if ar_steps:
            loss, xx = 0, x
            for t in range(ar_steps):
                im = model(xx)
                loss = loss + loss_fn(im.reshape(B, -1), y[..., t : t + 1].reshape(B, -1))
                xx = torch.cat((xx[..., 1:], im), dim=-1)
        else:
            out = model(x).reshape(B, *tshape)
            loss = loss_fn(out.reshape(B, -1), y.reshape(B, -1))
        loss.backward()
        if after is not None:
            after()
        return loss

    return step


class GraphedStep:
    """``step = GraphedStep(model, loss_fn, x_example, y_example, ar_steps=10); loss = step(x, y)``.

    ``loss`` is a static 0-dim tensor that every replay overwrites; ``param.grad`` are views of ``step.reducer.flat``.
    ``step.eager_step`` runs the same step without the graph (same buffers, same result)."""

    mode = "graph"

    def __init__(self, model, loss_fn, x_example: torch.Tensor, y_example: torch.Tensor, ar_steps: int = 0,
                 reducer: Optional[GradReducer] = None, warmup: int = 2):
        if not x_example.is_cuda:
            raise RuntimeError("GraphedStep needs CUDA tensors (there is no CPU path)")
        self.model, self.loss_fn = model, loss_fn
        self.B = int(x_example.shape[0])
        tshape = tuple(y_example.shape[1:])
        # hooks off: inside a capture nothing may talk to NCCL behind the graph's back; the reduction runs after the replay
        self.reducer = reducer if reducer is not None else GradReducer(model, overlap=False)
        if self.reducer.overlap:
            raise ValueError("GraphedStep needs a GradReducer built with overlap=False")
        self.static_x = x_example.clone()
        self.static_y = y_example.clone()
        self._body = make_eager_step(model, loss_fn, self.B, tshape, ar_steps, zero=self.reducer.zero_grad)
        self.stream = torch.cuda.Stream(device=x_example.device)
        self.stream.wait_stream(torch.cuda.current_stream(x_example.device))
        with torch.cuda.stream(self.stream):      # warm-up on the capture stream: plans, operand images, .grad views
            for _ in range(max(warmup, 1)):
                self._body(self.static_x, self.static_y)
        torch.cuda.current_stream(x_example.device).wait_stream(self.stream)
        torch.cuda.synchronize(x_example.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.static_loss = self._body(self.static_x, self.static_y)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        self.reducer.allreduce_now()
        return self.static_loss

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if x.data_ptr() != self.static_x.data_ptr():
            self.static_x.copy_(x, non_blocking=True)
        if y.data_ptr() != self.static_y.data_ptr():
            self.static_y.copy_(y, non_blocking=True)
        return self.replay()

    def run_from_host(self, x_host: torch.Tensor, y_host: torch.Tensor) -> torch.Tensor:
        """Pinned host buffers straight into the graph's static inputs (no staging copy), then replay."""
        self.static_x.copy_(x_host, non_blocking=True)
        self.static_y.copy_(y_host, non_blocking=True)
        return self.replay()

    def eager_step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """The same step without the graph, on the stream the capture used (autograd's AccumulateGrad nodes remember the stream
        they were created on; running them from another stream makes the engine insert synchronisation and warn)."""
        cur = torch.cuda.current_stream(x.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            loss = self._body(x, y)
        cur.wait_stream(self.stream)
        self.reducer.allreduce_now()
        return loss
