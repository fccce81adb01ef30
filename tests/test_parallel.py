"""CPU, gloo, world_size 2: the flat-buffer gradient all-reduce reproduces single-process gradients
of a SUM-reduced loss over the global batch, including complex parameters."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.lin = torch.nn.Linear(6, 5)
        self.w = torch.nn.Parameter(torch.randn(5, 4, 3, dtype=torch.cfloat))
        self.out = torch.nn.Linear(4, 1)

    def forward(self, x):
        h = torch.nn.functional.gelu(self.lin(x))                       # [B, 5]
        spec = torch.einsum("bi,iok->bok", h.to(torch.cfloat), self.w)  # [B, 4, 3]
        r = torch.fft.irfft(spec, n=4)                                  # [B, 4, 4]
        return self.out(r.transpose(1, 2)).sum(dim=(1, 2))             # [B]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, overlap, q):
    sys.path.insert(0, ROOT)
    from uno_b200.parallel import GradReducer, shard_batch

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(1)
    x = torch.randn(8, 6)
    model = Toy()
    red = GradReducer(model, bucket_mb=0.0001, overlap=overlap)
    for _ in range(2):  # second step checks zero_grad keeps the views alive
        red.zero_grad()
        loss = model(shard_batch(x, rank, world)).sum()
        loss.backward()
        red.finish()
    # the reference's training loops call optimizer.zero_grad(), whose default (set_to_none=True) drops the views into the
    # flat buffer: the reducer must notice, move the fresh gradients into their slots and still reduce the right values
    model.zero_grad()
    assert all(p.grad is None for p in model.parameters())
    model(shard_batch(x, rank, world)).sum().backward()
    red.finish()
    assert all(p.grad.data_ptr() >= red.flat.data_ptr() for p in model.parameters())
    # gradient accumulation: two half micro-batches under no_sync + one reduction == the full local batch
    full = [torch.view_as_real(p.grad).clone() if p.grad.is_complex() else p.grad.clone() for p in model.parameters()]
    red.zero_grad()
    xs = shard_batch(x, rank, world)
    with red.no_sync():
        model(xs[: xs.shape[0] // 2]).sum().backward()
        red.finish()
    model(xs[xs.shape[0] // 2 :]).sum().backward()
    red.finish()
    for p, f in zip(model.parameters(), full):
        g = torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad
        assert float((g - f).abs().max()) < 1e-5 * max(float(f.abs().max()), 1e-3)
    grads = [torch.view_as_real(p.grad).clone() if p.grad.is_complex() else p.grad.clone() for p in model.parameters()]
    if rank == 0:
        q.put([g.numpy() for g in grads])
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_two_rank_allreduce_matches_single_process(overlap):
    torch.manual_seed(1)
    x = torch.randn(8, 6)
    ref = Toy()
    ref(x).sum().backward()
    want = [torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad for p in ref.parameters()]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for g, w in zip(got, want):
        assert abs(torch.tensor(g) - w).max() < 1e-5 * max(float(w.abs().max()), 1e-3)


def test_single_process_is_a_noop():
    from uno_b200.parallel import GradReducer

    m = Toy()
    red = GradReducer(m)
    red.zero_grad()
    m(torch.randn(3, 6)).sum().backward()
    red.finish()
    assert all(p.grad is not None and p.grad.data_ptr() >= red.flat.data_ptr() for p in m.parameters())
    assert red.payload_bytes >= sum(p.numel() * (8 if p.is_complex() else 4) for p in m.parameters())
    assert all((p.grad.data_ptr() - red.flat.data_ptr()) % 16 == 0 for p in m.parameters())
