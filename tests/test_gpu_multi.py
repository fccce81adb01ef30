"""GPU, >= 2 devices (`gpurun --gpus 2 -- python -m pytest tests -m multigpu`): batch-sharded data parallelism over NCCL.

SURVEY.md section 4 / 8(e): a global batch B on one GPU and B/N per rank on N ranks with the SUM all-reduce of
uno_b200.parallel.GradReducer must give the same loss and the same gradient of every parameter -- eagerly (bucketed all-reduce
launched from autograd hooks, overlapping backward) and with the whole step replayed from a CUDA graph (all-reduce after the
replay).  Tolerance: 5e-5 of each gradient's largest magnitude (the summation order over the batch differs).
"""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.multigpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grads(model):
    return [(torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).detach().float().cpu().clone() for p in model.parameters()]


def _worker(rank, world, port, case, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from uno_b200 import models
    from uno_b200.graphed import GraphedStep, make_eager_step
    from uno_b200.losses import LpLoss
    from uno_b200.parallel import GradReducer, shard_batch

    torch.cuda.set_device(rank)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cls, args, kw, xshape, tshape, B, ar = case
    torch.manual_seed(0)
    model = getattr(models, cls)(*args, **kw).cuda()
    torch.manual_seed(1)
    x = torch.randn(B, *xshape)
    y = torch.randn(B, *tshape)
    out = {}
    if rank == 0:       # the whole batch on one GPU, no collective
        step = make_eager_step(model, LpLoss(size_average=False), B, tshape if not ar else None, ar)
        out["loss_full"] = float(step(x.cuda(), y.cuda()).detach())
        out["grads_full"] = _grads(model)
    dist.barrier()
    xs, ys = shard_batch(x, rank, world).cuda(), shard_batch(y, rank, world).cuda()
    Bl = xs.shape[0]
    # eager, hooks overlap the bucketed all-reduce with backward
    red = GradReducer(model, bucket_mb=1.0)
    step = make_eager_step(model, LpLoss(size_average=False), Bl, tshape if not ar else None, ar, zero=red.zero_grad, after=red.finish)
    for _ in range(2):
        loss = step(xs, ys)
    lt = loss.detach().clone()
    dist.all_reduce(lt)
    out["loss_eager"] = float(lt)
    out["grads_eager"] = _grads(model)
    # the same step replayed from a CUDA graph, all-reduce after the replay
    model2 = getattr(models, cls)(*args, **kw).cuda()
    model2.load_state_dict(model.state_dict())
    gs = GraphedStep(model2, LpLoss(size_average=False), xs, ys, ar_steps=ar, reducer=GradReducer(model2, overlap=False))
    lg = gs(xs, ys).detach().clone()
    dist.all_reduce(lg)
    out["loss_graph"] = float(lg)
    out["grads_graph"] = _grads(model2)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


CASES = {
    "darcy_small": ("UNO_9", (3, 16), dict(pad=5), (85, 85, 1), (85, 85), 4, 0),
    "ns2d_rollout": ("UNO", (14, 16), {}, (64, 64, 10), (64, 64, 3), 4, 3),
    "ns3d_small": ("Uno3D_T10", (6, 4), dict(pad=3), (32, 32, 10, 1), (32, 32, 10), 2, 0),
}


@pytest.mark.parametrize("name", list(CASES))
def test_batch_shard_over_nccl_matches_one_gpu(name):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    from uno_b200 import build

    build.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, CASES[name], q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for tag in ("eager", "graph"):
        assert abs(out[f"loss_{tag}"] - out["loss_full"]) < 1e-5 * abs(out["loss_full"]), (tag, out[f"loss_{tag}"], out["loss_full"])
        gmax = max(float(g.abs().max()) for g in out["grads_full"])
        for i, (a, b) in enumerate(zip(out[f"grads_{tag}"], out["grads_full"])):
            scale = max(float(b.abs().max()), 1e-4 * gmax)
            assert float((a - b).abs().max()) <= 5e-5 * scale, (tag, i, float((a - b).abs().max()), scale)
