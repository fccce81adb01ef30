"""GPU: whole U-NO models on the CUDA blocks against the reference-generated golden outputs, plus the
behavioural contract of the drop-in modules."""
import numpy as np
import pytest
import torch

from conftest import BWD_TOL, FWD_TOL, rel_err
from test_oracle import MODEL_CASES, check_golden_gradients

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", list(MODEL_CASES))
def test_models_match_reference_golden(tag, golden, cuda_lib):
    from uno_b200 import models

    cls, args, kw, xshape, tshape = MODEL_CASES[tag]
    g = golden("models")
    torch.manual_seed(0)
    np.random.seed(0)
    model = getattr(models, cls)(*args, **kw)
    assert list(model.state_dict().keys()) == [str(k) for k in g[f"{tag}.keys"]]
    model = model.cuda()
    torch.manual_seed(1)
    x = torch.randn(*xshape)
    tgt = torch.randn(*tshape)
    y = model(x.cuda())
    # end-to-end tolerance: rel-L2 <= 1e-4; observed far below
    ref = g[f"{tag}.y"]
    assert np.linalg.norm(y.detach().cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-4
    assert rel_err(y.detach().cpu().numpy(), ref) < 10 * FWD_TOL
    B = xshape[0]
    tg = tgt.cuda()
    loss = torch.sum(torch.norm(y.reshape(B, -1) - tg.reshape(B, -1), 2, 1) / torch.norm(tg.reshape(B, -1), 2, 1))
    assert abs(loss.item() - float(g[f"{tag}.loss"])) < 1e-4 * abs(float(g[f"{tag}.loss"]))
    loss.backward()
    # element-wise against the reference's own gradients at the fixture's sample positions of EVERY parameter
    # (InstanceNorm and 3-D models included), relative to each gradient's largest magnitude
    check_golden_gradients(model, g, tag, 4 * BWD_TOL)


def test_model_gradients_match_cpu_port(cuda_lib):
    """Same weights on the CUDA blocks and on the CPU fp32 oracle port: every parameter gradient agrees."""
    from oracle import uno_torch_port as port
    from uno_b200 import models

    torch.manual_seed(0)
    ref = models.UNO(14, 8, ops=port)
    torch.manual_seed(0)
    ours = models.UNO(14, 8).cuda()
    ours.load_state_dict(ref.state_dict())
    torch.manual_seed(3)
    x = torch.randn(2, 64, 64, 10)
    (ref(x) ** 2).sum().backward()
    (ours(x.cuda()) ** 2).sum().backward()
    for (k, a), (_, b) in zip(ours.named_parameters(), ref.named_parameters()):
        ga = torch.view_as_real(a.grad).cpu() if a.grad.is_complex() else a.grad.cpu()
        gb = torch.view_as_real(b.grad) if b.grad.is_complex() else b.grad
        assert float((ga - gb).abs().max()) < 4 * BWD_TOL * max(float(gb.abs().max()), 1e-6), k


def test_sticky_dims_and_notebook_shapes(cuda_lib):
    from uno_b200 import integral_operators as ops

    c = ops.SpectralConv2d_Uno(2, 2, 16, 16, 4, 4).cuda()
    x = torch.randn(1, 2, 16, 16, device="cuda")
    assert c(x).shape == (1, 2, 16, 16)
    assert c(x, 24, 20).shape == (1, 2, 24, 20)
    assert c(x).shape == (1, 2, 24, 20)                       # sticky (integral_operators.py:182-184)
    # UNO_Tutorial.ipynb:266 / :420 known shapes
    blk = ops.OperatorBlock_2D(2, 4, 50, 50, 10, 10).cuda()
    assert blk(torch.randn(1, 2, 100, 100, device="cuda")).shape == (1, 4, 50, 50)
    assert blk(torch.randn(1, 2, 200, 200, device="cuda"), 100, 100).shape == (1, 4, 100, 100)
    assert blk.conv.dim1 == 100 and blk.w.dim1 == 50         # conv mutated, pointwise not
    # default modes (integral_operators.py:157-158, :331-333)
    d = ops.SpectralConv2d_Uno(2, 2, 16, 16)
    assert (d.modes1, d.modes2) == (7, 8)
    e = ops.SpectralConv3d_Uno(1, 1, 4, 4, 6)
    assert (e.modes1, e.modes2, e.modes3) == (4, 4, 4)
    # float co-dimensions are int()-cast (models pass 2*factor*width with factor=3/4)
    f = ops.OperatorBlock_2D(48.0, 96.0, 8, 8, 3, 3)
    assert f.conv.weights1.shape == (48, 96, 3, 3) and f.w.conv.weight.shape == (96, 48, 1, 1)


def test_error_contract(cuda_lib):
    from uno_b200 import integral_operators as ops

    c = ops.SpectralConv2d_Uno(2, 2, 8, 8, 3, 6).cuda()
    with pytest.raises(RuntimeError):                         # modes2 > W//2+1 (reference: einsum size error)
        c(torch.randn(1, 2, 8, 8, device="cuda"))
    c = ops.SpectralConv2d_Uno(2, 2, 16, 16, 4, 4).cuda()
    with pytest.raises(RuntimeError):                         # fp64 input (reference: ComplexDouble vs ComplexFloat)
        c(torch.randn(1, 2, 16, 16, device="cuda", dtype=torch.float64))
    with pytest.raises(RuntimeError):
        c(torch.randn(1, 2, 16, 16, device="cuda", dtype=torch.bfloat16))
    with pytest.raises(RuntimeError, match="CUDA"):           # no CPU fallback
        ops.SpectralConv2d_Uno(2, 2, 16, 16, 4, 4)(torch.randn(1, 2, 16, 16))
    with pytest.raises(ValueError):                           # pointwise_op_1D raises on torch >= 1.11 upstream too
        ops.OperatorBlock_1D(2, 2, 16, 4).cuda()(torch.randn(1, 2, 16, device="cuda"))


def test_non_contiguous_and_no_grad(cuda_lib):
    from uno_b200 import integral_operators as ops

    torch.manual_seed(0)
    blk = ops.OperatorBlock_2D(3, 5, 12, 10, 4, 4, Normalize=True).cuda()
    xc = torch.randn(2, 16, 20, 3, device="cuda")
    x_nc = xc.permute(0, 3, 1, 2)                             # NCHW view of channels-last memory
    y1 = blk(x_nc, 12, 10)
    y2 = blk(x_nc.contiguous(), 12, 10)
    assert torch.equal(y1, y2)
    with torch.no_grad():
        y3 = blk(x_nc, 12, 10)
    assert rel_err(y3.cpu().numpy(), y1.detach().cpu().numpy()) < 1e-6
    blk.eval()
    assert torch.equal(blk(x_nc, 12, 10), y1)                 # no train/eval difference (B.1)


def test_linearity_and_batch_independence_full_size(cuda_lib):
    """Size-independent properties at a BASELINE-scale level (Darcy conv0: 32->64, 481^2 -> 240^2, 18 modes)."""
    from uno_b200 import integral_operators as ops

    torch.manual_seed(0)
    c = ops.SpectralConv2d_Uno(32, 64, 240, 240, 18, 18).cuda()
    x1 = torch.randn(2, 32, 481, 481, device="cuda")
    x2 = torch.randn(2, 32, 481, 481, device="cuda")
    with torch.no_grad():
        y1, y2, y12 = c(x1), c(x2), c(2.0 * x1 - 3.0 * x2)
        assert rel_err((2.0 * y1 - 3.0 * y2).cpu().numpy(), y12.cpu().numpy()) < FWD_TOL
        ya = c(x1[:1])
        assert rel_err(ya.cpu().numpy(), y1[:1].cpu().numpy()) < 1e-6
        # a band-limited input is reproduced exactly by same-size identity weights
        ident = ops.SpectralConv2d_Uno(1, 1, 64, 64, 8, 8).cuda()
        ident.weights1.fill_(1.0)
        ident.weights2.fill_(1.0)
        n = torch.arange(64, device="cuda", dtype=torch.float32)
        img = torch.cos(2 * torch.pi * 3 * n / 64)[:, None] * torch.sin(2 * torch.pi * 5 * n / 64)[None, :] + 0.5
        out = ident(img[None, None])
        assert rel_err(out[0, 0].cpu().numpy(), img.cpu().numpy()) < FWD_TOL


def test_autoregressive_rollout_gradients_match_cpu_port(cuda_lib):
    """BASELINE config 3 as the reference trains it (ns_train_2d.py:52-67): predictions are fed back as inputs, so the
    backward pass runs through the lift kernel's input gradient at every step.  CUDA path vs the CPU fp32 oracle port."""
    import bench
    from oracle import uno_torch_port as port
    from uno_b200 import models
    from uno_b200.losses import LpLoss

    torch.manual_seed(0)
    ref = models.UNO(14, 8, ops=port)
    torch.manual_seed(0)
    ours = models.UNO(14, 8).cuda()
    ours.load_state_dict(ref.state_dict())
    torch.manual_seed(5)
    B, T = 2, 3
    x = torch.randn(B, 64, 64, 10)
    y = torch.randn(B, 64, 64, T)
    lr_ = bench.make_step(ref, port.LpLoss(size_average=False), B, (64, 64, T), ar_steps=T)(x, y)
    lo = bench.make_step(ours, LpLoss(size_average=False), B, (64, 64, T), ar_steps=T)(x.cuda(), y.cuda())
    assert abs(float(lo) - float(lr_)) < 1e-4 * abs(float(lr_))
    for (k, a), (_, b) in zip(ours.named_parameters(), ref.named_parameters()):
        ga = torch.view_as_real(a.grad).cpu() if a.grad.is_complex() else a.grad.cpu()
        gb = torch.view_as_real(b.grad) if b.grad.is_complex() else b.grad
        assert float((ga - gb).abs().max()) < 8 * BWD_TOL * max(float(gb.abs().max()), 1e-6), k


def test_training_step_is_cuda_graph_capturable(cuda_lib):
    """everything the library enqueues goes to the caller's stream and nothing synchronises or allocates
    after the first call of a shape, so a whole forward + loss + backward (here also the fused Adam step) can be captured
    in a CUDA graph and replayed on new data.  Replay must reproduce the eager results."""
    from uno_b200 import models
    from uno_b200.losses import LpLoss
    from uno_b200.optim import Adam

    torch.manual_seed(0)
    model = models.UNO_9(3, 8, pad=5).cuda()
    ref = models.UNO_9(3, 8, pad=5).cuda()
    ref.load_state_dict(model.state_dict())
    loss_fn = LpLoss(size_average=False)
    B, S = 2, 85
    xs = torch.randn(3, B, S, S, 1, device="cuda")
    ys = torch.randn(3, B, S, S, device="cuda")
    opt = Adam(model.parameters(), lr=1e-3)
    opt_ref = Adam(ref.parameters(), lr=1e-3)
    static_x, static_y = xs[0].clone(), ys[0].clone()

    def step(m, o, x, y):
        o.zero_grad(set_to_none=False)
        loss = loss_fn(m(x).reshape(B, -1), y.reshape(B, -1))
        loss.backward()
        o.step()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):           # warm-up on the capture stream: plans, scratch, optimiser state, .grad buffers
        for _ in range(2):
            step(model, opt, static_x, static_y)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(2):
        step(ref, opt_ref, xs[0], ys[0])
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        static_loss = step(model, opt, static_x, static_y)
    # NOTE: the step counter of the optimiser is host state baked into the capture, so only the first replay matches an
    # eager step bit for bit in its bias correction; compare that one
    static_x.copy_(xs[1])
    static_y.copy_(ys[1])
    graph.replay()
    torch.cuda.synchronize()
    eager_loss = step(ref, opt_ref, xs[1], ys[1])
    assert abs(float(static_loss) - float(eager_loss)) < 1e-5 * abs(float(eager_loss))
    for (k, a), (_, b) in zip(model.named_parameters(), ref.named_parameters()):
        ga = torch.view_as_real(a.grad) if a.grad.is_complex() else a.grad
        gb = torch.view_as_real(b.grad) if b.grad.is_complex() else b.grad
        # (same arithmetic, but several reductions finish with atomics -- the weight gradients, the split contraction -- so the
        # summation order differs from run to run: the bound is the backward tolerance, not bit equality)
        assert float((ga - gb).abs().max()) <= BWD_TOL * max(float(gb.abs().max()), 1e-6), k
        # Adam turns a gradient into a step of size ~lr whatever its magnitude, so elements whose gradient is at the
        # round-off level (their value depends on the order of the atomics) may legitimately differ by a fraction of lr;
        # the update is compared where the gradient is well above that level, and bounded by lr everywhere
        pa = torch.view_as_real(a.detach()) if a.is_complex() else a.detach()
        pb = torch.view_as_real(b.detach()) if b.is_complex() else b.detach()
        mag = (gb[..., 0] ** 2 + gb[..., 1] ** 2).sqrt().unsqueeze(-1).expand_as(gb) if a.is_complex() else gb.abs()
        solid = mag > 1e-3 * float(mag.max())
        assert float((pa - pb).abs().max()) <= 1e-3, k
        if bool(solid.any()):
            assert float((pa - pb)[solid].abs().max()) <= 5e-6, k


@pytest.mark.parametrize("ar_steps", [0, 3])
def test_graphed_step_reproduces_the_eager_step(ar_steps, cuda_lib):
    """uno_b200/graphed.py: zero_grad + forward (or the autoregressive rollout of ns_train_2d.py:52-67) + loss + backward
    captured in one CUDA graph; a replay on NEW data must give the eager step's loss and every parameter gradient."""
    from uno_b200 import models
    from uno_b200.graphed import GraphedStep, make_eager_step
    from uno_b200.losses import LpLoss

    torch.manual_seed(0)
    model = models.UNO(14, 8).cuda()
    ref = models.UNO(14, 8).cuda()
    ref.load_state_dict(model.state_dict())
    B, T = 2, max(ar_steps, 1)
    torch.manual_seed(4)
    xs = torch.randn(2, B, 64, 64, 10, device="cuda")
    ys = torch.randn(2, B, 64, 64, T, device="cuda") if ar_steps else torch.randn(2, B, 64, 64, device="cuda")
    step = GraphedStep(model, LpLoss(size_average=False), xs[0], ys[0], ar_steps=ar_steps)
    loss_g = float(step(xs[1], ys[1]))          # replay on data the capture never saw
    eager = make_eager_step(ref, LpLoss(size_average=False), B, tuple(ys.shape[2:]), ar_steps)
    loss_e = float(eager(xs[1], ys[1]))
    assert abs(loss_g - loss_e) < 1e-5 * abs(loss_e)
    for (k, a), (_, b) in zip(model.named_parameters(), ref.named_parameters()):
        ga = torch.view_as_real(a.grad) if a.grad.is_complex() else a.grad
        gb = torch.view_as_real(b.grad) if b.grad.is_complex() else b.grad
        assert float((ga - gb).abs().max()) <= 2e-5 * max(float(gb.abs().max()), 1e-6), k
    # a second replay must not accumulate onto the first (zero_grad is inside the graph)
    loss_g2 = float(step(xs[1], ys[1]))
    assert abs(loss_g2 - loss_g) < 1e-6 * abs(loss_g)
    for (k, a), (_, b) in zip(model.named_parameters(), ref.named_parameters()):
        ga = torch.view_as_real(a.grad) if a.grad.is_complex() else a.grad
        gb = torch.view_as_real(b.grad) if b.grad.is_complex() else b.grad
        assert float((ga - gb).abs().max()) <= 2e-5 * max(float(gb.abs().max()), 1e-6), k
    # the same object's eager path gives the same numbers (bench.py profiles through it)
    assert abs(float(step.eager_step(xs[1], ys[1])) - loss_e) < 1e-5 * abs(loss_e)


def test_parameters_on_another_device_are_rejected(cuda_lib):
    """A module that was never moved to the input's device must raise (the reference's torch ops do), not hand a host pointer
    to a kernel; double backward raises instead of silently cutting the graph."""
    from uno_b200 import integral_operators as ops

    blk = ops.OperatorBlock_2D(2, 3, 8, 8, 3, 3)               # parameters on the CPU
    with pytest.raises(RuntimeError):
        blk(torch.randn(1, 2, 8, 8, device="cuda"), 8, 8)
    blk = blk.cuda()
    x = torch.randn(1, 2, 8, 8, device="cuda", requires_grad=True)
    y = blk(x, 8, 8)
    (g,) = torch.autograd.grad(y.sum(), x, create_graph=True)
    with pytest.raises(RuntimeError):
        g.sum().backward()
