"""Shared helpers of the lift / project tests: seeded inputs and the fp64 autograd oracle
(oracle/uno_torch_port.py lift / project, which the model golden fixtures pin to the reference)."""
import numpy as np
import torch

from oracle import uno_torch_port as port


def lift_inputs(case, seed=0):
    B, dims, lo, hi, raw, gch, hid, out = case
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    cin = raw + gch
    t = dict(a=f(B, *dims, raw), grid=f(*dims, gch), w_a=f(hid, cin) / np.sqrt(cin), b_a=0.3 * f(hid),
             w_b=f(out, hid) / np.sqrt(hid), b_b=0.3 * f(out))
    t["gh"] = f(B, out, *[n + l + h for n, l, h in zip(dims, lo, hi)])
    return {k: v.astype(np.float32) for k, v in t.items()}   # (float32 / np.float64 scalar promotes to float64)


def lift_oracle(case, t):
    _, _, lo, hi, *_ = case
    td = {k: torch.tensor(v, dtype=torch.float64, requires_grad=k not in ("grid", "gh")) for k, v in t.items()}
    h = port.lift(td["a"], td["grid"], td["w_a"], td["b_a"], td["w_b"], td["b_b"], lo, hi)
    h.backward(td["gh"])
    return dict(h=h.detach().numpy(), ga=td["a"].grad.numpy(), gw_a=td["w_a"].grad.numpy(), gb_a=td["b_a"].grad.numpy(),
                gw_b=td["w_b"].grad.numpy(), gb_b=td["b_b"].grad.numpy())


def project_inputs(case, seed=0):
    B, dims, lo, hi, src_ch, hid, out = case
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    full = [n + l + h for n, l, h in zip(dims, lo, hi)]
    ct = sum(src_ch)
    t = dict(srcs=[f(B, c, *full) for c in src_ch], w1=f(hid, ct) / np.sqrt(ct), b1=0.3 * f(hid), w2=f(out, hid) / np.sqrt(hid),
             b2=0.3 * f(out), gout=f(B, *dims, out))
    return {k: ([x.astype(np.float32) for x in v] if k == "srcs" else v.astype(np.float32)) for k, v in t.items()}


def project_oracle(case, t):
    _, _, lo, hi, *_ = case
    srcs = [torch.tensor(s, dtype=torch.float64, requires_grad=True) for s in t["srcs"]]
    td = {k: torch.tensor(t[k], dtype=torch.float64, requires_grad=True) for k in ("w1", "b1", "w2", "b2")}
    out = port.project(srcs, td["w1"], td["b1"], td["w2"], td["b2"], lo, hi)
    out.backward(torch.tensor(t["gout"], dtype=torch.float64))
    return dict(out=out.detach().numpy(), gsrcs=[s.grad.numpy() for s in srcs], gw1=td["w1"].grad.numpy(), gb1=td["b1"].grad.numpy(),
                gw2=td["w2"].grad.numpy(), gb2=td["b2"].grad.numpy())
