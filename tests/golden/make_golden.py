"""Generate the golden fixtures in tests/golden/ from the REAL reference.

Run in the build container only (it imports /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Everything is CPU fp32 torch 2.11 with fixed seeds.  Small cases store inputs, parameters, outputs
and gradients in full; the larger ones (config 1 of BASELINE.json, whole models) are regenerated
from the seed at test time and the fixture stores outputs / sub-samples / parameter checksums.
"""
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
for name in ("matplotlib", "matplotlib.pyplot"):  # the model files import it, unused on this path
    sys.modules.setdefault(name, types.ModuleType(name))

import integral_operators as ref  # noqa: E402

sys.path.insert(0, os.path.dirname(HERE))
from cases import grad_sample_indices  # noqa: E402

torch.set_num_threads(4)


def n(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(arrays)} arrays")


# ------------------------------------------------------------------------------------------------
SPECTRAL_CASES = {
    # name: (B, Ci, Co, in_dims, out_dims, modes)
    "s1d": (2, 3, 4, (32,), (24,), (5,)),
    "s2d_same": (2, 3, 5, (24, 24), (24, 24), (7, 8)),
    "s2d_down_odd": (2, 3, 4, (37, 41), (19, 23), (5, 6)),
    "s2d_up_odd": (2, 2, 3, (8, 8), (32, 31), (3, 3)),
    "s2d_overlap_out": (1, 4, 2, (16, 16), (8, 8), (6, 4)),
    "s2d_overlap_in": (1, 2, 3, (10, 10), (12, 12), (6, 5)),
    "s3d_down": (1, 2, 3, (16, 16, 13), (12, 12, 13), (4, 4, 5)),
    "s3d_up": (1, 2, 2, (8, 8, 20), (16, 16, 31), (3, 3, 7)),
    "s3d_overlap": (1, 2, 2, (16, 16, 20), (8, 8, 20), (6, 6, 7)),
}


def make_spectral():
    out = {}
    for name, (B, Ci, Co, idim, odim, modes) in SPECTRAL_CASES.items():
        torch.manual_seed(100 + len(out))
        cls = {1: ref.SpectralConv1d_Uno, 2: ref.SpectralConv2d_Uno, 3: ref.SpectralConv3d_Uno}[len(idim)]
        m = cls(Ci, Co, *odim, *modes)
        x = torch.randn(B, Ci, *idim, requires_grad=True)
        y = m(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        out[f"{name}.x"] = n(x)
        out[f"{name}.y"] = n(y)
        out[f"{name}.gy"] = n(gy)
        out[f"{name}.gx"] = n(x.grad)
        for i in range(2 ** (len(idim) - 1)):
            w = getattr(m, f"weights{i + 1}")
            out[f"{name}.w{i + 1}"] = n(w)
            out[f"{name}.gw{i + 1}"] = n(w.grad)
    save("spectral.npz", **out)


POINTWISE_CASES = {
    "p2d_down": (2, 3, 4, (16, 16), (8, 8)),
    "p2d_up": (2, 5, 2, (8, 8), (16, 16)),
    "p2d_odd": (1, 2, 3, (37, 41), (19, 23)),
    "p2d_same": (1, 3, 3, (12, 12), (12, 12)),
    "p2d_mixed": (1, 4, 2, (12, 14), (30, 9)),
    "p3d_down": (1, 2, 3, (16, 16, 13), (12, 12, 13)),
    "p3d_up": (1, 3, 2, (8, 8, 20), (16, 16, 31)),
    "p3d_same": (1, 2, 2, (12, 12, 10), (12, 12, 10)),
}


def make_pointwise():
    out = {}
    for name, (B, Ci, Co, idim, odim) in POINTWISE_CASES.items():
        torch.manual_seed(200 + len(out))
        cls = {2: ref.pointwise_op_2D, 3: ref.pointwise_op_3D}[len(idim)]
        m = cls(Ci, Co, *odim)
        x = torch.randn(B, Ci, *idim, requires_grad=True)
        y = m(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        out.update({
            f"{name}.x": n(x), f"{name}.y": n(y), f"{name}.gy": n(gy), f"{name}.gx": n(x.grad),
            f"{name}.cw": n(m.conv.weight), f"{name}.cb": n(m.conv.bias),
            f"{name}.gcw": n(m.conv.weight.grad), f"{name}.gcb": n(m.conv.bias.grad),
        })
    # resample matrices of the Darcy levels (identity pushed through F.interpolate), fp32
    for a, b in ((481, 240), (240, 120), (120, 240), (240, 481), (446, 223), (223, 111), (111, 223), (223, 446), (64, 48), (48, 32), (32, 16), (16, 32), (32, 48), (48, 64)):
        eye = torch.eye(a).reshape(1, 1, a, a)
        R = torch.nn.functional.interpolate(eye, size=(b, a), mode="bicubic", align_corners=True, antialias=True)[0, 0]
        # store the band only: first non-zero column and the non-zero run of every row
        Rn = n(R)
        taps = int((np.abs(Rn) > 0).sum(1).max())
        start = np.array([int(np.argmax(np.abs(r) > 0)) for r in Rn], np.int32)
        band = np.zeros((b, taps), np.float32)
        for i in range(b):
            seg = Rn[i, start[i] : start[i] + taps]
            band[i, : len(seg)] = seg
        out[f"R_{a}_{b}.start"] = start
        out[f"R_{a}_{b}.band"] = band
    save("pointwise.npz", **out)


BLOCK_CASES = {
    # name: (B, Ci, Co, in, out, modes, Normalize, Non_Lin)
    "b2d_down_norm": (2, 3, 4, (20, 24), (10, 12), (4, 4), True, True),
    "b2d_down": (2, 3, 4, (20, 24), (10, 12), (4, 4), False, True),
    "b2d_up": (2, 4, 2, (10, 12), (20, 24), (4, 4), False, True),
    "b2d_up_norm_lin": (2, 4, 2, (10, 12), (20, 24), (4, 4), True, False),
    "b2d_same_lin": (2, 4, 3, (10, 12), (10, 12), (4, 4), False, False),
    "b3d_down_norm": (1, 2, 3, (12, 12, 10), (8, 8, 10), (3, 3, 4), True, True),
    "b3d_up": (1, 3, 2, (8, 8, 10), (12, 12, 10), (3, 3, 4), False, True),
}


def make_blocks():
    out = {}
    for name, (B, Ci, Co, idim, odim, modes, norm, nl) in BLOCK_CASES.items():
        torch.manual_seed(300 + len(out))
        cls = {2: ref.OperatorBlock_2D, 3: ref.OperatorBlock_3D}[len(idim)]
        m = cls(Ci, Co, *odim, *modes, Normalize=norm, Non_Lin=nl)
        if norm:
            with torch.no_grad():
                m.normalize_layer.weight.uniform_(0.5, 1.5)
                m.normalize_layer.bias.uniform_(-0.5, 0.5)
        x = torch.randn(B, Ci, *idim, requires_grad=True)
        y = m(x, *odim)
        gy = torch.randn_like(y)
        y.backward(gy)
        out.update({f"{name}.x": n(x), f"{name}.y": n(y), f"{name}.gy": n(gy), f"{name}.gx": n(x.grad)})
        for k, p in m.named_parameters():
            out[f"{name}.param.{k}"] = n(p)
            out[f"{name}.grad.{k}"] = n(p.grad)
    save("blocks.npz", **out)


def make_config1():
    """BASELINE.json configs[0]: SpectralConv2d single layer, batch 2, 64x64, 20 modes, 32->32."""
    torch.manual_seed(0)
    m = ref.SpectralConv2d_Uno(32, 32, 64, 64, 20, 20)
    x = torch.randn(2, 32, 64, 64, requires_grad=True)
    y = m(x)
    torch.manual_seed(1)
    gy = torch.randn_like(y)
    y.backward(gy)
    save(
        "config1.npz",
        y_sum=np.float64(y.double().sum().item()),
        y_head=n(y[0, 0, 0, :3]),
        y_sub=n(y[:, ::4, ::4, ::4]),
        gx_sub=n(x.grad[:, ::4, ::4, ::4]),
        gw1_sub=n(m.weights1.grad[::4, ::4, ::2, ::2]),
        gw2_sub=n(m.weights2.grad[::4, ::4, ::2, ::2]),
        w1_abs_sum=np.float64(m.weights1.detach().abs().double().sum().item()),
        x_abs_sum=np.float64(x.detach().abs().double().sum().item()),
    )


def _state_fingerprint(model):
    return np.array([float(torch.view_as_real(v).double().abs().sum()) if v.is_complex() else float(v.double().abs().sum()) for v in model.state_dict().values()])


def make_models():
    import darcy_flow_uno2d as d2
    import navier_stokes_uno2d as n2
    import navier_stokes_uno3d as n3

    out = {}

    def run(tag, ctor, xshape, reshape):
        torch.manual_seed(0)
        np.random.seed(0)
        model = ctor()
        torch.manual_seed(1)
        x = torch.randn(*xshape)
        y = model(x)
        tgt = torch.randn(*reshape)
        # rel-L2 loss summed over the batch, as utilities3.LpLoss(size_average=False) (utilities3.py:86-100)
        B = xshape[0]
        diff = torch.norm(y.reshape(B, -1) - tgt.reshape(B, -1), 2, 1)
        loss = torch.sum(diff / torch.norm(tgt.reshape(B, -1), 2, 1))
        loss.backward()
        out[f"{tag}.y"] = n(y)
        out[f"{tag}.loss"] = np.float64(loss.item())
        out[f"{tag}.state_fp"] = _state_fingerprint(model)
        out[f"{tag}.keys"] = np.array(list(model.state_dict().keys()))
        out[f"{tag}.grad_fp"] = np.array([float(torch.view_as_real(p.grad).double().abs().sum()) if p.grad.is_complex() else float(p.grad.double().abs().sum()) for p in model.parameters()])
        # element-wise gradients at fixed sample positions of every parameter (tests/cases.py grad_sample_indices) and their scale
        vals, off, gmax = [], [0], []
        for i, p in enumerate(model.parameters()):
            g = (torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).reshape(-1)
            idx = grad_sample_indices(i, g.numel())
            vals.append(n(g)[idx])
            off.append(off[-1] + len(idx))
            gmax.append(float(g.abs().max()))
        out[f"{tag}.grad_sub"] = np.concatenate(vals).astype(np.float32)
        out[f"{tag}.grad_sub_off"] = np.array(off, dtype=np.int64)
        out[f"{tag}.grad_max"] = np.array(gmax, dtype=np.float64)

    run("uno9_pad5", lambda: d2.UNO_9(3, 8, pad=5), (1, 85, 85, 1), (1, 85, 85))
    run("uno_ns2d", lambda: n2.UNO(14, 8), (1, 64, 64, 10), (1, 64, 64))
    run("uno_p_ns2d", lambda: n2.UNO_P(14, 8), (1, 64, 64, 10), (1, 64, 64))
    run("uno3d_t10", lambda: n3.Uno3D_T10(6, 4, pad=3), (1, 64, 64, 10, 1), (1, 64, 64, 10))
    save("models.npz", **out)


def make_training():
    """The reference's optimiser (Adam.py, complex-aware second moment |g|^2) and loss (utilities3.LpLoss) on a mixed
    set of real and complex tensors: parameters, seeded gradients, states after every one of 5 steps."""
    import Adam as ref_adam
    import utilities3 as ref_util

    out = {}
    for tag, kw in (("plain", dict(lr=1e-2)), ("wd", dict(lr=3e-3, weight_decay=1e-2, betas=(0.8, 0.95), eps=1e-6)),
                    ("amsgrad", dict(lr=1e-2, amsgrad=True))):
        torch.manual_seed(7)
        shapes = [((5, 3), False), ((17,), False), ((2, 3, 4, 4), True), ((1,), False), ((3, 2, 5), True)]
        if kw.get("amsgrad"):
            shapes = [s for s in shapes if not s[1]]        # torch.maximum has no complex kernel: upstream raises
        params = [torch.nn.Parameter(torch.randn(*sh, dtype=torch.cfloat if cx else torch.float)) for sh, cx in shapes]
        opt = ref_adam.Adam(params, **kw)
        out[f"{tag}.n"] = np.array(len(params))
        for i, q in enumerate(params):
            out[f"{tag}.p0.{i}"] = n(q).copy()          # (numpy() aliases the parameter, which is updated in place)
        for step in range(5):
            for i, q in enumerate(params):
                q.grad = torch.randn_like(q) * (0.5 + step)
                out[f"{tag}.g{step}.{i}"] = n(q.grad)
            opt.step()
            for i, q in enumerate(params):
                out[f"{tag}.p{step + 1}.{i}"] = n(q).copy()
        for i, q in enumerate(params):
            st = opt.state[q]
            out[f"{tag}.m.{i}"] = n(st["exp_avg"])
            out[f"{tag}.v.{i}"] = n(st["exp_avg_sq"])
    torch.manual_seed(11)
    x = torch.randn(4, 9, 7, requires_grad=True)
    y = torch.randn(4, 9, 7)
    for tag, kw in (("sum", dict(size_average=False)), ("mean", dict(size_average=True)), ("none", dict(reduction=False))):
        loss = ref_util.LpLoss(**kw)(x, y)
        gl = torch.randn_like(loss)
        (gx,) = torch.autograd.grad(loss, x, gl)
        out[f"loss.{tag}"] = n(loss)
        out[f"loss.{tag}.gl"] = n(gl)
        out[f"loss.{tag}.gx"] = n(gx)
    out["loss.x"] = n(x)
    out["loss.y"] = n(y)
    save("training.npz", **out)


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, help="regenerate a single fixture (e.g. training)")
    only = ap.parse_args().only
    if only:
        globals()["make_" + only]()
        sys.exit(0)
    make_training()
    make_spectral()
    make_pointwise()
    make_blocks()
    make_config1()
    make_models()
