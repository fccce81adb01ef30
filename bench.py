#!/usr/bin/env python
"""Benchmark of the U-NO hot path on B200 (contract: see the task prompt / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload darcy|ns2d|ns3d|ns2d_ar]

Headline workload = BASELINE.json configs[1]: UNO_9(3, 32, pad=12) (darcy_flow_main.py:95) on synthetic
421x421 Darcy inputs, batch 32 per GPU, forward + rel-L2 loss + backward (train_darcy.py:50-54), no
optimizer step.  One JSON line is printed by rank 0.

  value        samples/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same step driven from pinned HOST buffers (H2D of x,y and D2H of the loss inside the timed region)
  roofline     dominant kernel family of the step: algorithmic bytes / CUDA-event time vs measured HBM peak
  secondary    the other half of BASELINE's metric in the same invocation, at every --gpus N: configs[3] (Uno3D_T10 64^3,
               batch 8) and configs[2] (UNO 10-step autoregressive rollout + BPTT, batch 64) -- value, e2e, roofline each;
               at N > 1 both the weak (per-GPU batch fixed) and the strong (global batch fixed) split
  strong       (N > 1) the headline workload with the GLOBAL batch fixed at the BASELINE batch
  reference_gpu (N = 1) the reference's OWN model files (oracle/_ref, staged byte for byte by build()) run through
               torch-CUDA (cuFFT / cuBLAS / cuDNN, TF32 off) on the same B200 at the same batch -- the real bar
  sweep        (N = 1) BASELINE configs[4]: SpectralConv2d S x m x C grid, GB/s and fraction of HBM, vs stock cuFFT path
  cpu_baseline / --impl reference: the reference's own model (kind "reference") on the host cores, bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORKLOADS = {
    # name: (model class, ctor args, ctor kwargs, input shape w/o batch, target shape w/o batch, default batch, cpu sample batch)
    "darcy": ("UNO_9", (3, 32), dict(pad=12), (421, 421, 1), (421, 421), 32, 2),
    "ns2d": ("UNO", (14, 32), {}, (64, 64, 10), (64, 64), 64, 8),
    "ns3d": ("Uno3D_T10", (6, 8), dict(pad=3), (64, 64, 64, 1), (64, 64, 64), 8, 1),
    # BASELINE.json configs[2] exactly as ns_train_2d.py:46-67 trains it: 10 autoregressive model calls, summed loss, ONE backward
    "ns2d_ar": ("UNO", (14, 32), {}, (64, 64, 10), (64, 64, 10), 64, 4),
}
AR_STEPS = {"ns2d_ar": 10}
WORKLOAD_DESC = {
    "darcy": "UNO_9(3,32,pad=12) Darcy 421x421 fwd+loss+bwd",
    "ns2d": "UNO(14,32) Navier-Stokes 64x64x10 single-call fwd+loss+bwd",
    "ns3d": "Uno3D_T10(6,8,pad=3) Navier-Stokes 64x64x64 fwd+loss+bwd",
    "ns2d_ar": "UNO(14,32) Navier-Stokes 64x64, 10-step autoregressive rollout + BPTT (ns_train_2d.py:46-67)",
}
SECONDARY = ("ns3d", "ns2d_ar")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return json.load(open(path))
    except Exception:
        return {}


def source_sha16():
    """Identity of the device code a traffic capture belongs to: sha256 over the kernel sources."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "uno_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".cpp", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(role, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `role` from the committed ncu capture of one step of this
    workload (profiles/r02_traffic_<workload>.json, tools/ncu_traffic.py) -- only when that capture was taken on the device
    code that is running now (its `src_sha16` matches); otherwise None: a stale capture says nothing about this build."""
    path = os.path.join(ROOT, "profiles", f"r02_traffic_{workload}.json")
    try:
        doc = json.load(open(path))
        if doc.get("src_sha16") != source_sha16():
            return None
        return float(doc["roles"][role]["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while self.ok and not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.05)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def build_model(workload, ops=None, device="cuda"):
    from uno_b200 import models

    cls, args, kw, *_ = WORKLOADS[workload]
    torch.manual_seed(0)
    kw = dict(kw)
    if ops is not None:
        kw["ops"] = ops
    return getattr(models, cls)(*args, **kw).to(device)


def make_step(model, loss_fn, B, tshape, reducer=None, ar_steps=0, zero_grad=None):
    """One training step without the optimizer: zero_grad -> forward -> loss -> backward, exactly the reference's loops
    (train_darcy.py:50-54, ns_train_3d.py:51-65; ns_train_2d.py:52-67 for the autoregressive rollout)."""

    def step(x, y, before_loss=None):
        """`before_loss`: called once after the first forward, before the target is first read (the end-to-end arm waits there
        for the target's host-to-device copy, which it lets overlap the forward)."""
        if reducer is not None:
            reducer.zero_grad()
        elif zero_grad is not None:
            zero_grad()
        else:
            model.zero_grad(set_to_none=True)
        if ar_steps:
            # ns_train_2d.py:52-67: feed each prediction back as the newest input frame, sum the per-step losses
            loss, xx = 0, x
            for t in range(ar_steps):
                im = model(xx)
                if t == 0 and before_loss is not None:
                    before_loss()
                loss = loss + loss_fn(im.reshape(B, -1), y[..., t : t + 1].reshape(B, -1))
                xx = torch.cat((xx[..., 1:], im), dim=-1)
        else:
            out = model(x).reshape(B, *tshape)
            if before_loss is not None:
                before_loss()
            loss = loss_fn(out.reshape(B, -1), y.reshape(B, -1))
        loss.backward()
        if reducer is not None:
            reducer.finish()
        return loss

    return step


# ----------------------------------------------------------------------------------------------------------------------
# the reference itself (oracle/_ref: its own model files, unmodified) -- CPU arm and torch-CUDA leg
# ----------------------------------------------------------------------------------------------------------------------
def reference_model(workload, device):
    """(model, loss class, kind): the reference's own nn.Module when oracle/_ref (or /root/reference) is there, else the
    functional port of its operators under the same model table (kind "port")."""
    cls, args, kw, *_ = WORKLOADS[workload]
    try:
        from oracle import ref_loader

        torch.manual_seed(0)
        model = ref_loader.model_class(workload)(*args, **kw).to(device)
        return model, ref_loader.lp_loss(), "reference"
    except ImportError:
        from oracle import uno_torch_port as port

        return build_model(workload, ops=port, device=device), port.LpLoss, "port"


def cpu_reference_run(workload, steps, warmup, batch=None, budget_s=60.0):
    """The reference's own model on the host cores: fwd + loss + bwd on a bounded sample of the workload (the per-GPU batch
    when one step of it fits the time budget, else the largest power-of-two fraction that does)."""
    _, _, _, xshape, tshape, full_b, cpu_b = WORKLOADS[workload]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model, LpLoss, kind = reference_model(workload, "cpu")
    ar = AR_STEPS.get(workload, 0)

    def run(B, n):
        torch.manual_seed(1)
        x, y = torch.randn(B, *xshape), torch.randn(B, *tshape)
        step = make_step(model, LpLoss(size_average=False), B, tshape, ar_steps=ar)
        t0 = time.perf_counter()
        for _ in range(n):
            step(x, y)
        return (time.perf_counter() - t0) / n

    B = batch or cpu_b
    t_small = run(B, 1)                        # also the warm-up (MKL plans, allocator)
    if batch is None:
        want = full_b
        try:
            import psutil

            mem_cap = psutil.virtual_memory().available * 0.25
        except Exception:
            mem_cap = 16e9
        per_sample = 1.0e9      # autograd-saved activations per sample: upper bound (measured 0.5 GB Darcy, < 1 GB NS-3D)
        while want > B and (t_small * want / B * (steps + max(warmup - 1, 0)) > budget_s or want * per_sample > mem_cap):
            want //= 2
        B = max(B, want)
    for _ in range(max(warmup - 1, 0)):
        run(B, 1)
    dt = run(B, steps)
    src = "oracle/_ref: the reference's own model files" if kind == "reference" else "oracle/uno_torch_port.py"
    return {"value": B / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"batch {B} of the same workload (full per-GPU batch {full_b}), {steps} steps after {max(warmup, 1)} warm-up, "
                      f"fp32 torch CPU ({src})",
            "ms_per_step": dt * 1e3, "batch": B}


def reference_gpu_run(workload, B, steps, warmup, dev):
    """The reference's own model through torch-CUDA on this GPU (cuFFT + cuBLAS cgemm + cuDNN / ATen kernels; TF32 off so
    that it computes what its CPU path computes), same batch, same step, CUDA-event timed."""
    _, _, _, xshape, tshape, _, _ = WORKLOADS[workload]
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        model, LpLoss, kind = reference_model(workload, dev)
        step = make_step(model, LpLoss(size_average=False), B, tshape, ar_steps=AR_STEPS.get(workload, 0))
        torch.manual_seed(1)
        x = torch.randn(B, *xshape, device=dev)
        y = torch.randn(B, *tshape, device=dev)
        for _ in range(warmup):
            step(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(x, y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "batch": B, "kind": kind, "steps": steps,
                "path": "torch-CUDA eager (cuFFT, cuBLAS, cuDNN/ATen), allow_tf32=False, same B200"}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        torch.cuda.empty_cache()


# ----------------------------------------------------------------------------------------------------------------------
def spectral_levels(lib, nprof, hbm_gbs):
    """Per U-level roofline of the fused spectral convolution (north_star: 'achieved fraction of HBM and tensor-core roofline
    reported per U-level'): for every distinct call shape and direction, the summed CUDA-event time of the kernels that call
    launched against its ALGORITHMIC bytes and contraction flops.  The tensor-core peak is taken as half the
    measured dense bf16 rate (tf32 operands), MEASURED_PEAKS.json.  None if the library cannot report it."""
    try:
        n = lib.uno_profile_report_levels(None, 0)
        buf = C.create_string_buffer(n + 16)
        lib.uno_profile_report_levels(buf, n + 16)
        rep = json.loads(buf.value.decode())
        tf32_peak = 0.5 * measured_peaks().get("bf16_tflops", 1665.0)
        out = []
        for label, v in rep.items():
            if v["ms"] <= 0 or v["calls"] <= 0:
                continue
            sec = v["ms"] * 1e-3
            gbs = v["bytes"] / sec / 1e9
            tfl = v["flops"] / sec / 1e12
            out.append({"level": label, "calls_per_step": v["calls"] / nprof, "launches_per_call": v["launches"] / v["calls"],
                        "ms_per_call": v["ms"] / v["calls"], "algorithmic_bytes_per_call": v["bytes"] / v["calls"],
                        "GBps": gbs, "hbm_frac": gbs / hbm_gbs, "contraction_TFLOPs": tfl, "tensor_frac": tfl / tf32_peak})
        return out or None
    except Exception as e:   # the headline line must survive a reporting problem
        return {"error": repr(e)}


class Ctx:
    def __init__(self, rank, world, local, dev, lib):
        self.rank, self.world, self.local, self.dev, self.lib = rank, world, local, dev, lib

    def sync_all(self):
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(self, fn, n):
        """n calls of fn between barrier + synchronize on both sides, CUDA events, max over ranks (ms total)."""
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        self.sync_all()
        return ms


def run_workload(ctx, workload, B, steps, warmup, profile=True, breakdown_top=None, graph_mode="auto"):
    """Measure one workload on this job's GPUs: device-resident value, end-to-end from pinned host buffers, per-kernel roofline."""
    from uno_b200.losses import LpLoss
    from uno_b200.parallel import GradReducer

    lib, dev, world, rank = ctx.lib, ctx.dev, ctx.world, ctx.rank
    _, _, _, xshape, tshape, _, _ = WORKLOADS[workload]
    model = build_model(workload, device=dev)
    ar = AR_STEPS.get(workload, 0)
    torch.manual_seed(1 + rank)
    x_host = torch.randn(B, *xshape).pin_memory()
    y_host = torch.randn(B, *tshape).pin_memory()
    x = x_host.to(dev)
    y = y_host.to(dev)

    # Execution mode.  "graph": the whole step (zero_grad, forward / rollout, loss, backward) captured once in a CUDA graph and
    # replayed (uno_b200/graphed.py); the gradient all-reduce follows the replay.  "eager": autograd drives the C ABI call by
    # call and the bucketed all-reduce overlaps backward.  Replay wins wherever host work per step rivals the kernel time --
    # since round 2 that is everywhere: the eager Darcy step issues 1820 launches and measured 18.8 ms on 8 GPUs when the
    # replayed step takes 17.5 ms on one, i.e. the host was the limit; the exposed all-reduce after the replay (65.6 MB, ~0.3 ms
    # over NVLink) costs less.  `--graph off` keeps the eager path.
    step, reducer, rollout, graph_error = None, None, "eager", None
    want_graph = graph_mode in ("on", "auto")
    if want_graph:
        ok = 1
        try:
            from uno_b200.graphed import GraphedStep

            reducer = GradReducer(model, overlap=False)
            step = GraphedStep(model, LpLoss(size_average=False), x, y, ar_steps=ar, reducer=reducer)
            rollout = "graph"
        except Exception as exc:
            ok, graph_error = 0, repr(exc)
        if world > 1:   # every rank runs the same mode
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if not ok:
            step, reducer, rollout = None, None, "eager"
            model.zero_grad(set_to_none=True)
            torch.cuda.empty_cache()
    if step is None:
        reducer = GradReducer(model) if world > 1 else None
        step = make_step(model, LpLoss(size_average=False), B, tshape, reducer, ar_steps=ar)

    for _ in range(warmup):
        step(x, y)
    # --- device-resident timing
    sampler = ClockSampler(ctx.local)
    sampler.start()
    l0 = lib.uno_launch_count()
    total_ms = ctx.timed(lambda: step(x, y), steps)
    launches = lib.uno_launch_count() - l0
    if rollout == "graph":      # replays do not pass through the library's host code: count one eager step of the same work
        l0 = lib.uno_launch_count()
        step.eager_step(x, y)
        torch.cuda.synchronize()
        launches = (lib.uno_launch_count() - l0) * steps
    clocks = sampler.finish()
    ms_per_step = total_ms / steps
    value = B * world / (ms_per_step * 1e-3)

    # --- end to end from pinned host memory
    def e2e_step():
        if rollout == "graph":      # pinned host buffers -> the graph's static inputs -> replay -> loss read back
            return float(step.run_from_host(x_host, y_host).item())
        xd = x_host.to(dev, non_blocking=True)
        yd = y_host.to(dev, non_blocking=True)
        return float(step(xd, yd).item())

    e2e_step()
    e2e_ms = ctx.timed(e2e_step, steps) / steps
    e2e = {"value": B * world / (e2e_ms * 1e-3), "unit": "samples/s",
           "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 4), "d2h_bytes_per_step": 4,
           "ms_per_step": e2e_ms, "copies": "input and target copied from pinned host memory on the compute stream before the step"
                                           + (" (into the captured graph's static inputs)" if rollout == "graph" else "")}

    # Same steps, same bytes, same loss read-back every step, but the inputs are double buffered the way any data loader feeds a
    # training loop: while step i computes, the host-to-device copies of step i+1 run on a copy stream into the other staging
    # pair.  Every step's copy is still issued and completed inside the timed region (the first one is fully exposed, none is
    # issued for a step that does not run).  Reported as `e2e` when it reproduced the loss and was faster; the serial figure
    # (copies on the compute stream ahead of each step) stays next to it.
    try:
        copy_stream = torch.cuda.Stream(device=dev)
        stage = [(torch.empty_like(x), torch.empty_like(y)) for _ in range(2)]
        ready, consumed = [None, None], [None, None]
        state = {"i": 0, "n": 0}

        def prefetch(slot):
            with torch.cuda.stream(copy_stream):
                if consumed[slot] is not None:
                    copy_stream.wait_event(consumed[slot])      # the step that last read this pair has finished with it
                stage[slot][0].copy_(x_host, non_blocking=True)
                stage[slot][1].copy_(y_host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                ready[slot] = ev

        def e2e_step_pipelined():
            i, n = state["i"], state["n"]
            slot = i & 1
            cur = torch.cuda.current_stream(dev)
            if i == 0:
                prefetch(0)
            cur.wait_event(ready[slot])
            if i + 1 < n:
                prefetch(slot ^ 1)
            xd, yd = stage[slot]
            loss = step(xd, yd)
            ev = torch.cuda.Event()
            ev.record(cur)
            consumed[slot] = ev
            state["i"] = i + 1
            return float(loss.item())

        def run_pipelined(n):
            state["i"], state["n"] = 0, n
            for k in range(2):
                ready[k] = consumed[k] = None
            return ctx.timed(e2e_step_pipelined, n)

        ok, why = 1, ""
        try:
            ref_loss = e2e_step()
            state["i"], state["n"] = 0, 2
            got = [e2e_step_pipelined(), e2e_step_pipelined()]
            if any(abs(g_ - ref_loss) > 1e-4 * max(1.0, abs(ref_loss)) for g_ in got):
                ok, why = 0, f"pipelined-input step changed the loss: {got} vs {ref_loss}"
        except Exception as exc:
            ok, why = 0, repr(exc)
        if world > 1:   # every rank takes the same branch (the timed loop below contains collectives)
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if not ok:
            raise RuntimeError(why or "another rank could not run the pipelined-input step")
        pl_ms = run_pipelined(steps) / steps
        if pl_ms < e2e_ms:
            e2e = {"value": B * world / (pl_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
                   "d2h_bytes_per_step": 4, "ms_per_step": pl_ms,
                   "copies": "double-buffered inputs: step i+1's input and target are copied from pinned host memory on a copy stream "
                             "while step i computes; the loss is read back every step",
                   "serial_copy_value": e2e["value"], "serial_copy_ms_per_step": e2e_ms}
        else:   # no gain on this box: the serial-copy figure stays the headline
            e2e["pipelined_copy_value"] = B * world / (pl_ms * 1e-3)
            e2e["pipelined_copy_ms_per_step"] = pl_ms
    except Exception as exc:   # the serial-copy measurement above stands
        e2e["pipeline_error"] = repr(exc)

    # --- per-kernel roofline, CUDA events around every launch of OUR kernels (separate steps so the
    #     event records do not perturb `value`; a captured rollout is re-run eagerly for it: events cannot be captured)
    roofline, breakdown, levels = None, None, None
    if profile:
        hbm, how = peaks()
        prof_step = step.eager_step if hasattr(step, "eager_step") else step
        ctx.sync_all()
        # one kernel at a time while events bracket every launch: the block's two branches otherwise run side by side
        # (switch `overlap`) and each other's time would leak into the per-kernel figures
        from uno_b200 import config as _cfg

        was_overlap = _cfg.get("overlap")
        _cfg.set("overlap", 0)
        prof_step(x, y)      # untimed: first use of the single-stream configuration (per-stream scratch is allocated on first use)
        ctx.sync_all()
        if rank == 0:
            lib.uno_profile_enable(1)
        nprof = 2
        for _ in range(nprof):
            prof_step(x, y)
        ctx.sync_all()
        _cfg.set("overlap", was_overlap)
        if rank == 0:
            n = lib.uno_profile_report(None, 0)
            buf = C.create_string_buffer(n + 16)
            lib.uno_profile_report(buf, n + 16)
            levels = spectral_levels(lib, nprof, hbm)
            lib.uno_profile_enable(0)
            prof = json.loads(buf.value.decode())
            tot = sum(v["ms"] for v in prof.values()) or 1.0
            breakdown = {k: {"launches_per_step": v["launches"] // nprof, "ms_per_step": v["ms"] / nprof,
                             "share_of_uno_kernels": v["ms"] / tot,
                             "GBps": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None,
                             "hbm_frac": v["bytes"] / (v["ms"] * 1e-3) / 1e9 / hbm if v["ms"] > 0 else None,
                             "TFLOPs": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else None}
                         for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
            top, tv = max(prof.items(), key=lambda kv: kv[1]["ms"])
            achieved = tv["bytes"] / (tv["ms"] * 1e-3) / 1e9
            roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                        "traffic": measured_traffic(top, workload), "peak_source": how, "avg_launch_ms": tv["ms"] / tv["launches"],
                        "algorithmic_bytes_per_launch": tv["bytes"] / tv["launches"],
                        "uno_kernel_ms_per_step": tot / nprof}
            if breakdown_top:
                breakdown = dict(list(breakdown.items())[:breakdown_top])
        elif rank != 0:
            pass
    del model, step, reducer, x, y
    torch.cuda.empty_cache()
    return {"workload": WORKLOAD_DESC[workload], "per_gpu_batch": B, "global_batch": B * world, "value": value, "unit": "samples/s",
            "ms_per_step": ms_per_step, "steps": steps, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "execution": rollout, "graph_error": graph_error, "roofline": roofline, "kernel_breakdown": breakdown, "spectral_levels": levels}


def spectral_sweep(iters=3):
    """BASELINE configs[4]: SpectralConv2d_Uno(C,C,S,S,m,m), 512 MiB inputs, fwd and bwd separately,
    against 8(d)'s algorithmic bytes and against the stock torch-CUDA layer (cuFFT rfft2 / complex einsum / irfft2)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sweep_spectral as sw

    hbm, _ = peaks()
    timer = sw.Timer(flush=False)
    rows = []
    for S in (64, 128, 256, 512):
        for m in (12, 20, 32):
            for Cc in (32, 64, 128):
                r = sw.measure(timer, (1 << 27) // (Cc * S * S), Cc, Cc, S, S, S, S, m, m, iters, hbm)
                rows.append({k: r[k] for k in ("B", "Ci", "in", "modes", "fwd_ms", "bwd_ms", "fwd_gbs", "bwd_gbs", "fwd_frac", "bwd_frac",
                                               "speedup_fwd", "speedup_bwd")})
    return {"what": "SpectralConv2d_Uno(C,C,S,S,m,m), 512 MiB input (> L2), median of %d, CUDA events; *_frac = SURVEY 8(d) algorithmic "
                    "bytes / time / measured HBM peak; speedup_* = stock torch-CUDA layer (cuFFT + complex einsum) time / ours" % iters,
            "rows": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="darcy", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the BASELINE config's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="headline split at N > 1: per-GPU batch fixed (weak) or global batch fixed (strong); the other one is reported beside it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step from a CUDA graph (auto / on: every workload at every N, eager fallback if capture fails; off: eager)")
    ap.add_argument("--lean", action="store_true", help="headline workload only (no secondary / reference / sweep legs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.lean:
        args.no_secondary = args.no_reference_gpu = args.no_sweep = args.no_cpu_baseline = True

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cls, cargs, ckw, xshape, tshape, def_b, _ = WORKLOADS[args.workload]
    B = args.batch or def_b

    # ------------------------------------------------------------------ reference arm (the reference on the host cores)
    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = min(args.steps, 3), min(args.warmup, 1)
        r = cpu_reference_run(args.workload, steps, warm)
        line = {
            "impl": "reference", "metric": "UNO samples/sec (fwd+bwd)", "value": r["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "per_gpu_batch": B, "cpu_sample_batch": r["batch"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: while the communicator is created, file descriptor 1 points at stderr so that
        # NCCL's own banner ("NCCL version ...", written straight to stdout under NCCL_DEBUG=VERSION/WARN) lands there
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    from uno_b200 import _lib, build as _build

    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()
    ctx = Ctx(rank, world, local, dev, _lib.get())

    def split(full, mode):
        """per-GPU batch of a workload whose BASELINE batch is `full`"""
        if mode == "weak" or world == 1:
            return full
        return max(full // world, 1)

    main_mode = args.scaling
    head = run_workload(ctx, args.workload, args.batch or split(def_b, main_mode), args.steps, args.warmup, profile=not args.no_profile, graph_mode=args.graph)
    other = None
    if world > 1 and args.batch is None:
        om = "strong" if main_mode == "weak" else "weak"
        o = run_workload(ctx, args.workload, split(def_b, om), args.steps, args.warmup, profile=False, graph_mode=args.graph)
        other = {k: o[k] for k in ("per_gpu_batch", "global_batch", "value", "unit", "ms_per_step", "e2e", "gpu_launches")}
        other["scaling"] = om

    secondary = None
    if not args.no_secondary and args.workload == "darcy":
        secondary = {}
        for wl in SECONDARY:
            full = WORKLOADS[wl][5]
            try:
                modes = ["weak"] if world == 1 else ["weak", "strong"]
                ent = {}
                for mode in modes:
                    r = run_workload(ctx, wl, split(full, mode), min(args.steps, 10), args.warmup, profile=(mode == "weak"), breakdown_top=10, graph_mode=args.graph)
                    r["scaling"] = mode
                    ent[mode] = r
                secondary[wl] = ent["weak"] if world == 1 else ent
            except Exception as exc:   # the headline must survive a problem in a secondary workload
                if world > 1:
                    raise
                secondary[wl] = {"error": repr(exc)}

    cpu_baseline = reference_gpu = sweep = None
    if rank == 0 and world == 1:
        if not args.no_reference_gpu:
            reference_gpu = {}
            mine = {args.workload: head}
            mine.update({k: v for k, v in (secondary or {}).items() if "value" in v})
            for wl, r in mine.items():
                try:
                    g = reference_gpu_run(wl, r["per_gpu_batch"], 5, 2, dev)
                    g["speedup"] = r["value"] / g["value"]
                    reference_gpu[wl] = g
                except Exception as exc:
                    reference_gpu[wl] = {"error": repr(exc)}
        if not args.no_sweep and args.workload == "darcy":
            try:
                sweep = spectral_sweep()
            except Exception as exc:
                sweep = {"error": repr(exc)}
        if not args.no_cpu_baseline:
            r = cpu_reference_run(args.workload, 3, 1, batch=WORKLOADS[args.workload][6])
            cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            for wl, ent in (secondary or {}).items():
                if "value" in ent:
                    try:
                        rr = cpu_reference_run(wl, 2, 1, batch=WORKLOADS[wl][6])
                        ent["cpu_baseline"] = {k: rr[k] for k in ("value", "unit", "cores", "kind", "sample")}
                    except Exception as exc:
                        ent["cpu_baseline"] = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": "UNO samples/sec (fwd+bwd)", "value": head["value"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": main_mode, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "per_gpu_batch": head["per_gpu_batch"], "global_batch": head["global_batch"],
                       "parallelism": f"dp{world} (batch shard, flat-buffer gradient all-reduce)" if world > 1 else "single GPU",
                       "l2": "per-step working set (several GB of activations) exceeds the 126 MB L2; no explicit flush"},
            "clocks": head["clocks"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
            "execution": head["execution"], "graph_error": head["graph_error"],
            "roofline": head["roofline"], "cpu_baseline": cpu_baseline, "reference_gpu": reference_gpu,
            ("strong" if main_mode == "weak" else "weak"): other, "secondary": secondary,
            "kernel_breakdown": head["kernel_breakdown"], "spectral_levels": head["spectral_levels"], "sweep": sweep,
            "src_sha16": source_sha16(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
