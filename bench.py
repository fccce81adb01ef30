#!/usr/bin/env python
"""Benchmark of the U-NO hot path on B200 (contract: see the task prompt / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload darcy|ns2d|ns3d]

Default workload = BASELINE.json configs[1]: UNO_9(3, 32, pad=12) (darcy_flow_main.py:95) on synthetic
421x421 Darcy inputs, batch 32 per GPU, forward + rel-L2 loss + backward (train_darcy.py:50-54), no
optimizer step.  One JSON line is printed by rank 0.

  value      samples/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same step driven from pinned HOST buffers (H2D of x,y and D2H of the loss inside the timed region)
  roofline   dominant kernel family of the step: algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline / --impl reference: the oracle torch port (same MKL/ATen calls as the reference's CPU path)
             timed on the host cores on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
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


def measured_traffic(role, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `role`, from the committed ncu --set full capture of
    one step of this workload (profiles/r01_traffic.json, made by tools/ncu_traffic.py); None when there is none."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if workload != "darcy" or not os.path.exists(path):
        return None
    try:
        return float(json.load(open(path))["roles"][role]["dram_bytes_per_launch"])
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


def make_step(model, loss_fn, B, tshape, reducer=None, ar_steps=0):
    def step(x, y, before_loss=None):
        """`before_loss`: called once after the first forward, before the target is first read (the end-to-end arm waits there
        for the target's host-to-device copy, which it lets overlap the forward)."""
        if reducer is not None:
            reducer.zero_grad()
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


def cpu_reference_run(workload, steps, warmup, batch=None):
    """The oracle torch port (kind "port") on the host cores: fwd + loss + bwd, bounded batch."""
    from oracle import uno_torch_port as port

    LpLoss = port.LpLoss          # the CPU arm runs no code of the product
    _, _, _, xshape, tshape, _, cpu_b = WORKLOADS[workload]
    B = batch or cpu_b
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = build_model(workload, ops=port, device="cpu")
    torch.manual_seed(1)
    x = torch.randn(B, *xshape)
    y = torch.randn(B, *tshape)
    step = make_step(model, LpLoss(size_average=False), B, tshape, ar_steps=AR_STEPS.get(workload, 0))
    for _ in range(warmup):
        step(x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(x, y)
    dt = (time.perf_counter() - t0) / steps
    return {"value": B / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"batch {B} of the same workload, {steps} steps after {warmup} warm-up, fp32 torch CPU (oracle/uno_torch_port.py)",
            "ms_per_step": dt * 1e3, "batch": B}


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="darcy", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the BASELINE config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cls, cargs, ckw, xshape, tshape, def_b, _ = WORKLOADS[args.workload]
    B = args.batch or def_b

    # ------------------------------------------------------------------ reference arm (CPU port)
    if args.impl == "reference":
        if rank != 0:
            return
        steps = min(args.steps, 5)
        r = cpu_reference_run(args.workload, steps, min(args.warmup, 1))
        line = {
            "impl": "reference", "metric": "UNO samples/sec (fwd+bwd)", "value": r["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
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
    from uno_b200.losses import LpLoss
    from uno_b200.parallel import GradReducer

    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()
    lib = _lib.get()

    model = build_model(args.workload, device=dev)
    reducer = GradReducer(model) if world > 1 else None
    loss_fn = LpLoss(size_average=False)
    step = make_step(model, loss_fn, B, tshape, reducer, ar_steps=AR_STEPS.get(args.workload, 0))

    torch.manual_seed(1 + rank)
    x_host = torch.randn(B, *xshape).pin_memory()
    y_host = torch.randn(B, *tshape).pin_memory()
    x = x_host.to(dev)
    y = y_host.to(dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, n):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        sync_all()
        return ms

    for _ in range(args.warmup):
        step(x, y)
    # --- device-resident timing
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.uno_launch_count()
    total_ms = timed(lambda: step(x, y), args.steps)
    launches = lib.uno_launch_count() - l0
    clocks = sampler.finish()
    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # --- end to end from pinned host memory
    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        yd = y_host.to(dev, non_blocking=True)
        return float(step(xd, yd).item())

    e2e_step()
    e2e_ms = timed(e2e_step, args.steps) / args.steps
    e2e = {"value": B * world / (e2e_ms * 1e-3), "unit": "samples/s",
           "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 4), "d2h_bytes_per_step": 4,
           "ms_per_step": e2e_ms, "copies": "input and target copied on the compute stream before the step"}

    # Same step, same bytes, same loss read-back, but the target's copy is issued on a copy stream right behind the input's and
    # the compute stream only waits for it where the loss first reads it: half of the host-to-device time hides behind the
    # forward.  Reported as `e2e` when it ran; the serial-copy figure stays next to it.
    try:
        copy_stream = None

        def e2e_step_overlap():
            cur = torch.cuda.current_stream(dev)
            xd = x_host.to(dev, non_blocking=True)
            x_done = torch.cuda.Event()
            x_done.record(cur)
            copy_stream.wait_event(x_done)
            with torch.cuda.stream(copy_stream):
                yd = y_host.to(dev, non_blocking=True)
                y_done = torch.cuda.Event()
                y_done.record(copy_stream)

            def before_loss():
                cur.wait_event(y_done)
                yd.record_stream(cur)

            return float(step(xd, yd, before_loss).item())

        ok, why = 1, ""
        try:
            copy_stream = torch.cuda.Stream(device=dev)
            ref_loss = e2e_step()
            got_loss = e2e_step_overlap()
            if abs(got_loss - ref_loss) > 1e-4 * max(1.0, abs(ref_loss)):
                ok, why = 0, f"overlapped-copy step changed the loss: {got_loss} vs {ref_loss}"
        except Exception as exc:
            ok, why = 0, repr(exc)
        if world > 1:   # every rank takes the same branch (the timed loop below contains collectives)
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if not ok:
            raise RuntimeError(why or "another rank could not run the overlapped-copy step")
        ov_ms = timed(e2e_step_overlap, args.steps) / args.steps
        if ov_ms < e2e_ms:
            e2e = {"value": B * world / (ov_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
                   "d2h_bytes_per_step": 4, "ms_per_step": ov_ms,
                   "copies": "input copied on the compute stream, target on a copy stream behind it (waited for at the loss)",
                   "serial_copy_value": e2e["value"], "serial_copy_ms_per_step": e2e_ms}
        else:   # no gain on this box: the serial-copy figure stays the headline
            e2e["overlapped_copy_value"] = B * world / (ov_ms * 1e-3)
            e2e["overlapped_copy_ms_per_step"] = ov_ms
    except Exception as exc:   # the serial-copy measurement above stands
        e2e["overlap_error"] = repr(exc)

    # --- per-kernel roofline, CUDA events around every launch of OUR kernels (separate steps so the
    #     event records do not perturb `value`)
    roofline, breakdown, levels = None, None, None
    if not args.no_profile:
        # every rank runs these steps (they contain the gradient all-reduce); only rank 0 records events
        hbm, how = peaks()
        sync_all()
        if rank == 0:
            lib.uno_profile_enable(1)
        nprof = 2
        for _ in range(nprof):
            step(x, y)
        sync_all()
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
                             "TFLOPs": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else None}
                         for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
            top, tv = max(prof.items(), key=lambda kv: kv[1]["ms"])
            achieved = tv["bytes"] / (tv["ms"] * 1e-3) / 1e9
            roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                        "traffic": measured_traffic(top, args.workload), "peak_source": how, "avg_launch_ms": tv["ms"] / tv["launches"],
                        "algorithmic_bytes_per_launch": tv["bytes"] / tv["launches"],
                        "uno_kernel_ms_per_step": tot / nprof}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.workload, 3, 1)
        cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": "UNO samples/sec (fwd+bwd)", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "per_gpu_batch": B, "global_batch": B * world,
                       "parallelism": f"dp{world} (batch shard, flat-buffer gradient all-reduce)" if world > 1 else "single GPU",
                       "l2": "per-step working set (several GB of activations) exceeds the 126 MB L2; no explicit flush"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "kernel_breakdown": breakdown, "spectral_levels": levels,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
