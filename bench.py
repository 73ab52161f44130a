#!/usr/bin/env python
"""bench.py -- PDE-loss fwd+bwd Gpixels/s on B200 (BASELINE.json metric), with roofline, end-to-end
and CPU-baseline numbers on the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg5|cfg1] [--dtype f32|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU PyTorch loss (torch port) on host cores

A "step" = one forward + one backward of the Stage II loss (Dice + BCE + reaction-diffusion +
phase-field; D=5, a=0.5, eps=0.05, weights 0.5/0.5/1e-4/1e-4) over one batch of synthetic maps, logits
entry (sigmoid fused).  Default workload: 64 x 1 x 1024 x 1024 fp32 per GPU -- the shape the north-star
target is quoted on; weak scaling (every rank holds that batch; the only exchange is the 64-byte
all-reduce of the partial sums between the two kernels).

value  : K steps on tensors already resident in HBM, CUDA events, max over ranks.
e2e    : the same through the host-buffer C-ABI call (pil_session_run): pinned host maps -> H2D ->
         kernels -> D2H of the loss report and the gradient, all inside the timed region.
roofline: per-kernel CUDA-event durations inside the timed region; algorithmic bytes = 8 B/px forward
         (read x, t) and 12 B/px backward (read x, t, write grad) for fp32.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B, H, W, description)
    "cfg1": (8, 256, 256, "stage2_loss_fwd_bwd_8x1x256x256"),
    "cfg2": (32, 512, 512, "stage2_loss_fwd_bwd_32x1x512x512"),
    "cfg3": (64, 1024, 1024, "stage2_loss_fwd_bwd_64x1x1024x1024"),
    "cfg5": (128, 2048, 2048, "stage2_loss_fwd_bwd_128x1x2048x2048"),
}
STAGE2 = dict(dice_weight=0.5, bce_weight=0.5, pde_weight=1e-4, phase_field_weight=1e-4,
              diffusion_coeff=5.0, reaction_threshold=0.5, epsilon=0.05, smooth=1e-6)
METRIC = "pde_loss_fwd_bwd_gpixels_per_s"
UNIT = "Gpixel/s"


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth(B, H, W, seed, device, dtype):
    """iid synthetic maps of SURVEY.md 8d: z = 2*randn logits, t = Bernoulli(0.5) masks."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    z = (2.0 * torch.randn(B, 1, H, W, device=device, generator=g)).to(dtype)
    t = (torch.rand(B, 1, H, W, device=device, generator=g) > 0.5).to(dtype)
    return z.contiguous(), t.contiguous()


# --------------------------------------------------------------------------------------------------
# CPU arms: the reference's CPU PyTorch loss, restated op for op in oracle/torch_port.py
# --------------------------------------------------------------------------------------------------
def cpu_port_time(sample_B, H, W, iters, warm):
    import torch

    from oracle import pil_oracle as po
    from oracle import torch_port

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    z = 2.0 * torch.randn(sample_B, 1, H, W, generator=g)
    t = (torch.rand(sample_B, 1, H, W, generator=g) > 0.5).float()
    for _ in range(warm):
        torch_port.fwd_bwd(z, t, po.STAGE2, 1)
    times = []
    for _ in range(iters):
        t0 = time.perf_counter()
        torch_port.fwd_bwd(z, t, po.STAGE2, 1)
        times.append(time.perf_counter() - t0)
    px = sample_B * H * W
    return px, times, torch.get_num_threads()


def run_reference(args):
    """--impl reference: rank 0 only.  Each step = a bounded sample of the workload -- up to 4 whole images,
    fewer (down to a band of rows of one image) when --steps is large -- sized from one probe step so that
    the whole run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, H, W, name = WORKLOADS[args.workload]
    _, probe, _ = cpu_port_time(1, H, W, 1, 1)            # one image, after one warm-up
    budget_s = 150.0 / max(args.steps + args.warmup, 1)    # per step
    sample_H = H
    if probe[0] <= budget_s:
        sample_B = max(1, min(B, 4, int(budget_s / probe[0])))
    else:
        sample_B = 1
        sample_H = max(16, min(H, int(H * budget_s / probe[0]) // 16 * 16))
    px, times, cores = cpu_port_time(sample_B, sample_H, W, args.steps, args.warmup)
    total = sum(times)
    value = px * len(times) / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}_fp32", "stage2_params": STAGE2, "entry": "sigmoid -> loss -> backward on CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample_B}x1x{sample_H}x{W} per step (of {B}x1x{H}x{W}), torch CPU ops, op-for-op port of the reference"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200 import functional as Fn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this package has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)

    B, H, W, name = WORKLOADS[args.workload]
    dtype = torch.float32 if args.dtype == "f32" else torch.bfloat16
    esz = 4 if args.dtype == "f32" else 2
    z, t = synth(B, H, W, 1234 + rank, dev, dtype)
    grad = torch.empty_like(z)
    p = P.LossParams(**STAGE2)
    kind = Fn.X_LOGITS_SIGMOID
    sums = torch.empty(8, dtype=torch.float64, device=dev)
    report = torch.empty(8, dtype=torch.float32, device=dev)
    n_local = B * H * W
    n_global = n_local * world

    sb = torch.empty(8, dtype=torch.float64, device=dev)
    tot = torch.empty(8, dtype=torch.float64, device=dev)

    # Data-parallel exchange of the 8-double sums vectors: "peer" = the kernels swap them themselves through
    # peer-mapped mailboxes over NVLink (include/pil.h PilExchange; no collective call, 2 launches per
    # step); "nccl" = all-reduce between the kernels (the baseline this replaces).
    exchange = args.exchange if distributed else "none"
    px = None
    if exchange == "peer":
        from physics_informed_image_segmentation_b200.sharding import PeerExchange

        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            px = PeerExchange(dev)
        except Exception as exc:  # no peer access between these GPUs: every rank falls back together
            print(f"[rank {rank}] peer-memory exchange unavailable ({exc}); using the NCCL all-reduce path", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if px is not None:
                px.close()
            px, exchange = None, "nccl"

    def fwd_part(ex=None):
        """K1L: pointwise sums (I, P, T, BCE, double well), one flat pass over x and t."""
        if ex is not None:
            Fn.forward_pointwise_xchg(z, t, p, kind, ex, sums=sums)
        else:
            Fn.forward_pointwise(z, t, p, kind, sums=sums)

    def bwd_part(ex=None):
        """[exchange of the 8 doubles] -> K2: gradient + the two stencil sums -> loss report."""
        if ex is not None:
            Fn.backward_accumulate_xchg(z, t, p, kind, ex, n_global, grad_scale=float(world), out=grad, stencil_sums=sb,
                                        report=report, total_sums=tot)
        elif distributed:
            dist.all_reduce(sums)                                   # the gradient needs the global I, P, T
            Fn.backward_accumulate(z, t, p, kind, sums, n_global, grad_scale=float(world), out=grad, stencil_sums=sb)
            dist.all_reduce(sb)                                     # only the loss value needs these
            Fn.finalize_report(sums + sb, n_global, p, report=report)
        else:
            Fn.backward_accumulate(z, t, p, kind, sums, n_global, out=grad, stencil_sums=sb, report=report)

    def step():
        ex = px.next_step() if px is not None else None
        fwd_part(ex)
        bwd_part(ex)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # runs through warm-up and both timed regions
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- short instrumented pass (roofline of the individual kernels, same burst regime the measured copy peak
    # was taken in): an event between the two kernels of every step.
    K = args.steps
    KA = max(3, min(K, 30))
    eva = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(KA)]
    for k in range(KA):
        ex = px.next_step() if px is not None else None
        eva[k][0].record()
        fwd_part(ex)
        eva[k][1].record()
        bwd_part(ex)
        eva[k][2].record()
    torch.cuda.synchronize()
    fwd_burst = [eva[k][0].elapsed_time(eva[k][1]) for k in range(KA)]
    bwd_burst = [eva[k][1].elapsed_time(eva[k][2]) for k in range(KA)]

    # ---- timed region 1 (the headline `value`): exactly K steps back to back, one event pair around them.
    # The two kernels of a step are launched with programmatic dependent launch, so the next kernel's
    # blocks fill the SMs while the previous one drains; nothing is recorded between them.
    k0 = Fn.launch_info().kernels_launched
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    e_beg.record()
    for k in range(K):
        step()
    e_end.record()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    wall = time.perf_counter() - t_wall0
    info = Fn.launch_info()
    launches = info.kernels_launched - k0
    total_ms = e_beg.elapsed_time(e_end)

    # ---- sustained instrumented pass: the same steps with an event between the two kernels, repeated for at
    # least ~0.3 s so that the clock sampler (50 ms period) sees the load.
    K2 = max(K, min(20000, int(0.3 / max(total_ms * 1e-3 / K, 1e-6)) + 1))
    tk = torch.tensor([K2], dtype=torch.int64, device=dev)
    if distributed:
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)  # every rank must run the same number of exchange steps
    K2 = int(tk.item())
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K2)]
    torch.cuda.synchronize()
    for k in range(K2):
        ex = px.next_step() if px is not None else None
        ev[k][0].record()
        fwd_part(ex)
        ev[k][1].record()
        bwd_part(ex)
        ev[k][2].record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    fwd_ms = [ev[k][0].elapsed_time(ev[k][1]) for k in range(K2)]
    bwd_ms = [ev[k][1].elapsed_time(ev[k][2]) for k in range(K2)]  # multi-GPU: includes the exchange wait
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms_max = float(tt.item())
    value = n_global * K / (total_ms_max * 1e-3) / 1e9
    loss_val = float(report[0].item())
    xchg_timeout = bool(px.timed_out()) if px is not None else False

    # ---- end to end through the host-buffer C ABI --------------------------------------------------
    e2e = None
    if not args.no_e2e:
        zh = z.cpu().pin_memory()
        th = t.cpu().pin_memory()
        gh = torch.empty_like(zh).pin_memory()
        x_code = 0 if args.dtype == "f32" else 1
        with P.HostSession(B, H, W, device=local_rank, x_dtype=x_code, t_dtype=x_code) as sess:
            Ke = max(3, min(K, 10))

            def timed(**kw):
                for _ in range(2):
                    r = sess.run(zh, th, p, **kw)
                if distributed:
                    dist.barrier()
                t0 = time.perf_counter()
                for _ in range(Ke):
                    r = sess.run(zh, th, p, **kw)
                dt_ = time.perf_counter() - t0
                te_ = torch.tensor([dt_], dtype=torch.float64, device=dev)
                if distributed:
                    dist.all_reduce(te_, op=dist.ReduceOp.MAX)
                return float(te_.item()), r

            # headline: what a training step moves -- inputs in, loss report out, gradient stays on the device
            dt_dev, rep = timed(grad_on_device=True)
            # and the same with the gradient map copied back to the host as well
            dt_host, rep_h = timed(grad_host=gh)
        e2e = {"value": n_local * world * Ke / dt_dev / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": 2 * n_local * esz * world, "d2h_bytes_per_step": 32 * world,
               "steps": Ke, "ms_per_step": 1e3 * dt_dev / Ke,
               "note": "pil_session_run_ex(PIL_SESSION_GRAD_ON_DEVICE): pinned host x,t -> H2D (chunked, overlapped with the "
                       "pointwise forward) -> backward over the batch (gradient stays in HBM, as a training step consumes it) "
                       "-> D2H of the loss report; host wall clock, PCIe-bound"
                       + ("; per-rank session (independent shard losses)" if distributed else ""),
               "loss": float(rep[0]),
               "with_gradient_d2h": {"value": n_local * world * Ke / dt_host / 1e9, "ms_per_step": 1e3 * dt_host / Ke,
                                     "d2h_bytes_per_step": (n_local * esz + 32) * world, "loss": float(rep_h[0])}}

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        # per-kernel durations: the short pass before the headline region (burst); the medians over the >= 0.3 s
        # pass after it (clocks settling under the power cap) are reported beside them
        fwd_med, bwd_med = statistics.median(fwd_burst), statistics.median(bwd_burst)
        fwd_sus, bwd_sus = statistics.median(fwd_ms), statistics.median(bwd_ms)
        bpp_f, bpp_b = 2 * esz, 3 * esz
        ach_b = bpp_b * n_local / (bwd_med * 1e-3) / 1e9
        ach_f = bpp_f * n_local / (fwd_med * 1e-3) / 1e9
        ach_step = (bpp_f + bpp_b) * n_local * K / (total_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch of the backward kernel from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            if tj.get("workload") == f"{name}_{'fp32' if args.dtype == 'f32' else 'bf16'}":
                kb = tj["kernels"]["pil_bwd_kernel"]
                traffic, traffic_src = kb["dram_bytes_read"] + kb["dram_bytes_write"], tj.get("source")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{name}_{'fp32' if args.dtype == 'f32' else 'bf16'}", "per_gpu_batch": [B, 1, H, W],
                       "global_batch": B * world, "stage2_params": STAGE2, "entry": "logits (sigmoid fused)",
                       "parallelism": f"dp{world} (batch shards; exchange of 8 doubles between the two kernels: {exchange})",
                       "l2_policy": "inputs+gradient %.0f MB per step >> 126 MB L2, no flush needed" % (3 * n_local * esz / 1e6),
                       "step": "pil_forward_pointwise -> pil_backward_accumulate (stencils evaluated once per step)",
                       "tiling": {"fwd_blocks": info.fwd_blocks, "bwd_blocks": info.bwd_blocks,
                                  "bwd_rows_per_range": info.bwd_rows_per_segment}},
            "roofline": {"bound": "hbm", "kernel": "pil_bwd_kernel (gradient + stencil sums)", "achieved": ach_b, "peak": peak, "unit": "GB/s",
                         "frac": ach_b / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bpp_b * n_local, "kernel_ms": bwd_med, "kernel_launches_timed": KA,
                         "sustained": {"kernel_ms": bwd_sus, "launches": K2, "frac": bpp_b * n_local / (bwd_sus * 1e-3) / 1e9 / peak},
                         "note": "kernel_ms: median CUDA-event interval around the launch in a short pass right after warm-up (an event "
                                 "between the two kernels of every step; burst regime, like the measured copy peak); sustained: the same "
                                 "over a >= 0.3 s pass after the headline region (clocks settle under the power cap); "
                                 + ("multi-GPU: the interval includes the wait for the other ranks' sums" if distributed else "single GPU")},
            "roofline_fwd": {"kernel": "pil_point_kernel (pointwise sums)", "achieved": ach_f, "frac": ach_f / peak, "kernel_ms": fwd_med,
                             "sustained_kernel_ms": fwd_sus,
                             "algorithmic_bytes_per_launch": bpp_f * n_local},
            "roofline_step": {"achieved": ach_step, "frac": ach_step / peak, "bytes_per_pixel": bpp_f + bpp_b},
            "gpu_launches": int(launches), "clocks": clocks, "loss": loss_val, "wall_ms_per_step": 1e3 * wall / K,
            "exchange_timeout": xchg_timeout,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu and args.dtype == "f32":
            # the reference's op sequence (eager PyTorch: reflect pad, 3 one-channel conv2d, BCELoss, ~40 kernels forward,
            # ~100 backward) on THIS GPU -- what the unmodified reference costs once its tensors are on the B200
            try:
                from oracle import pil_oracle as po_
                from oracle import torch_port

                sb_ = max(1, min(B, (16 * 1024 * 1024) // (H * W)))
                ze, te = z[:sb_].float().clone(), t[:sb_].float().clone()
                for _ in range(3):
                    torch_port.fwd_bwd(ze, te, po_.STAGE2, 1)
                torch.cuda.synchronize()
                ee = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                ee[0].record()
                for _ in range(5):
                    torch_port.fwd_bwd(ze, te, po_.STAGE2, 1)
                ee[1].record()
                torch.cuda.synchronize()
                ms_e = ee[0].elapsed_time(ee[1]) / 5
                line["gpu_eager_baseline"] = {"value": sb_ * H * W / (ms_e * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_e,
                                              "sample": f"{sb_}x1x{H}x{W}, mean of 5 after 3 warm-ups",
                                              "kind": f"op-for-op port of the reference loss, eager torch {torch.__version__} on the same GPU"}
                del ze, te
                torch.cuda.empty_cache()
            except Exception as exc:  # informational leg: never fail the bench on it
                line["gpu_eager_baseline"] = {"unavailable": str(exc)[:200]}
        if world == 1 and not args.no_cpu:
            sample_B = max(1, min(B, (8 * 1024 * 1024) // (H * W)))
            npx, times, cores = cpu_port_time(sample_B, H, W, iters=3, warm=1)
            line["cpu_baseline"] = {"value": npx / statistics.median(times) / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample_B}x1x{H}x{W} (of {B}x1x{H}x{W}), median of 3 after 1 warm-up, "
                                              f"torch {torch.__version__} CPU ops, op-for-op port of the reference loss"}
        emit_line(line)
    if px is not None:
        px.close()
    if distributed:
        dist.destroy_process_group()


class _StdoutGuard:
    """Everything libraries write to stdout while the benchmark runs (NCCL prints its version banner there)
    goes to stderr; the ONE JSON line is written to the real stdout at the end."""

    def __enter__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text: str) -> None:
        sys.stdout.flush()
        os.write(self._real, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._real, 1)
        os.close(self._real)


_GUARD = None


def emit_line(line: dict) -> None:
    text = json.dumps(line)
    if _GUARD is not None:
        _GUARD.emit(text)
    else:
        print(text, flush=True)


def main():
    global _GUARD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cfg3")
    ap.add_argument("--dtype", choices=["f32", "bf16"], default="f32")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer",
                    help="multi-GPU: how the 8-double sums vectors cross ranks (peer-memory mailboxes | NCCL all-reduce)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    with _StdoutGuard() as guard:
        _GUARD = guard
        try:
            if args.impl == "reference":
                run_reference(args)
            else:
                run_ours(args)
        finally:
            _GUARD = None


if __name__ == "__main__":
    main()
