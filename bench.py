#!/usr/bin/env python
"""bench.py -- PDE-loss fwd+bwd Gpixels/s on B200 (BASELINE.json metric), with roofline, end-to-end
and CPU-baseline numbers on the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg5|cfg1|cfg4|cfg3_step] [--dtype f32|bf16]
                    [--scaling auto|strong|weak]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU PyTorch loss (torch port) on host cores

A "step" = one forward + one backward of the Stage II loss (Dice + BCE + reaction-diffusion +
phase-field; D=5, a=0.5, eps=0.05, weights 0.5/0.5/1e-4/1e-4) over one batch of synthetic maps, logits
entry (sigmoid fused).  Default workload: a GLOBAL batch of 64 x 1 x 1024 x 1024 fp32 -- the shape the north-star
target is quoted on.  With N > 1 GPUs the default is STRONG scaling: the global batch is sharded by whole images
(sharding.shard_bounds), the kernels exchange their partial sums over peer memory, every rank holds the global loss;
the weak-scaling figure (every rank holds the full batch) and the cfg5 (128 x 2048^2) strong-scaling point are
measured in the same run and reported beside it.  cfg4 = the S2/S3 parameter sweep as one batched loss evaluation;
cfg3_step = U-Net-shaped forward -> fused loss -> backward -> AdamW under DDP (BASELINE config 3).

value  : K steps on tensors already resident in HBM, CUDA events, max over ranks.
e2e    : the same through the host-buffer C-ABI call (pil_session_run): pinned host maps -> H2D ->
         kernels -> D2H of the loss report and the gradient, all inside the timed region.
roofline: per-kernel CUDA-event durations inside the timed region; algorithmic bytes = 8 B/px forward
         (read x, t) and 12 B/px backward (read x, t, write grad) for fp32.
"""
from __future__ import annotations

import argparse
import contextlib
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
    "ref128": (8, 128, 128, "stage2_loss_fwd_bwd_8x1x128x128"),          # the reference's real batch (src/dataset.py:18)
    "cfg4": (256, 512, 512, "s2_s3_sweep_11_settings_256x1x512x512"),      # BASELINE config 4 (run_ablation.py:159-224)
    "cfg3_step": (64, 1024, 1024, "unet_train_step_64x1x1024x1024"),       # BASELINE config 3
}
L2_BYTES = 126 * 1024 * 1024
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
    if args.workload in ("cfg4", "cfg3_step"):
        emit_line({"impl": "reference", "unavailable": f"the reference arm times the loss workloads (cfg1/2/3/5); {args.workload} has no CPU leg"})
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
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak" if args.scaling == "weak" else "strong",
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
class Ctx:
    """Process / device / group of this rank (one process per GPU)."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: this package has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.distributed = self.world > 1
        if self.distributed:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.distributed:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.distributed:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok: bool) -> bool:
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        if self.distributed:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def close(self):
        if self.distributed:
            self.dist.destroy_process_group()


class Exchanges:
    """The two peer-memory exchanges of a run: host-epoch mailboxes for direct calls (per-kernel events, host-buffer
    sessions) and device-epoch mailboxes for captured steps.  None when single-GPU or when --exchange nccl."""

    def __init__(self, C: Ctx, mode: str):
        self.mode = mode if C.distributed else "none"
        self.host = self.device = None
        if self.mode != "peer":
            return
        from physics_informed_image_segmentation_b200.sharding import PeerExchange

        ok = True
        try:
            self.host = PeerExchange(C.dev)
            self.device = PeerExchange(C.dev, device_epoch=True)
        except Exception as exc:  # no peer access between these GPUs: every rank falls back together
            print(f"[rank {C.rank}] peer-memory exchange unavailable ({exc}); using the NCCL all-reduce path", file=sys.stderr)
            ok = False
        if not C.all_ok(ok):
            self.close()
            self.mode = "nccl"

    def timed_out(self) -> bool:
        return any(px is not None and px.timed_out() for px in (self.host, self.device))

    def close(self):
        for px in (self.host, self.device):
            if px is not None:
                px.close()
        self.host = self.device = None


def global_batch(B, H, W, dev, dtype, lo, hi, seed=1234):
    """Images [lo, hi) of the GLOBAL synthetic batch, generated image by image so that every rank (and a single GPU
    evaluating the whole batch) sees bit-identical maps without materialising more than it needs."""
    import torch

    zs, ts = [], []
    for b in range(lo, hi):
        z, t = synth(1, H, W, seed * 1000003 + b, dev, dtype)
        zs.append(z)
        ts.append(t)
    return torch.cat(zs).contiguous(), torch.cat(ts).contiguous()


class LossShard:
    """This rank's shard of a loss workload: rotating buffer sets (L2 hygiene), direct and captured steps."""

    def __init__(self, C: Ctx, X: Exchanges, z, t, n_global: int, launch: str):
        import physics_informed_image_segmentation_b200 as P
        from physics_informed_image_segmentation_b200 import functional as Fn

        torch = C.torch
        self.C, self.X, self.Fn, self.P = C, X, Fn, P
        self.p = P.LossParams(**STAGE2)
        self.kind = Fn.X_LOGITS_SIGMOID
        self.n_local, self.n_global = z.numel(), n_global
        self.esz = z.element_size()
        self.footprint = 3 * self.n_local * self.esz
        # inputs + gradient of one step must not be served from L2 by the step before: either they are larger than
        # twice the L2, or the steps cycle through enough buffer sets that a set comes back only after > 2 x L2 of
        # other traffic
        self.n_sets = 1 if self.footprint >= 2 * L2_BYTES else min(64, -(-2 * L2_BYTES // self.footprint) + 1)
        self.sets = [(z, t, torch.empty_like(z))]
        for _ in range(self.n_sets - 1):
            self.sets.append((z.clone(), t.clone(), torch.empty_like(z)))
        dev = C.dev
        self.sums = torch.empty(8, dtype=torch.float64, device=dev)
        self.report = torch.empty(8, dtype=torch.float32, device=dev)
        self.sb = torch.empty(8, dtype=torch.float64, device=dev)
        self.tot = torch.empty(8, dtype=torch.float64, device=dev)
        self.use_graph = launch == "graph" or (launch == "auto" and self.n_local <= 16 * 1024 * 1024 and X.mode in ("none", "peer"))
        self.graphs = []
        self.k = 0
        if self.use_graph:
            for (zz, tt, gg) in self.sets:
                self.graphs.append(Fn.StepGraph(zz, tt, self.p, self.kind, grad=gg, exchange=X.device, n_global=n_global,
                                                grad_scale=float(C.world)))

    def fwd_part(self, s, ex):
        z, t, _ = self.sets[s]
        if ex is not None:
            self.Fn.forward_pointwise_xchg(z, t, self.p, self.kind, ex, sums=self.sums)
        else:
            self.Fn.forward_pointwise(z, t, self.p, self.kind, sums=self.sums)

    def bwd_part(self, s, ex):
        z, t, g = self.sets[s]
        Fn, C = self.Fn, self.C
        if ex is not None:
            Fn.backward_accumulate_xchg(z, t, self.p, self.kind, ex, self.n_global, grad_scale=float(C.world), out=g,
                                        stencil_sums=self.sb, report=self.report, total_sums=self.tot)
        elif C.distributed:
            C.dist.all_reduce(self.sums)                                 # the gradient needs the global I, P, T
            Fn.backward_accumulate(z, t, self.p, self.kind, self.sums, self.n_global, grad_scale=float(C.world), out=g,
                                   stencil_sums=self.sb)
            C.dist.all_reduce(self.sb)                                   # only the loss value needs these
            Fn.finalize_report(self.sums + self.sb, self.n_global, self.p, report=self.report)
        else:
            Fn.backward_accumulate(z, t, self.p, self.kind, self.sums, self.n_global, out=g, stencil_sums=self.sb, report=self.report)

    def direct_step(self, events=None):
        s = self.k % self.n_sets
        self.k += 1
        ex = self.X.host.next_step() if self.X.host is not None else None
        if events is not None:
            events[0].record()
        self.fwd_part(s, ex)
        if events is not None:
            events[1].record()
        self.bwd_part(s, ex)
        if events is not None:
            events[2].record()

    def step(self):
        if self.use_graph:
            g = self.graphs[self.k % self.n_sets]
            self.k += 1
            self.last_report = g.launch()
        else:
            self.direct_step()
            self.last_report = self.report

    def instrumented(self, n):
        """n direct steps with an event between the two kernels; returns (fwd_ms[], bwd_ms[])."""
        torch = self.C.torch
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
        for k in range(n):
            self.direct_step(ev[k])
        torch.cuda.synchronize()
        return [e[0].elapsed_time(e[1]) for e in ev], [e[1].elapsed_time(e[2]) for e in ev]

    def timed(self, K):
        """Exactly K steps between a barrier + synchronize on both sides; CUDA events; max over ranks.  Returns
        (ms of the slowest rank, host wall seconds, launches of the library's kernels on this rank)."""
        C, torch = self.C, self.C.torch
        k0 = self.Fn.launch_info().kernels_launched
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        C.barrier()
        w0 = time.perf_counter()
        e0.record()
        for _ in range(K):
            self.step()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        C.barrier()
        launches = (2 * K) if self.use_graph else (self.Fn.launch_info().kernels_launched - k0)
        return C.max_over_ranks(e0.elapsed_time(e1)), wall, int(launches)

    def l2_policy(self) -> str:
        mb = self.footprint / 1e6
        if self.n_sets == 1:
            return "inputs+gradient %.0f MB per step > 2 x 126 MB L2, no flush needed" % mb
        return "inputs+gradient %.1f MB per step: steps rotate through %d buffer sets (%.0f MB > 2 x 126 MB L2 between reuses)" % (
            mb, self.n_sets, mb * self.n_sets)

    def close(self):
        for g in self.graphs:
            g.close()
        self.graphs = []


def quick_measure(C, X, z, t, n_global, K, warm, launch):
    """value / ms_per_step of a secondary configuration (no per-kernel passes)."""
    sh = LossShard(C, X, z, t, n_global, launch)
    for _ in range(max(warm, 3)):
        sh.step()
    C.barrier()
    ms, _, _ = sh.timed(K)
    out = {"value": n_global * K / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms / K, "steps": K,
           "launch": "cuda graph" if sh.use_graph else "direct", "buffer_sets": sh.n_sets, "loss": float(sh.last_report[0].item())}
    sh.close()
    return out


def parity_block(C, X, B, H, W, dtype, args):
    """One sharded step through the exchange path on the workload's own GLOBAL batch, checked before anything is timed:
    (a) the global loss of the sharded step against the single-shard evaluation of the gathered batch on rank 0 (2e-6),
    and every rank's gradient slice against that evaluation's; (b) rank 0's first image against the fp64 CPU oracle
    given the global sums (1e-5 fp32, 1e-2 bf16 storage)."""
    import numpy as np
    torch = C.torch
    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200 import functional as Fn
    from physics_informed_image_segmentation_b200.sharding import shard_bounds

    p = P.LossParams(**STAGE2)
    kind = Fn.X_LOGITS_SIGMOID
    b0, b1 = shard_bounds(B, C.rank, C.world)
    z, t = global_batch(B, H, W, C.dev, dtype, b0, b1)
    n_global = B * H * W
    sh = LossShard(C, X, z, t, n_global, "direct")
    sh.direct_step()
    torch.cuda.synchronize()
    rep_sh = sh.report.clone()
    grad_sh = sh.sets[0][2].float().clone() / float(C.world)  # the step scales by world size for DDP's averaging
    tot = (sh.tot if X.host is not None else (sh.sums + sh.sb)).clone() if C.distributed else (sh.sums + sh.sb).clone()
    sh.close()
    del sh
    out = {"checked": True, "world": C.world, "batch": [B, 1, H, W], "exchange": X.mode}
    ok = True
    tol_g = 2e-6 if dtype == torch.float32 else 1e-2
    # (a) single-shard evaluation of the whole batch on rank 0 (image by image gathered = regenerated, bit-identical)
    slices = [None] * C.world
    if C.distributed:
        mine = grad_sh.cpu()
        gathered = [None] * C.world
        C.dist.all_gather_object(gathered, (b0, b1, mine[:1].clone() if C.rank else mine))  # rank 0 keeps its own
        slices = gathered
    if C.rank == 0:
        zf, tf = global_batch(B, H, W, C.dev, dtype, 0, B)
        rep_f, sums_f, grad_f = Fn.loss_fwd_bwd(zf, tf, p, kind)
        torch.cuda.synchronize()
        e_loss = abs(rep_sh[0].item() - rep_f[0].item()) / abs(rep_f[0].item())
        gf = grad_f.float()
        den = gf.abs().max().item()
        e_grad = (grad_sh - gf[b0:b1]).abs().max().item() / den
        out.update({"loss_sharded": rep_sh[0].item(), "loss_single_shard": rep_f[0].item(), "loss_rel_err": e_loss,
                    "grad_slice_rel_err_rank0": e_grad})
        ok = ok and e_loss <= 2e-6 and e_grad <= tol_g
        if C.distributed:
            worst = 0.0
            for (a, b, g1) in slices[1:]:   # every other rank's first image
                worst = max(worst, (g1.to(C.dev) - gf[a:a + 1]).abs().max().item() / den)
            out["grad_first_image_rel_err_other_ranks"] = worst
            ok = ok and worst <= tol_g
        # (b) fp64 oracle on rank 0's first image, given the global sums
        from oracle import pil_oracle as po

        gs = tot.cpu().numpy().astype(np.float64)
        og = po.backward(zf[:1].float().cpu().numpy().astype(np.float64), tf[:1].float().cpu().numpy().astype(np.float64), po.STAGE2,
                         gs, int(n_global), po.X_LOGITS_SIGMOID)
        e_or = float(np.max(np.abs(grad_sh[:1].cpu().numpy() - og)) / np.max(np.abs(og)))
        comps = po.finalize(gs, int(n_global), po.STAGE2)
        e_or_loss = abs(rep_sh[0].item() - comps[0]) / abs(comps[0])
        out.update({"grad_vs_oracle_rel_err": e_or, "loss_vs_oracle_finalize_rel_err": e_or_loss,
                    "tolerances": {"loss": 2e-6, "grad_vs_single_shard": tol_g, "grad_vs_oracle": 1e-5 if dtype == torch.float32 else 1e-2}})
        ok = ok and e_or <= (1e-5 if dtype == torch.float32 else 1e-2) and e_or_loss <= 1e-5
        del zf, tf, grad_f, gf
        torch.cuda.empty_cache()
    out["ok"] = C.all_ok(ok)
    if not out["ok"]:
        if C.rank == 0:
            print("PARITY CHECK FAILED: " + json.dumps(out), file=sys.stderr)
        raise SystemExit(3)
    return out


def numa_local_affinity(C):
    """Best effort: run this rank on the CPUs next to its GPU before pinned staging memory is allocated (first touch
    places the pages), so that N host-buffer sessions do not all stage through one NUMA node."""
    try:
        props = C.torch.cuda.get_device_properties(C.local_rank)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read().strip())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as exc:
        return {"numa_node": None, "note": str(exc)[:80]}


def run_loss(args, C: Ctx):
    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200 import functional as Fn
    from physics_informed_image_segmentation_b200.sharding import shard_bounds

    torch = C.torch
    B, H, W, name = WORKLOADS[args.workload]
    dtype = torch.float32 if args.dtype == "f32" else torch.bfloat16
    esz = 4 if args.dtype == "f32" else 2
    wl_name = f"{name}_{'fp32' if args.dtype == 'f32' else 'bf16'}"
    # the default series (N = 1, 2, 4, 8 on the same GLOBAL batch) is a strong-scaling series; at N = 1 both modes coincide
    scaling = "weak" if args.scaling == "weak" else "strong"
    X = Exchanges(C, args.exchange)

    sampler = ClockSampler(C.local_rank)
    if C.rank == 0:
        sampler.start()  # runs through warm-up and all timed regions

    parity = None
    if not args.no_parity:
        pB = B if B * H * W <= 64 * 1024 * 1024 or C.distributed else 64 * 1024 * 1024 // (H * W)  # cfg5 on one GPU: a 16-image batch
        parity = parity_block(C, X, max(pB, C.world), H, W, dtype, args)

    # ---- the workload's shard on this rank ---------------------------------------------------------------------
    if scaling == "strong":
        b0, b1 = shard_bounds(B, C.rank, C.world)
        z, t = global_batch(B, H, W, C.dev, dtype, b0, b1)
        n_global = B * H * W
        Bg = B
    else:
        z, t = synth(B, H, W, 1234 + C.rank, C.dev, dtype)
        n_global = B * H * W * C.world
        Bg = B * C.world
    sh = LossShard(C, X, z, t, n_global, args.launch)
    for _ in range(max(args.warmup, 3)):
        sh.step()
        sh.direct_step()
    C.barrier()

    K = args.steps
    KA = max(3, min(K, 30))
    fwd_burst, bwd_burst = sh.instrumented(KA)                       # burst regime, like the measured copy peak
    total_ms, wall, launches = sh.timed(K)                           # the headline region
    K2 = max(K, min(20000, int(0.3 / max(total_ms * 1e-3 / K, 1e-6)) + 1))
    K2 = int(C.max_over_ranks(float(K2)))                            # every rank runs the same number of exchange steps
    fwd_sus, bwd_sus = sh.instrumented(K2)                           # >= 0.3 s: clocks settle under the power cap
    value = n_global * K / (total_ms * 1e-3) / 1e9
    loss_val = float(sh.last_report[0].item())
    info = Fn.launch_info()

    # ---- secondary configurations measured in the same run ------------------------------------------------------
    extras = {}
    if not args.no_extras:
        Kx = max(5, min(K, 50))
        if C.distributed and scaling == "strong":
            zw, tw = synth(B, H, W, 1234 + C.rank, C.dev, dtype)
            extras["weak_scaling"] = dict(quick_measure(C, X, zw, tw, B * H * W * C.world, Kx, 3, args.launch), per_gpu_batch=[B, 1, H, W])
            del zw, tw
        if args.workload == "cfg3":
            B5, H5, W5, n5 = WORKLOADS["cfg5"]
            for dname, dt5 in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
                a5, e5 = shard_bounds(B5, C.rank, C.world)
                torch.cuda.empty_cache()
                z5, t5 = global_batch(B5, H5, W5, C.dev, dt5, a5, e5)
                extras[f"cfg5_strong_{dname}"] = dict(quick_measure(C, X, z5, t5, B5 * H5 * W5, max(5, min(K, 20)), 3, args.launch),
                                                      workload=f"{n5}_{dname}", global_batch=[B5, 1, H5, W5], scaling="strong")
                del z5, t5
            torch.cuda.empty_cache()
        if not C.distributed and args.workload == "cfg3" and args.dtype == "f32":
            small = {}
            X0 = Exchanges(C, "none")
            for wn in ("ref128", "cfg1", "cfg2"):
                Bs, Hs, Ws, ns = WORKLOADS[wn]
                zs, ts = synth(Bs, Hs, Ws, 1234, C.dev, dtype)
                r = {}
                for mode in ("direct", "graph"):
                    q = quick_measure(C, X0, zs, ts, Bs * Hs * Ws, 500, 10, mode)
                    r[mode] = {"us_per_step": q["ms_per_step"] * 1e3, "value": q["value"], "buffer_sets": q["buffer_sets"]}
                small[ns] = r
            extras["small_batch"] = dict(small, note="one CUDA-graph launch per step (pil_step_graph_*) against two direct C-ABI calls; "
                                                       "8x1x128x128 is the reference's real batch shape (src/dataset.py:18)")
            # the module API a user calls: criterion.forward_logits(x, t); loss.backward()
            crit = P.DiceBCEPDELoss(**{k: v for k, v in STAGE2.items()}).to(C.dev)
            xm = z.detach().clone().requires_grad_(True)
            for _ in range(3):
                xm.grad = None
                crit.forward_logits(xm, t).backward()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(Kx):
                xm.grad = None
                crit.forward_logits(xm, t).backward()
            e1.record()
            torch.cuda.synchronize()
            ms_m = e0.elapsed_time(e1) / Kx
            extras["module_api"] = {"value": B * H * W / (ms_m * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_m, "steps": Kx,
                                    "call": "loss = DiceBCEPDELoss.forward_logits(x, t); loss.backward()  (autograd, fresh gradient buffer per step)"}
            del xm
            # the model tail: 64-channel features -> 1x1 output convolution -> activation -> loss -> backward,
            # fused (criterion.forward_features) against torch's convolution around the fused loss
            Bt, Ct = 8, 64
            feat = torch.randn(Bt, Ct, H, W, device=C.dev)
            conv = torch.nn.Conv2d(Ct, 1, 1).to(C.dev)
            tt8 = t[:Bt].contiguous()

            def tail_fused():
                f = feat.requires_grad_(True)
                f.grad = None
                conv.zero_grad(set_to_none=True)
                crit.forward_features(f, conv.weight, conv.bias, tt8).backward()

            def tail_unfused():
                f = feat.requires_grad_(True)
                f.grad = None
                conv.zero_grad(set_to_none=True)
                crit.forward_logits(conv(f), tt8).backward()

            tail = {}
            for nm, fn in (("fused", tail_fused), ("torch_conv_plus_fused_loss", tail_unfused)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                tail[nm] = {"ms_per_step": e0.elapsed_time(e1) / 10}
            px_t = Bt * H * W
            fb = 4 * Ct  # feature bytes per pixel
            tail["fused"]["algorithmic_bytes_per_pixel"] = fb + 4 + 4 + 12 + 4 + 2 * fb   # T1: feat + t + z | K2: z, t, g | T2: g + feat + dfeat
            tail["fused"]["achieved_gbs"] = tail["fused"]["algorithmic_bytes_per_pixel"] * px_t / (tail["fused"]["ms_per_step"] * 1e-3) / 1e9
            tail["note"] = (f"{Bt}x{Ct}x{H}x{W} fp32 features; both legs pay the three feature passes ({3 * fb} B/px); fusing the tail saves "
                            "the logits round trips (~20 B/px) and the convolution's extra kernels, not feature traffic")
            extras["model_tail"] = tail
            del feat, conv, crit
            torch.cuda.empty_cache()

    # ---- end to end through the host-buffer C ABI --------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        numa = numa_local_affinity(C) if C.distributed else None
        zh = z.cpu().pin_memory()
        th = t.cpu().pin_memory()
        x_code = 0 if args.dtype == "f32" else 1
        Bl = z.shape[0]
        with P.HostSession(Bl, H, W, device=C.local_rank, x_dtype=x_code, t_dtype=x_code) as sess:
            Ke = max(3, min(K, 10))

            def timed_e2e(fn):
                for _ in range(2):
                    r = fn()
                C.barrier()
                t0 = time.perf_counter()
                for _ in range(Ke):
                    r = fn()
                dt_ = time.perf_counter() - t0
                return C.max_over_ranks(dt_), r

            if C.distributed and X.host is not None:
                # ONE global loss: every rank stages its shard, the kernels exchange their sums, every rank reads the report
                dt_dev, rep = timed_e2e(lambda: sess.run_sharded(zh, th, sh.p, X.host, n_global=n_global, grad_scale=float(C.world)))
                note = ("pil_session_run_xchg: every rank stages ITS shard from pinned host memory (chunked H2D overlapped with the "
                        "pointwise forward), the shard sums cross over peer memory, one backward per rank (gradient stays in HBM), "
                        "D2H of the GLOBAL loss report on every rank; host wall clock, max over ranks")
                dt_host = rep_h = None
            else:
                dt_dev, rep = timed_e2e(lambda: sess.run(zh, th, sh.p, grad_on_device=True))
                note = ("pil_session_run_ex(PIL_SESSION_GRAD_ON_DEVICE): pinned host x,t -> H2D (chunked, overlapped with the pointwise "
                        "forward) -> backward over the batch (gradient stays in HBM, as a training step consumes it) -> D2H of the "
                        "loss report; host wall clock, PCIe-bound" + ("; per-rank session (independent shard losses, NCCL mode)" if C.distributed else ""))
                gh = torch.empty_like(zh).pin_memory()
                dt_host, rep_h = timed_e2e(lambda: sess.run(zh, th, sh.p, grad_host=gh))
        e2e = {"value": n_global * Ke / dt_dev / 1e9, "unit": UNIT, "h2d_bytes_per_step": 2 * n_global * esz,
               "d2h_bytes_per_step": 32 * C.world, "steps": Ke, "ms_per_step": 1e3 * dt_dev / Ke, "note": note, "loss": float(rep[0])}
        if numa is not None:
            e2e["staging"] = numa
        if dt_host is not None:
            e2e["with_gradient_d2h"] = {"value": n_global * Ke / dt_host / 1e9, "ms_per_step": 1e3 * dt_host / Ke,
                                        "d2h_bytes_per_step": n_global * esz + 32 * C.world, "loss": float(rep_h[0])}

    clocks = sampler.stop() if C.rank == 0 else None
    xchg_timeout = X.timed_out()
    if C.rank == 0:
        peak, peak_src = measured_hbm_peak()
        n_local = sh.n_local
        fwd_med, bwd_med = statistics.median(fwd_burst), statistics.median(bwd_burst)
        fwd_s, bwd_s = statistics.median(fwd_sus), statistics.median(bwd_sus)
        bpp_f, bpp_b = 2 * esz, 3 * esz
        ach_b = bpp_b * n_local / (bwd_med * 1e-3) / 1e9
        ach_f = bpp_f * n_local / (fwd_med * 1e-3) / 1e9
        ach_step = (bpp_f + bpp_b) * n_global / C.world * K / (total_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch of the backward kernel from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            if tj.get("workload") == wl_name and not C.distributed:
                kb = tj["kernels"]["pil_bwd_kernel"]
                traffic, traffic_src = kb["dram_bytes_read"] + kb["dram_bytes_write"], tj.get("source")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": C.world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": wl_name, "global_batch": [Bg, 1, H, W], "per_gpu_images": int(z.shape[0]),
                       "stage2_params": STAGE2, "entry": "logits (sigmoid fused)",
                       "parallelism": f"dp{C.world} ({scaling} scaling; batch shards by whole images; exchange of 8 doubles between the two kernels: {X.mode})",
                       "l2_policy": sh.l2_policy(),
                       "step": "pil_forward_pointwise -> pil_backward_accumulate (stencils evaluated once per step)",
                       "launch": "one CUDA-graph launch per step (pil_step_graph_*)" if sh.use_graph else "two direct C-ABI calls per step",
                       "tiling": {"fwd_blocks": info.fwd_blocks, "bwd_blocks": info.bwd_blocks, "bwd_rows_per_range": info.bwd_rows_per_segment,
                                  "bwd_row_staging": {0: "scalar loads", 1: "cp.async ring", 2: "TMA boxes + mbarrier"}.get(info.bwd_aligned)}},
            "roofline": {"bound": "hbm", "kernel": "pil_bwd_kernel (gradient + stencil sums)", "achieved": ach_b, "peak": peak, "unit": "GB/s",
                         "frac": ach_b / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bpp_b * n_local, "kernel_ms": bwd_med, "kernel_launches_timed": KA,
                         "sustained": {"kernel_ms": bwd_s, "launches": K2, "frac": bpp_b * n_local / (bwd_s * 1e-3) / 1e9 / peak},
                         "note": "kernel_ms: median CUDA-event interval around the launch in a short pass right after warm-up (an event "
                                 "between the two kernels of every step; burst regime, like the measured copy peak); sustained: the same "
                                 "over a >= 0.3 s pass after the headline region (clocks settle under the power cap); "
                                 + ("multi-GPU: the interval includes the wait for the other ranks' sums" if C.distributed else "single GPU")},
            "roofline_fwd": {"kernel": "pil_point_kernel (pointwise sums)", "achieved": ach_f, "frac": ach_f / peak, "kernel_ms": fwd_med,
                             "sustained_kernel_ms": fwd_s, "algorithmic_bytes_per_launch": bpp_f * n_local},
            "roofline_step": {"achieved": ach_step, "frac": ach_step / peak, "bytes_per_pixel": bpp_f + bpp_b, "per": "GPU"},
            "gpu_launches": int(launches), "clocks": clocks, "loss": loss_val, "wall_ms_per_step": 1e3 * wall / K,
            "exchange_timeout": xchg_timeout,
        }
        if parity is not None:
            line["parity"] = parity
        if extras:
            line["extras"] = extras
        if e2e is not None:
            line["e2e"] = e2e
        if C.world == 1 and not args.no_cpu and args.dtype == "f32":
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
        if C.world == 1 and not args.no_cpu:
            sample_B = max(1, min(B, (8 * 1024 * 1024) // (H * W)))
            npx, times, cores = cpu_port_time(sample_B, H, W, iters=3, warm=1)
            line["cpu_baseline"] = {"value": npx / statistics.median(times) / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample_B}x1x{H}x{W} (of {B}x1x{H}x{W}), median of 3 after 1 warm-up, "
                                              f"torch {torch.__version__} CPU ops, op-for-op port of the reference loss "
                                              "(the GPU box has no reference checkout; the port is pinned bit-identical to the real "
                                              "reference in the build container, tests/test_oracle.py)"}
        emit_line(line)
    sh.close()
    X.close()


def run_sweep(args, C: Ctx):
    """BASELINE config 4: the S2 (6 diffusion coefficients) and S3 (5 epsilons) sensitivity grids of run_ablation.py:159-224
    as ONE batched loss evaluation per step: a single pass over the sharded maps (pil_forward_moments), an all-reduce of
    the 16 moment sums, 11 loss reports from closed forms (pil_sweep_finalize)."""
    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200 import functional as Fn
    from physics_informed_image_segmentation_b200.sharding import shard_bounds

    torch = C.torch
    B, H, W, name = WORKLOADS["cfg4"]
    dtype = torch.float32 if args.dtype == "f32" else torch.bfloat16
    esz = 4 if args.dtype == "f32" else 2
    b0, b1 = shard_bounds(B, C.rank, C.world)
    z, t = global_batch(B, H, W, C.dev, dtype, b0, b1)
    grid = P.s2_grid() + P.s3_grid()
    n_local, n_global = z.numel(), B * H * W
    footprint = 2 * n_local * esz
    n_sets = 1 if footprint >= 2 * L2_BYTES else min(64, -(-2 * L2_BYTES // footprint) + 1)
    sets = [(z, t)] + [(z.clone(), t.clone()) for _ in range(n_sets - 1)]
    moments = torch.empty(16, dtype=torch.float64, device=C.dev)
    reports = torch.empty(len(grid), 8, dtype=torch.float32, device=C.dev)
    kind = Fn.X_LOGITS_SIGMOID
    state = {"k": 0}
    X = Exchanges(C, args.exchange)

    def step(events=None):
        zz, tt = sets[state["k"] % n_sets]
        state["k"] += 1
        ex = X.host.next_step() if X.host is not None else None
        if events:
            events[0].record()
        Fn.forward_moments(zz, tt, kind, moments=moments, ex=ex)
        if events:
            events[1].record()
        if ex is not None:       # the moment sums crossed over peer memory inside the kernel: no collective call
            Fn.sweep_finalize_xchg(ex, n_global, grid, C.dev, reports=reports)
        else:
            if C.distributed:
                C.dist.all_reduce(moments)
            Fn.sweep_finalize(moments, n_global, grid, reports=reports)

    # timed steps: one CUDA-graph launch each (pil_sweep_graph_*) for shards up to 16 Mpixel -- a sharded pass takes a few
    # tens of microseconds, less than the host needs for two calls (8 ranks: 85 us per step direct against the 38 us the
    # kernel takes); the per-kernel event pass above stays on direct calls
    use_graph = args.launch == "graph" or (args.launch == "auto" and n_local <= (16 << 20))
    if C.distributed and X.device is None:
        use_graph = False
    graphs = []
    if use_graph:
        graphs = [Fn.SweepGraph(zz, tt, kind, grid, exchange=X.device, n_global=n_global) for zz, tt in sets]

    def timed_step():
        if graphs:
            state["k"] += 1
            graphs[state["k"] % n_sets].launch()
        else:
            step()

    sampler = ClockSampler(C.local_rank)
    if C.rank == 0:
        sampler.start()
    # parity before timing: the 11 losses against the fp64 CPU oracle evaluated setting by setting on a small batch
    parity = None
    if not args.no_parity and C.rank == 0:
        import numpy as np
        from oracle import pil_oracle as po

        zs, ts = global_batch(2, 256, 256, C.dev, torch.float32, 0, 2)
        rs = Fn.sweep_finalize(Fn.forward_moments(zs, ts, kind), -1, grid).cpu().numpy()
        worst = 0.0
        for k, gp in enumerate(grid):
            pp = po.Params(dice_weight=gp.dice_weight, bce_weight=gp.bce_weight, pde_weight=gp.pde_weight, phase_field_weight=gp.phase_field_weight,
                           diffusion_coeff=gp.diffusion_coeff, reaction_threshold=gp.reaction_threshold, epsilon=gp.epsilon, smooth=gp.smooth)
            s = po.sums(zs.cpu().numpy().astype(np.float64), ts.cpu().numpy().astype(np.float64), pp, po.X_LOGITS_SIGMOID)
            worst = max(worst, abs(rs[k, 0] - po.finalize(s, int(s[7]), pp)[0]) / abs(po.finalize(s, int(s[7]), pp)[0]))
        parity = {"checked": True, "settings": len(grid), "max_loss_rel_err_vs_oracle": float(worst), "tolerance": 1e-5, "ok": bool(worst <= 1e-5)}
        if not parity["ok"]:
            print("PARITY CHECK FAILED: " + json.dumps(parity), file=sys.stderr)
            raise SystemExit(3)
    for _ in range(max(args.warmup, 3)):
        step()
        timed_step()
    C.barrier()
    K = args.steps
    KA = max(3, min(K, 30))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(KA)]
    for k in range(KA):
        step(ev[k])
    torch.cuda.synchronize()
    kern_ms = statistics.median(e[0].elapsed_time(e[1]) for e in ev)
    k0 = Fn.launch_info().kernels_launched
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    C.barrier()
    e0.record()
    for _ in range(K):
        timed_step()
    e1.record()
    torch.cuda.synchronize()
    C.barrier()
    launches = (2 * K) if graphs else Fn.launch_info().kernels_launched - k0   # a graph launch replays the two kernels
    if graphs:
        reports = graphs[state["k"] % n_sets].reports
    total_ms = C.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if C.rank == 0 else None
    if C.rank == 0:
        peak, peak_src = measured_hbm_peak()
        ach = 2 * esz * n_local / (kern_ms * 1e-3) / 1e9
        value = n_global * K / (total_ms * 1e-3) / 1e9
        emit_line({
            "metric": "sweep_batched_loss_eval_gpixels_per_s", "value": value, "unit": UNIT, "n_gpus": C.world, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{name}_{'fp32' if args.dtype == 'f32' else 'bf16'}", "global_batch": [B, 1, H, W],
                       "per_gpu_images": int(z.shape[0]), "settings_per_pass": len(grid),
                       "grids": "S2: D in {0.5,1,2,5,10,100} (pde_weight 1e-3); S3: eps in {0.001,0.01,0.05,0.1,0.2} (both weights 1e-4, D 5)",
                       "step": ("pil_forward_moments_xchg (one pass over x, t; the shard's 16 sums pushed into every rank's mailbox) -> "
                                "pil_sweep_finalize_xchg (11 global loss reports)") if X.host is not None else
                               "pil_forward_moments (one pass over x, t) -> all-reduce of 16 doubles -> pil_sweep_finalize (11 loss reports)",
                       "launch": "one CUDA-graph launch per step (pil_sweep_graph_*)" if graphs else "two direct C-ABI calls per step",
                       "l2_policy": ("maps %.0f MB per step > 2 x 126 MB L2" % (footprint / 1e6)) if n_sets == 1 else
                                    ("maps %.1f MB per step: steps rotate through %d buffer sets" % (footprint / 1e6, n_sets))},
            "pixel_evaluations_per_s": value * len(grid) * 1e9,
            "roofline": {"bound": "hbm", "kernel": "pil_fwd_kernel<MOMENTS>", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": 2 * esz * n_local, "kernel_ms": kern_ms,
                         "kernel_launches_timed": KA},
            "gpu_launches": int(launches), "clocks": clocks, "losses": [float(v) for v in reports[:, 0].cpu()],
            "exchange_timeout": X.timed_out(), **({"parity": parity} if parity else {}),
        })
    for gsw in graphs:
        gsw.close()
    X.close()


def run_train_step(args, C: Ctx):
    """BASELINE config 3: U-Net-shaped forward (bf16 autocast, channels_last) -> fused Stage II loss on the fp32 logits of
    the whole shard -> backward -> AdamW, data parallel (DDP) over the global batch 64 x 1 x 1024 x 1024.  The loss is
    batch-global, so when a rank's shard is larger than --microbatch the model runs in two passes: a no-grad forward
    of every micro-batch collects the shard's logits, ONE fused loss evaluation yields dL/dlogits for the whole shard,
    and every micro-batch is then re-run with autograd and back-propagated from its slice (activation memory of one
    micro-batch; one extra model forward)."""
    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200.integration import use_logits_head
    from physics_informed_image_segmentation_b200.sharding import enable_peer_exchange, shard_bounds
    from tools.unet_harness import UNetHarness, n_params

    torch = C.torch
    B, H, W, name = WORKLOADS["cfg3_step"]
    if args.shape:
        B, H, W = (int(v) for v in args.shape.split("x"))
    b0, b1 = shard_bounds(B, C.rank, C.world)
    nb = b1 - b0
    g = torch.Generator(device=C.dev).manual_seed(99 + C.rank)
    images = torch.randn(nb, 1, H, W, device=C.dev, generator=g)
    masks = (torch.rand(nb, 1, H, W, device=C.dev, generator=g) > 0.5).float()
    torch.manual_seed(5)
    model = UNetHarness(64, "sigmoid").to(C.dev).to(memory_format=torch.channels_last)
    group = None
    net = model
    if C.distributed:
        from torch.nn.parallel import DistributedDataParallel as DDP

        group = C.dist.group.WORLD
        if args.exchange == "peer":
            try:
                enable_peer_exchange(C.dev, group)
            except Exception as exc:
                print(f"[rank {C.rank}] peer exchange unavailable ({exc}); NCCL all-reduce of the sums", file=sys.stderr)
        net = DDP(model, device_ids=[C.local_rank], gradient_as_bucket_view=True)
    crit = P.DiceBCEPDELoss(**STAGE2, process_group=group, ddp_average=True).to(C.dev)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-5, weight_decay=1e-5)   # src/train.py:720-726: 0.1 * lr, wd 1e-5
    mb = max(1, min(args.microbatch, nb))
    ev_loss = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def fwd(x):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return net(x.contiguous(memory_format=torch.channels_last)).float()

    def step(time_loss=False):
        opt.zero_grad(set_to_none=True)
        with use_logits_head(model):
            if nb <= mb:
                logits = fwd(images)
                if time_loss:
                    ev_loss[0].record()
                loss = crit.forward_logits(logits.contiguous(), masks)
                if time_loss:
                    ev_loss[1].record()
                loss.backward()
            else:
                with torch.no_grad():
                    logits = torch.cat([fwd(images[i:i + mb]) for i in range(0, nb, mb)]).contiguous()
                logits.requires_grad_(True)
                if time_loss:
                    ev_loss[0].record()
                loss = crit.forward_logits(logits, masks)
                loss.backward()                                            # free: the gradient was written by the forward's second kernel
                if time_loss:
                    ev_loss[1].record()
                gl = logits.grad
                chunks = list(range(0, nb, mb))
                for j, i in enumerate(chunks):
                    last = j == len(chunks) - 1
                    ctx = contextlib.nullcontext() if (last or not C.distributed) else net.no_sync()
                    with ctx:
                        fwd(images[i:i + mb]).backward(gl[i:i + mb])
        opt.step()
        return loss

    sampler = ClockSampler(C.local_rank)
    if C.rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 2)):
        step()
    C.barrier()
    K = args.steps
    loss_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    C.barrier()
    e0.record()
    for _ in range(K):
        loss = step(time_loss=True)
        ev_loss[1].synchronize()
        loss_ms.append(ev_loss[0].elapsed_time(ev_loss[1]))
    e1.record()
    torch.cuda.synchronize()
    C.barrier()
    total_ms = C.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if C.rank == 0 else None
    if C.rank == 0:
        n_global = B * H * W
        lm = statistics.median(loss_ms)
        emit_line({
            "metric": "unet_train_step_gpixels_per_s", "value": n_global * K / (total_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": C.world,
            "steps": K, "warmup": max(args.warmup, 2), "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16 model (autocast) + f32 loss", "data": "synthetic",
            "config": {"workload": name if not args.shape else f"unet_train_step_{B}x1x{H}x{W}", "global_batch": [B, 1, H, W],
                       "per_gpu_images": nb, "microbatch": mb, "model": f"U-Net-shaped harness, base 64, {n_params(model)} parameters, random init",
                       "optimizer": "AdamW(lr 1e-5, wd 1e-5)", "parallelism": f"ddp{C.world}", "stage2_params": STAGE2,
                       "step": "model forward (logits head) -> DiceBCEPDELoss.forward_logits on the shard (global loss over the exchange) -> "
                               "backward -> AdamW" + ("; two-pass micro-batching (one extra no-grad forward)" if nb > mb else "")},
            "loss_ms": lm, "loss_share_of_step": lm / (total_ms / K),
            "loss_note": "CUDA-event interval around the fused loss evaluation (both kernels + the exchange wait) on rank 0",
            "clocks": clocks, "loss": float(loss.item()), "gpu_launches": 2 * K,
        })
    if C.distributed:
        from physics_informed_image_segmentation_b200.sharding import disable_peer_exchange

        disable_peer_exchange()


def run_ours(args):
    C = Ctx()
    try:
        if args.workload == "cfg4":
            run_sweep(args, C)
        elif args.workload == "cfg3_step":
            run_train_step(args, C)
        else:
            run_loss(args, C)
    finally:
        C.close()


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
    ap.add_argument("--scaling", choices=["auto", "strong", "weak"], default="auto",
                    help="multi-GPU: strong = the global batch is sharded (default), weak = every rank holds the full batch")
    ap.add_argument("--launch", choices=["auto", "graph", "direct"], default="auto",
                    help="how the timed steps are issued: one CUDA-graph launch per step (default for shards up to 16 Mpixel) or two direct calls")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity block")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configurations (weak scaling, cfg5, small batches, module API)")
    ap.add_argument("--microbatch", type=int, default=8, help="cfg3_step: images per model pass")
    ap.add_argument("--shape", default="", help="cfg3_step: BxHxW instead of 64x1024x1024 (development)")
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
