"""Run a few fwd+bwd steps of the Stage II loss on resident tensors (for ncu / quick timing).

    python tools/prof_step.py [--workload cfg3] [--steps 3] [--dtype f32] [--rps-fwd N --rps-bwd N]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import physics_informed_image_segmentation_b200 as P  # noqa: E402
from physics_informed_image_segmentation_b200 import _lib, functional as Fn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--rps-fwd", type=int, default=0)
ap.add_argument("--rps-bwd", type=int, default=0)
ap.add_argument("--kind", type=int, default=1)
ap.add_argument("--shape", default="")
ap.add_argument("--mode", default="split", choices=["split", "full", "fused"])
ap.add_argument("--rotate", type=int, default=1, help="fused mode: cycle through this many gradient buffers (no L2 carry-over of dirty lines)")
a = ap.parse_args()
B, H, W, name = bench.WORKLOADS[a.workload]
if a.shape:
    B, H, W = (int(v) for v in a.shape.split("x"))
    name = a.shape
dev = torch.device("cuda:0")
dt = torch.float32 if a.dtype == "f32" else torch.bfloat16
z, t = bench.synth(B, H, W, 1234, dev, dt)
if a.kind == 0:
    z = torch.sigmoid(z.float()).to(dt)
g = torch.empty_like(z)
p = P.LossParams(**bench.STAGE2)
sums = torch.empty(8, dtype=torch.float64, device=dev)
rep = torch.empty(8, dtype=torch.float32, device=dev)
sb = torch.empty(8, dtype=torch.float64, device=dev)
_lib.lib().pil_set_tuning(a.rps_fwd, a.rps_bwd)
n_it = a.steps + 3
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_it)]
torch.cuda.synchronize()
for k in range(n_it):  # no host sync inside: the CPU runs ahead, intervals are pure GPU time
    ev[k][0].record()
    if a.mode == "full":
        Fn.forward_sums(z, t, p, a.kind, sums=sums, report=rep)
        ev[k][1].record()
        Fn.backward_grad(z, t, p, a.kind, sums, z.numel(), out=g)
    else:
        Fn.forward_pointwise(z, t, p, a.kind, sums=sums)
        ev[k][1].record()
        Fn.backward_accumulate(z, t, p, a.kind, sums, z.numel(), out=g, stencil_sums=sb, report=rep)
    ev[k][2].record()
torch.cuda.synchronize()
fm = [ev[k][0].elapsed_time(ev[k][1]) for k in range(3, n_it)]
bm = [ev[k][1].elapsed_time(ev[k][2]) for k in range(3, n_it)]
if a.mode == "fused":
    # the single C-ABI call, no events between the two kernels: per-step time over the whole loop
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        Fn.loss_fwd_bwd(z, t, p, a.kind, grad=g, sums=sums, report=rep)
    torch.cuda.synchronize()
    gs = [g] + [torch.empty_like(g) for _ in range(a.rotate - 1)]
    e0.record()
    for k in range(a.steps):
        Fn.loss_fwd_bwd(z, t, p, a.kind, grad=gs[k % a.rotate], sums=sums, report=rep)
    e1.record()
    torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / a.steps
    print(f"{name} {a.dtype} fused: {per*1e3:.1f} us/step  {z.numel()/per/1e6:.1f} Gpx/s  {5*z.element_size()*z.numel()/per/1e6:.0f} GB/s loss {rep[0].item():.6f}")
    sys.exit(0)
info = Fn.launch_info()
n = z.numel()
esz = z.element_size()
f, b = min(fm), min(bm)
print(f"{name} {a.dtype} {a.mode} kind={a.kind} rps=({info.fwd_rows_per_segment},{info.bwd_rows_per_segment}) blocks=({info.fwd_blocks},{info.bwd_blocks}) "
      f"fwd {f*1e3:.1f} us {2*esz*n/f/1e6:.0f} GB/s | bwd {b*1e3:.1f} us {3*esz*n/b/1e6:.0f} GB/s | "
      f"step {n/(f+b)/1e6:.1f} Gpx/s {5*esz*n/(f+b)/1e6:.0f} GB/s loss {rep[0].item():.6f}")
