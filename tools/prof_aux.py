"""Run the auxiliary kernels a few times (for ncu): full forward, moments forward (sweep), pointwise forward
with per-image metrics counts.    python tools/prof_aux.py [--workload cfg3] [--steps 3]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import physics_informed_image_segmentation_b200 as P  # noqa: E402
from physics_informed_image_segmentation_b200 import functional as Fn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3")
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
B, H, W, name = bench.WORKLOADS[a.workload]
dev = torch.device("cuda:0")
z, t = bench.synth(B, H, W, 1234, dev, torch.float32)
p = P.LossParams(**bench.STAGE2)
grid = P.s2_grid() + P.s3_grid()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for k in range(a.steps + 2):
    ev[0].record()
    Fn.forward_sums(z, t, p, 1)
    ev[1].record()
    rep = P.sweep_losses(z, t, grid, activation="sigmoid")
    ev[2].record()
    sums, counts = Fn.forward_pointwise_metrics(z, t, p, 1, 0.5)
    dice, iou = Fn.image_metrics(counts)
    ev[3].record()
torch.cuda.synchronize()
n = z.numel()
print(f"{name}: full forward {ev[0].elapsed_time(ev[1])*1e3:.1f} us ({8*n/ev[0].elapsed_time(ev[1])/1e6:.0f} GB/s) | "
      f"sweep of {len(grid)} settings {ev[1].elapsed_time(ev[2])*1e3:.1f} us | pointwise+metrics {ev[2].elapsed_time(ev[3])*1e3:.1f} us "
      f"(dice[0] {dice[0].item():.4f}, iou[0] {iou[0].item():.4f}, sweep loss[0] {rep[0, 0].item():.6f})")
