"""Small-batch step cost: the two direct C-ABI calls against one CUDA-graph launch (functional.StepGraph).

    python tools/graph_bench.py [--shapes 8x128x128,8x256x256,32x512x512,8x1024x1024] [--steps 2000]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import physics_informed_image_segmentation_b200 as P  # noqa: E402
from physics_informed_image_segmentation_b200 import functional as Fn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="8x128x128,8x256x256,32x512x512,8x1024x1024")
ap.add_argument("--steps", type=int, default=2000)
a = ap.parse_args()
dev = torch.device("cuda:0")
p = P.LossParams(**bench.STAGE2)
for shp in a.shapes.split(","):
    B, H, W = (int(v) for v in shp.split("x"))
    z, t = bench.synth(B, H, W, 1234, dev, torch.float32)
    grad = torch.empty_like(z)
    sums = torch.empty(8, dtype=torch.float64, device=dev)
    rep = torch.empty(8, dtype=torch.float32, device=dev)
    g = Fn.StepGraph(z, t, p, Fn.X_LOGITS_SIGMOID, grad=grad)

    def timed(fn):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record()
        host = (time.perf_counter() - w0) / a.steps * 1e6
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps * 1e3, host

    d_us, d_host = timed(lambda: Fn.loss_fwd_bwd(z, t, p, Fn.X_LOGITS_SIGMOID, grad=grad, sums=sums, report=rep))
    g_us, g_host = timed(g.launch)
    print(f"{shp:>14s}: direct {d_us:6.1f} us/step (host {d_host:5.1f}) | graph {g_us:6.1f} us/step (host {g_host:5.1f}) | "
          f"{B * H * W / g_us / 1e3:.1f} Gpx/s with the graph")
    g.close()
