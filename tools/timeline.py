"""Per-block timeline of the two hot kernels (development tool; needs a -DPIL_TIMELINE build):

    cd physics_informed_image_segmentation_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC \
         -shared -DPIL_TIMELINE -DPIL_DEV_F32_ONLY -I ../../include -I . -o dev/libpil_tl.so pil_unity.cu pil_session.cu
    PIL_LIB=physics_informed_image_segmentation_b200/csrc/dev/libpil_tl.so python tools/timeline.py [--workload cfg3 | --shape BxHxW] [--steps 4]

Every block stamps %globaltimer at entry, after the PDL wait, at the end of its main loop and at exit.
Prints, per kernel of the LAST step: start spread, loop-end spread, tail, and the gap/overlap between
consecutive kernels -- where the microseconds outside the steady state go."""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import physics_informed_image_segmentation_b200 as P  # noqa: E402
from physics_informed_image_segmentation_b200 import _lib, functional as Fn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3")
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--side-stream", action="store_true")
ap.add_argument("--shape", default="", help="BxHxW instead of a named workload")
ap.add_argument("--graph", action="store_true", help="steps are CUDA-graph launches (no host gaps); the stamps are those of the last one")
a = ap.parse_args()
B, H, W, name = bench.WORKLOADS[a.workload]
if a.shape:
    B, H, W = (int(v) for v in a.shape.split("x"))
    name = a.shape
dev = torch.device("cuda:0")
L = _lib.lib()
L.pil_debug_timeline.restype = ctypes.c_int
L.pil_debug_timeline.argtypes = [ctypes.c_void_p]
z, t = bench.synth(B, H, W, 1234, dev, torch.float32)
g = torch.empty_like(z)
p = P.LossParams(**bench.STAGE2)
sums = torch.empty(8, dtype=torch.float64, device=dev)
rep = torch.empty(8, dtype=torch.float32, device=dev)
sb = torch.empty(8, dtype=torch.float64, device=dev)
tl = torch.zeros(a.steps, 2, 4096, 8, dtype=torch.int64, device=dev)
stream = torch.cuda.Stream() if a.side_stream else torch.cuda.current_stream()
with torch.cuda.stream(stream):
    for _ in range(3):
        Fn.forward_pointwise(z, t, p, 1, sums=sums)
        Fn.backward_accumulate(z, t, p, 1, sums, z.numel(), out=g, stencil_sums=sb, report=rep)
    torch.cuda.synchronize()
    if a.graph:
        sg = Fn.StepGraph(z, t, p, 1, grad=g)
        for _ in range(3):
            sg.launch()
        assert L.pil_debug_timeline(tl[a.steps - 1].data_ptr()) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_rep = 50
        e0.record()
        for _ in range(n_rep):
            sg.launch()
        e1.record()
        torch.cuda.synchronize()
        print(f"graph launches: {e0.elapsed_time(e1) * 1e3 / n_rep:.1f} us per step")
    # one buffer per step, the pointer is switched by a tiny async symbol copy between steps
    for k in range(0 if a.graph else a.steps):
        assert L.pil_debug_timeline(tl[k].data_ptr()) == 0
        Fn.forward_pointwise(z, t, p, 1, sums=sums)
        Fn.backward_accumulate(z, t, p, 1, sums, z.numel(), out=g, stencil_sums=sb, report=rep)
    torch.cuda.synchronize()
info = Fn.launch_info()
T = tl.cpu().numpy().astype(np.int64)
nb = [info.fwd_blocks, info.bwd_blocks]
k = a.steps - 1
t0 = T[k, 0, :nb[0], 0].min()
print(f"{name}: blocks fwd {nb[0]} bwd {nb[1]}; times in us relative to the first forward block of step {k}")
names = ["pointwise fwd", "backward"]
for ker in range(2):
    X = (T[k, ker, :nb[ker], :4] - t0) / 1e3
    ent, wait, loop, ex = X[:, 0], X[:, 1], X[:, 2], X[:, 3]
    print(f"  {names[ker]:14s} entry {ent.min():7.1f}..{ent.max():7.1f} | after wait {wait.min():7.1f}..{wait.max():7.1f} | "
          f"loop end {loop.min():7.1f}..{loop.max():7.1f} (p50 {np.median(loop):7.1f}, p95 {np.percentile(loop, 95):7.1f}) | exit max {ex.max():7.1f}")
    if ker == 1 and X.shape[0] and T[k, ker, :nb[ker], 4].max() > 0:
        # stamps 4: every warp of the block arrived (the LAST block re-stamps it when all partials are added),
        # 5: ticket taken + partial sent, 6: loss assembled (last block only)
        lb = int(np.argmax(T[k, ker, :nb[ker], 3]))
        Y = (T[k, ker, :nb[ker], 4:6] - t0) / 1e3
        oth = np.arange(nb[ker]) != lb
        if oth.any():
            print(f"      epilogue: block arrived -> ticket taken, partial sent +{np.median((Y[:, 1] - Y[:, 0])[oth]):.2f} us "
                  f"(max +{(Y[:, 1] - Y[:, 0])[oth].max():.2f})")
        Z = (T[k, ker, lb, :8] - t0) / 1e3
        print(f"      last block {lb}: loop end {Z[2]:.1f} | ticket taken {Z[5]:.1f} | partials added {Z[4]:.1f} | loss assembled {Z[6]:.1f} | exit {Z[3]:.1f}")
    smid = T[k, ker, :nb[ker], 7]
    dur = loop - wait
    print(f"      main-loop duration per block: min {dur.min():.1f} p50 {np.median(dur):.1f} p95 {np.percentile(dur, 95):.1f} max {dur.max():.1f} us; "
          f"{len(np.unique(smid))} SMs")
if a.steps > 1 and not a.graph:
    prev_exit = (T[k - 1, 1, :nb[1], 3].max() - t0) / 1e3
    print(f"  previous step's backward exit at {prev_exit:.1f} us; step period {(t0 - T[k - 1, 0, :nb[0], 0].min()) / 1e3:.1f} us")
