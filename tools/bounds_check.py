"""Memory-safety check of the fused kernels with a -DPIL_BOUNDS build (every global read / cp.async source /
store is checked against the extents of the call's tensors; compute-sanitizer is not available on the pool):

    cd physics_informed_image_segmentation_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC \
         -shared -DPIL_BOUNDS -DPIL_DEV_F32_ONLY -I ../../include -I . -o dev/libpil_bounds.so pil_unity.cu pil_session.cu \
         pil_graph.cu pil_boundary.cu
    PIL_LIB=physics_informed_image_segmentation_b200/csrc/dev/libpil_bounds.so python tools/bounds_check.py
    (add PIL_BWD_STAGE=cpasync PIL_FWD_STAGE=cpasync for the cp.async ring; the TMA boxes are bounds-safe by construction --
    the TMA unit zero-fills what lies outside the tensor map -- and their direct loads, stores and results are checked)

The tensors are carved out of the MIDDLE of a larger allocation, so an access that strays outside them lands in
poisoned (NaN) memory of the same allocation instead of faulting -- it is counted, and would also corrupt the sums."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import physics_informed_image_segmentation_b200 as P  # noqa: E402
from physics_informed_image_segmentation_b200 import _lib, functional as Fn  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
if not hasattr(L, "pil_debug_bounds"):
    raise SystemExit("needs a -DPIL_BOUNDS build: set PIL_LIB (see the docstring)")
L.pil_debug_bounds.restype = ctypes.c_int
L.pil_debug_bounds.argtypes = [ctypes.c_void_p]


def errors():
    out = (ctypes.c_ulonglong * 4)()
    assert L.pil_debug_bounds(out) == 0
    return int(out[0]), int(out[1]), int(out[2])


def carve(B, H, W, offset_elems=0):
    """(B,1,H,W) view in the middle of a NaN-poisoned buffer"""
    n = B * H * W
    buf = torch.full((n + 4096 + offset_elems,), float("nan"), device=dev)
    v = buf[2048 + offset_elems: 2048 + offset_elems + n].view(B, 1, H, W)
    return buf, v


shapes = [(1, 2, 2), (2, 3, 5), (1, 7, 1001), (3, 127, 129), (2, 70, 1024), (5, 17, 36), (1, 33, 240), (4, 64, 248), (2, 256, 256),
          (3, 100, 120), (2, 64, 124), (1, 1024, 8), (6, 8, 2048),
          (48, 1024, 1000)]   # large enough for the automatic partition: dynamically claimed 48-row ranges + the 16-row tail phase
p = P.LossParams(pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0)
errors()
bad = 0
g = torch.Generator(device=dev).manual_seed(3)
for (B, H, W) in shapes:
    for off in (0, 1):          # 16-byte aligned base, and a base that forces the scalar path
        for kind in (0, 1):
            for rows in (0, 8, 13):   # static partition, and short dynamically claimed ranges
                bx, x = carve(B, H, W, off)
                bt, t = carve(B, H, W, off)
                bg, gr = carve(B, H, W, off)
                z = 2.0 * torch.randn(B, 1, H, W, device=dev, generator=g)
                x.copy_(torch.sigmoid(z) if kind == 0 else z)
                t.copy_((torch.rand(B, 1, H, W, device=dev, generator=g) > 0.5).float())
                L.pil_set_tuning(rows, rows)
                sums, rep = Fn.forward_sums(x, t, p, kind)
                Fn.backward_grad(x, t, p, kind, sums, x.numel(), out=gr)
                rep2, sums2, _ = Fn.loss_fwd_bwd(x, t, p, kind, grad=gr)
                Fn.forward_moments(x, t, kind)
                Fn.forward_pointwise_metrics(x, t, p, kind, 0.5)
                L.pil_set_tuning(0, 0)
                r, w, first = errors()
                finite = bool(torch.isfinite(rep).all() and torch.isfinite(rep2).all() and torch.isfinite(gr).all())
                if r or w or not finite:
                    bad += 1
                    print(f"shape {B}x{H}x{W} offset {off} kind {kind} rows {rows}: bad reads {r} bad writes {w} first {first:#x} finite {finite}")
print(f"bounds check: {len(shapes) * 2 * 2 * 3} configurations, {bad} with out-of-extent accesses or non-finite results")
sys.exit(1 if bad else 0)
