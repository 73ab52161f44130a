"""CPU restatement of the reference's boundary-F1 metric.  TEST INFRASTRUCTURE ONLY: imported by tests/ (and by
tests/golden/make_golden_bf1.py, which pins it against the real reference); never by the product package.

Follows reference src/evaluate.py:
  extract_boundaries  :102-122   cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) + cv2.drawContours(thickness 1)
  compute_boundary_f1 :125-193   two cv2.distanceTransform(1 - boundary, DIST_L2, 5) <= tolerance, float32 ratios
  compute_boundary_f1_batch :196-229

OpenCV is a third-party dependency that is not vendored in the reference; its two algorithms are restated here from
their published definitions and pinned against cv2 itself in tests/test_boundary_oracle.py (cv2 is installed in the
build container) and against outputs of the real reference in tests/golden/ref_bf1.npz:
  * Suzuki & Abe border following with 8-connected foreground: the outer border of a component consists of its pixels
    that have a pixel of the SURROUNDING background component in their 4-neighbourhood; RETR_EXTERNAL keeps only borders
    whose surrounding component is the frame's (the outside of the image counts as frame background, 4-connected).
  * Borgefors 5x5 chamfer distance with OpenCV's DIST_L2 weights in 16-bit fixed point: 65536, 91750, 143976 for the
    (1,0), (1,1), (2,1) steps; computed here by the classical two raster passes.
"""
from __future__ import annotations

import numpy as np

HV, DIAG, LONG = 65536, 91750, 143976   # cvRound(w * 2**16) for w = 1.0f, 1.4f, 2.1969f (OpenCV distanceTransform_5x5)
_BIG = 1 << 40


def frame_connected_background(mask: np.ndarray) -> np.ndarray:
    """Background pixels 4-connected to the image frame (flood fill on the image padded by one ring of background)."""
    H, W = mask.shape
    pad = np.zeros((H + 2, W + 2), dtype=bool)
    pad[1:-1, 1:-1] = mask
    seen = np.zeros_like(pad)
    stack = [(0, 0)]
    seen[0, 0] = True
    while stack:
        r, c = stack.pop()
        for dr, dc in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            rr, cc = r + dr, c + dc
            if 0 <= rr < H + 2 and 0 <= cc < W + 2 and not seen[rr, cc] and not pad[rr, cc]:
                seen[rr, cc] = True
                stack.append((rr, cc))
    return seen  # padded (H+2, W+2)


def extract_boundaries(mask: np.ndarray) -> np.ndarray:
    """src/evaluate.py:102-122 without OpenCV.  mask: (H, W) float/bool; foreground = uint8(mask * 255) != 0."""
    fg = (np.asarray(mask, dtype=np.float32) * 255).astype(np.uint8) != 0
    out = frame_connected_background(fg)
    near = out[:-2, 1:-1] | out[2:, 1:-1] | out[1:-1, :-2] | out[1:-1, 2:]
    return (fg & near).astype(np.float32)


def chamfer5_fixed(zero_mask: np.ndarray) -> np.ndarray:
    """Fixed-point 5x5 chamfer distance to the nearest True pixel of zero_mask (two raster passes, paths stay inside
    the image, as in OpenCV's bordered buffer).  int64 array; _BIG where there is no True pixel at all."""
    H, W = zero_mask.shape
    d = np.where(zero_mask, 0, _BIG).astype(np.int64)
    fwd = ((-1, 0, HV), (0, -1, HV), (-1, -1, DIAG), (-1, 1, DIAG), (-2, -1, LONG), (-2, 1, LONG), (-1, -2, LONG), (-1, 2, LONG))
    for r in range(H):
        for c in range(W):
            v = d[r, c]
            for dr, dc, w in fwd:
                rr, cc = r + dr, c + dc
                if 0 <= rr < H and 0 <= cc < W and d[rr, cc] + w < v:
                    v = d[rr, cc] + w
            d[r, c] = v
    for r in range(H - 1, -1, -1):
        for c in range(W - 1, -1, -1):
            v = d[r, c]
            for dr, dc, w in fwd:
                rr, cc = r - dr, c - dc
                if 0 <= rr < H and 0 <= cc < W and d[rr, cc] + w < v:
                    v = d[rr, cc] + w
            d[r, c] = v
    return d


def chamfer5_norm_fixed(dy: int, dx: int) -> int:
    a, b = sorted((abs(dy), abs(dx)))
    return LONG * a + HV * (b - 2 * a) if b >= 2 * a else LONG * (b - a) + DIAG * (2 * a - b)


def tolerance_offsets(tolerance: int):
    return [(dy, dx) for dy in range(-tolerance, tolerance + 1) for dx in range(-tolerance, tolerance + 1)
            if chamfer5_norm_fixed(dy, dx) <= tolerance * 65536]


def within_tolerance(boundary: np.ndarray, tolerance: int, exact_transform: bool = False) -> np.ndarray:
    """(distanceTransform(1 - boundary, DIST_L2, 5) <= tolerance) as a bool map."""
    b = boundary > 0
    if exact_transform:
        d = chamfer5_fixed(b)
        return (d.astype(np.float64) / 65536.0).astype(np.float32) <= np.float32(tolerance)
    H, W = b.shape
    out = np.zeros_like(b)
    for dy, dx in tolerance_offsets(tolerance):
        if abs(dy) >= H or abs(dx) >= W:
            continue
        src = b[max(0, -dy):H - max(0, dy), max(0, -dx):W - max(0, dx)]
        out[max(0, dy):H - max(0, -dy), max(0, dx):W - max(0, -dx)] |= src
    return out


def boundary_counts(pred_mask: np.ndarray, target: np.ndarray, tolerance: int = 2, exact_transform: bool = False):
    """(|Bp|, |Bt|, #Bp within tolerance of Bt, #Bt within tolerance of Bp) for one image; tolerance 0: [2] = [3] = |Bp & Bt|."""
    bp, bt = extract_boundaries(pred_mask), extract_boundaries(target)
    if tolerance > 0:
        near_t, near_p = within_tolerance(bt, tolerance, exact_transform), within_tolerance(bp, tolerance, exact_transform)
        a, b = int(((bp > 0) & near_t).sum()), int(((bt > 0) & near_p).sum())
    else:
        a = b = int(((bp > 0) & (bt > 0)).sum())
    return int(bp.sum()), int(bt.sum()), a, b


def f1_from_counts(counts, tolerance: int = 2, smooth: float = 1e-6) -> np.float32:
    """src/evaluate.py:171-191 as NumPy 2 evaluates it: float32 sums, the Python-float smooth is a weak scalar."""
    n_p, n_t, a, b = (np.float32(v) for v in counts)
    s = np.float32(smooth)
    if tolerance > 0:
        precision = (a + s) / (n_p + s)
        recall = (b + s) / (n_t + s)
        return np.float32((np.float32(2.0) * precision * recall + s) / (precision + recall + s))
    return np.float32((np.float32(2.0) * a + s) / (n_p + n_t + s))


def boundary_f1_batch(predictions: np.ndarray, targets: np.ndarray, threshold: float = 0.5, tolerance: int = 2,
                      smooth: float = 1e-6) -> np.ndarray:
    """src/evaluate.py:196-229: (B,1,H,W) probabilities and masks -> float32[B]."""
    p = np.asarray(predictions, dtype=np.float32)
    t = np.asarray(targets, dtype=np.float32)
    B = p.shape[0]
    p = p.reshape(B, p.shape[-2], p.shape[-1])
    t = t.reshape(B, t.shape[-2], t.shape[-1])
    return np.array([f1_from_counts(boundary_counts((p[i] > np.float32(threshold)).astype(np.float32), t[i], tolerance), tolerance, smooth)
                     for i in range(B)], dtype=np.float32)
