"""Generate tests/golden/ref_bf1.npz by RUNNING THE REAL REFERENCE (build container only; needs cv2): golden vectors of
the per-step boundary-F1 metric -- src/evaluate.py extract_boundaries, compute_boundary_f1, compute_boundary_f1_batch.

    python tests/golden/make_golden_bf1.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_bf1.npz")


def structured_masks(H, W):
    """Masks that exercise what RETR_EXTERNAL and the chamfer tolerance distinguish."""
    m = []
    a = np.zeros((H, W), np.float32); a[4:20, 5:30] = 1; a[8:16, 10:24] = 0; a[10:14, 14:20] = 1   # ring with an island in its hole
    m.append(a)
    b = np.zeros((H, W), np.float32); b[0:10, 0:12] = 1; b[H - 6:, W - 9:] = 1; b[0, W - 1] = 1     # objects touching the frame
    m.append(b)
    c = np.zeros((H, W), np.float32)
    for k in range(min(H, W) - 4):
        c[2 + k, 2 + k] = 1                                                                           # 8-connected diagonal line
    m.append(c)
    d = ((np.indices((H, W)).sum(0) % 2) == 0).astype(np.float32)                                     # checkerboard
    m.append(d)
    m.append(np.zeros((H, W), np.float32))                                                            # empty
    m.append(np.ones((H, W), np.float32))                                                             # full
    e = np.zeros((H, W), np.float32); e[H // 2, 3:W - 3] = 1; e[5:H - 5, W // 2] = 1; e[7, 7] = 1      # thin cross + single pixel
    m.append(e)
    f = np.ones((H, W), np.float32); f[3:H - 3, 3:W - 3] = 0; f[6:H - 6, 6:W - 6] = 1; f[H // 2, 0:3] = 0  # frame with a gap + filled core
    m.append(f)
    return np.stack(m)


def main():
    assert ref_loader.available(), "needs the reference checkout (build container)"
    ev = ref_loader.evaluate()
    import cv2

    rng = np.random.default_rng(77)
    g = torch.Generator().manual_seed(77)
    out, meta = {}, {"torch": torch.__version__, "numpy": np.__version__, "cv2": cv2.__version__}
    H, W = 40, 56
    st = structured_masks(H, W)
    # predictions: the structured masks shifted / perturbed, as probabilities around the threshold
    shifted = np.roll(st, shift=(1, 2), axis=(1, 2))
    pred_struct = np.clip(0.5 + (shifted - 0.5) * 0.8 + 0.05 * rng.standard_normal(st.shape), 0, 1).astype(np.float32)
    # smooth random blobs and iid noise
    lo = torch.randn(6, 1, 5, 7, generator=g)
    sm = torch.nn.functional.interpolate(lo, size=(H, W), mode="bicubic", align_corners=False)
    t_blob = (sm > 0.2).float().numpy()[:, 0]
    p_blob = torch.sigmoid(3.0 * sm - 0.5 + 0.8 * torch.randn(6, 1, H, W, generator=g)).numpy()[:, 0]
    t_iid = (rng.random((3, H, W)) > 0.6).astype(np.float32)
    p_iid = rng.random((3, H, W)).astype(np.float32)
    targets = np.concatenate([st, t_blob, t_iid])[:, None]
    preds = np.concatenate([pred_struct, p_blob, p_iid])[:, None].astype(np.float32)
    out["predictions"], out["targets"] = preds, targets.astype(np.float32)
    B = preds.shape[0]
    out["boundary_pred"] = np.stack([ev.extract_boundaries((preds[i, 0] > 0.5).astype(np.float32)) for i in range(B)])
    out["boundary_target"] = np.stack([ev.extract_boundaries(targets[i, 0]) for i in range(B)])
    pt, tt = torch.from_numpy(preds), torch.from_numpy(targets.astype(np.float32))
    for tol in (0, 1, 2, 3):
        out[f"f1_tol{tol}"] = ev.compute_boundary_f1_batch(pt, tt, threshold=0.5, tolerance=tol).numpy()
    out["f1_thr03"] = ev.compute_boundary_f1_batch(pt, tt, threshold=0.3, tolerance=2).numpy()
    out["f1_first"] = np.array([ev.compute_boundary_f1(pt, tt).item()], dtype=np.float32)
    # the tolerance masks themselves (distanceTransform <= tol), for the chamfer restatement
    for tol in (1, 2, 3, 4, 5, 6):
        out[f"near_target_tol{tol}"] = np.stack([
            cv2.distanceTransform((1 - out["boundary_target"][i]).astype(np.uint8), cv2.DIST_L2, 5) <= tol for i in range(B)])
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {B} images of {H}x{W}; f1(tol 2) = {np.round(out['f1_tol2'], 4).tolist()}")


if __name__ == "__main__":
    main()
