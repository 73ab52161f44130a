"""Generate tests/golden/ref_ext.npz by RUNNING THE REAL REFERENCE (build container only): golden vectors for
the two widenings of the path --

  * the S2/S3 sensitivity grids (run_ablation.py:159-224) as loss evaluations: the reference's
    DiceBCEPDELoss constructed once per setting and evaluated on the same probability maps;
  * the per-step accuracy metrics: src/metrics.py compute_dice_score(_batch) and src/evaluate.py
    compute_iou(_batch) at thresholds 0.5 and 0.3.

    python tests/golden/make_golden_ext.py
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

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_ext.npz")
S2_D = [0.5, 1.0, 2.0, 5.0, 10.0, 100.0]        # run_ablation.py:176-188 (pde_weight 1e-3, no phase field)
S3_EPS = [0.001, 0.01, 0.05, 0.1, 0.2]          # run_ablation.py:210-224 (both weights 1e-4, D 5, a 0.5)


def blob(g, B, H, W):
    lo = torch.randn(B, 1, max(H // 8, 2), max(W // 8, 2), generator=g)
    sm = torch.nn.functional.interpolate(lo, size=(H, W), mode="bicubic", align_corners=False)
    t = (sm > 0.3).float()
    z = 4.0 * sm - 1.2 + 0.3 * torch.randn(B, 1, H, W, generator=g)
    return z, t


def main():
    assert ref_loader.available(), "needs the reference checkout (build container)"
    ref_loss, ref_metrics, ref_eval = ref_loader.loss(), ref_loader.metrics(), ref_loader.evaluate()
    src = open(os.path.join(ref_loader.REF_ROOT, "run_ablation.py")).read()
    assert "enumerate([0.5, 1.0, 2.0, 5.0, 10.0, 100.0])" in src and "enumerate([0.001, 0.01, 0.05, 0.1, 0.2])" in src
    torch.set_num_threads(1)
    out, meta = {}, {"torch": torch.__version__}
    g = torch.Generator().manual_seed(4242)
    z, t = blob(g, 3, 48, 64)
    u = torch.sigmoid(z)                 # fp32 probabilities, what train.py hands to the criterion and the metrics
    out["z"], out["t"], out["u"] = z.numpy(), t.numpy(), u.numpy()

    grid = [dict(pde_weight=1e-3, phase_field_weight=0.0, diffusion_coeff=d, reaction_threshold=0.5, epsilon=0.05) for d in S2_D]
    grid += [dict(pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0, reaction_threshold=0.5, epsilon=e) for e in S3_EPS]
    rows32, rows64 = [], []
    for kw in grid:
        for dt, rows in ((torch.float32, rows32), (torch.float64, rows64)):
            crit = ref_loss.DiceBCEPDELoss(dice_weight=0.5, bce_weight=0.5, smooth=1e-6, **kw)
            if dt == torch.float64:
                crit = crit.double()
            with torch.no_grad():
                uu, tt = u.to(dt), t.to(dt)
                total = crit(uu, tt)
                uf, tf = uu.view(-1), tt.view(-1)
                dice = 1 - (2.0 * (uf * tf).sum() + crit.smooth) / (uf.sum() + tf.sum() + crit.smooth)
                rd = crit.pde_regularization.compute_loss(uu)
                pf = crit.pde_regularization.compute_phase_field_loss(uu, epsilon=crit.epsilon)
                rows.append([float(total), float(dice), float(crit.bce(uu, tt)), float(rd), float(pf)])
    out["sweep_f32"], out["sweep_f64"] = np.asarray(rows32), np.asarray(rows64)
    meta["grid"] = grid

    for thr in (0.5, 0.3):
        tag = f"thr{int(thr * 10)}"
        out[f"dice_batch_{tag}"] = ref_metrics.compute_dice_score_batch(u, t, threshold=thr).numpy()
        out[f"iou_batch_{tag}"] = ref_eval.compute_iou_batch(u, t, threshold=thr).numpy()
        out[f"dice_{tag}"] = ref_metrics.compute_dice_score(u, t, threshold=thr).numpy()
        out[f"iou_{tag}"] = ref_eval.compute_iou(u, t, threshold=thr).numpy()
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if k != "meta"})


if __name__ == "__main__":
    main()
