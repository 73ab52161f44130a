"""Parameter sweep (BASELINE config 4): K loss evaluations from one pass over the maps, against the CPU
oracle evaluated K times.  Tolerance 1e-5 relative on the total loss and on every component (fp32)."""
import numpy as np
import pytest
import torch

from tests.helpers import blob_inputs, iid_inputs, rel_scalar

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _oracle_rows(po, z64, t64, grid, kind):
    rows = []
    for p in grid:
        pp = po.Params(dice_weight=p.dice_weight, bce_weight=p.bce_weight, pde_weight=p.pde_weight,
                       phase_field_weight=p.phase_field_weight, diffusion_coeff=p.diffusion_coeff,
                       reaction_threshold=p.reaction_threshold, epsilon=p.epsilon, smooth=p.smooth)
        s = po.sums(z64, t64, pp, kind)
        rows.append(po.finalize(s, int(s[7]), pp))
    return np.asarray(rows)


@pytest.mark.parametrize("maker,B,H,W", [(blob_inputs, 4, 512, 512), (iid_inputs, 2, 128, 256), (blob_inputs, 3, 67, 45)])
@pytest.mark.parametrize("activation", ["sigmoid", "none"])
def test_s2_s3_grids_match_oracle(dev, maker, B, H, W, activation):
    from oracle import pil_oracle as po
    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200 import functional as Fn

    z, t = maker(B, H, W, seed=11)
    if activation == "none":
        z = torch.sigmoid(z)
        kind = po.X_PROB
    else:
        kind = po.X_LOGITS_SIGMOID
    grid = P.s2_grid() + P.s3_grid() + [Fn.LossParams(pde_weight=1e-2, phase_field_weight=1e-3, diffusion_coeff=3.0,
                                                     reaction_threshold=0.3, epsilon=0.07)]
    rep = P.sweep_losses(z.to(dev), t.to(dev), grid, activation=activation).cpu().numpy().astype(np.float64)
    ref = _oracle_rows(po, z.numpy().astype(np.float64), t.numpy().astype(np.float64), grid, kind)
    assert rep.shape == (len(grid), 8)
    for k, p in enumerate(grid):
        for c in range(5):
            if c == 4 and not p.phase_field_weight > 0:
                continue  # the reference never evaluates the phase-field term when its weight is 0
            assert rel_scalar(rep[k, c], ref[k][c]) < 1e-5, (k, c, rep[k, c], ref[k][c], p)


def test_sweep_row_equals_fused_forward(dev):
    """Each sweep row must agree with the ordinary fused forward of that setting (same kernel family)."""
    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200 import functional as Fn

    z, t = blob_inputs(2, 96, 160, seed=3)
    x, tt = z.to(dev), t.to(dev)
    grid = P.s2_grid() + P.s3_grid()
    rep = P.sweep_losses(x, tt, grid, activation="sigmoid")
    for k, p in enumerate(grid):
        _, one = Fn.forward_sums(x, tt, p, Fn.X_LOGITS_SIGMOID)
        for c in range(4):
            assert rel_scalar(rep[k, c].item(), one[c].item()) < 2e-6, (k, c)


def test_sweep_validates_every_setting(dev):
    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200 import functional as Fn

    z, t = iid_inputs(1, 8, 8)
    with pytest.raises(ValueError, match="diffusion_coeff must be positive"):
        P.sweep_losses(z.to(dev), t.to(dev), [Fn.LossParams(), Fn.LossParams(diffusion_coeff=0.0)], activation="sigmoid")


def test_sweep_and_metrics_against_reference_golden(dev):
    """CUDA path against vectors produced by the REAL reference (tests/golden/make_golden_ext.py): the S2+S3
    grids as one batched evaluation, and the per-image / global thresholded Dice and IoU."""
    import json
    import os

    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200 import functional as Fn

    data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ext.npz"))
    meta = json.loads(bytes(data["meta"]).decode())
    u, t = torch.from_numpy(data["u"]).to(dev), torch.from_numpy(data["t"]).to(dev)
    grid = [Fn.LossParams(dice_weight=0.5, bce_weight=0.5, smooth=1e-6, **kw) for kw in meta["grid"]]
    assert [g.diffusion_coeff for g in P.s2_grid()] == [kw["diffusion_coeff"] for kw in meta["grid"][:6]]
    assert [g.epsilon for g in P.s3_grid()] == [kw["epsilon"] for kw in meta["grid"][6:]]
    rep = P.sweep_losses(u, t, grid, activation="none").cpu().numpy().astype(np.float64)
    ref = data["sweep_f64"]
    for k, kw in enumerate(meta["grid"]):
        for c in range(5):
            if c == 4 and not kw["phase_field_weight"] > 0:
                continue
            assert rel_scalar(rep[k, c], ref[k, c]) < 1e-5, (k, c, rep[k, c], ref[k, c])
    for thr, tag in ((0.5, "thr5"), (0.3, "thr3")):
        assert np.allclose(P.compute_dice_score_batch(u, t, threshold=thr).cpu().numpy(), data[f"dice_batch_{tag}"], rtol=1e-5, atol=1e-7)
        assert np.allclose(P.compute_iou_batch(u, t, threshold=thr).cpu().numpy(), data[f"iou_batch_{tag}"], rtol=1e-5, atol=1e-7)
        assert abs(P.compute_dice_score(u, t, threshold=thr).item() - float(data[f"dice_{tag}"])) < 1e-5
        assert abs(P.compute_iou(u, t, threshold=thr).item() - float(data[f"iou_{tag}"])) < 1e-5
