"""GPU tests of the reference-facing interface: the nn.Module drop-ins, their attribute surface as
src/train.py / src/ablation.py use it, autograd behaviour, the stand-alone PDE operators, and the
host-buffer session.  Checker: oracle/ (C oracle in fp64, torch port run on the GPU)."""
import numpy as np
import pytest
import torch

from tests.helpers import blob_inputs, iid_inputs, rel_l2, rel_max, rel_scalar

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def P():
    import physics_informed_image_segmentation_b200 as pkg

    return pkg


@pytest.fixture(scope="module")
def po():
    from oracle import pil_oracle

    return pil_oracle


STAGE2_KW = dict(dice_weight=0.5, bce_weight=0.5, pde_weight=1e-4, phase_field_weight=1e-4,
                 diffusion_coeff=5.0, reaction_threshold=0.5, epsilon=0.05)


def test_drop_in_forward_backward_like_train_step(P, po, dev):
    """criterion(outputs, masks) -> loss.backward() exactly as src/train.py:113-117,:163 with a tiny model"""
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(1, 1, 3, padding=1), torch.nn.Sigmoid()).to(dev)
    images, masks = torch.rand(4, 1, 64, 64, device=dev), (torch.rand(4, 1, 64, 64, device=dev) > 0.5).float()
    criterion = P.DiceBCEPDELoss(**STAGE2_KW).to(dev)
    outputs = model(images)
    loss = criterion(outputs, masks)
    assert loss.dim() == 0 and loss.dtype == torch.float32 and loss.device == outputs.device and loss.requires_grad
    outputs.retain_grad()
    loss.backward()
    assert all(p_.grad is not None and torch.isfinite(p_.grad).all() for p_ in model.parameters())
    u64, t64 = outputs.detach().cpu().numpy().astype(np.float64), masks.cpu().numpy().astype(np.float64)
    comps, og = po.loss_and_grad(u64, t64, po.STAGE2, po.X_PROB)
    assert rel_scalar(loss.item(), comps[0]) < TOL
    assert rel_max(outputs.grad.cpu().numpy(), og) < TOL
    # same step through the reference-equivalent torch ops on the GPU (the eager path being replaced)
    from oracle import torch_port

    model.zero_grad()
    out2 = model(images)
    out2.retain_grad()
    L2 = torch_port.loss(out2, masks, po.STAGE2)
    L2.backward()
    assert rel_scalar(loss.item(), L2.item()) < TOL
    assert rel_max(outputs.grad.cpu().numpy(), out2.grad.cpu().numpy()) < TOL


def test_logging_block_surface_and_cache(P, po, dev):
    """src/train.py:120-150 re-evaluates every term under no_grad through these attributes."""
    from physics_informed_image_segmentation_b200 import _lib, functional

    z, t = blob_inputs(3, 48, 56, seed=3)
    outputs = torch.sigmoid(z).to(dev).requires_grad_(True)
    masks = t.to(dev)
    criterion = P.DiceBCEPDELoss(**STAGE2_KW).to(dev)
    loss = criterion(outputs, masks)
    n0 = functional.launch_info().kernels_launched
    with torch.no_grad():
        smooth = criterion.smooth
        pf, tf = outputs.view(-1), masks.view(-1)
        dice_loss = 1 - (2.0 * (pf * tf).sum() + smooth) / (pf.sum() + tf.sum() + smooth)
        bce_loss = criterion.bce(outputs, masks)
        assert isinstance(criterion, P.DiceBCEPDELoss) and criterion.pde_weight > 0 and criterion.phase_field_weight > 0
        pde_loss = criterion.pde_regularization.compute_loss(outputs)
        pf_loss = criterion.pde_regularization.compute_phase_field_loss(outputs, epsilon=criterion.epsilon)
    # served from the fused forward's report: no kernel of ours ran for the three component calls
    assert functional.launch_info().kernels_launched == n0
    u64, t64 = outputs.detach().cpu().numpy().astype(np.float64), t.numpy().astype(np.float64)
    comps, _ = po.loss_and_grad(u64, t64, po.STAGE2, po.X_PROB)
    for got, want in ((dice_loss, comps[1]), (bce_loss, comps[2]), (pde_loss, comps[3]), (pf_loss, comps[4])):
        assert rel_scalar(got.item(), want) < TOL
    rep = criterion.components()
    assert rel_scalar(rep["dice_loss"].item(), comps[1]) < TOL and rel_scalar(rep["loss"].item(), comps[0]) < TOL
    # a different tensor, or a different epsilon, must NOT hit the cache
    other = outputs.detach().clone()
    with torch.no_grad():
        v = criterion.pde_regularization.compute_loss(other)
        w = criterion.pde_regularization.compute_phase_field_loss(outputs, epsilon=0.1)
    assert functional.launch_info().kernels_launched == n0 + 2
    assert rel_scalar(v.item(), comps[3]) < TOL
    import dataclasses
    comps_eps, _ = po.loss_and_grad(u64, t64, dataclasses.replace(po.STAGE2, epsilon=0.1), po.X_PROB)
    assert rel_scalar(w.item(), comps_eps[4]) < TOL
    loss.backward()
    # in-place modification invalidates the cache
    with torch.no_grad():
        outputs.mul_(0.5)
        v2 = criterion.pde_regularization.compute_loss(outputs)
    assert functional.launch_info().kernels_launched >= n0 + 4


def test_component_methods_are_differentiable(P, po, dev):
    """PDERegularization.compute_loss / compute_phase_field_loss / criterion.bce used on their own"""
    import dataclasses

    z, t = iid_inputs(2, 40, 44, seed=12)
    u0 = torch.sigmoid(z)
    reg = P.PDERegularization(5.0, 0.5).to(dev)
    t64 = t.numpy().astype(np.float64)
    cases = [
        (lambda u: reg.compute_loss(u), dataclasses.replace(po.STAGE2, dice_weight=0, bce_weight=0, pde_weight=1.0, phase_field_weight=0.0), 3),
        (lambda u: reg.compute_phase_field_loss(u, epsilon=0.07), dataclasses.replace(po.STAGE2, dice_weight=0, bce_weight=0, pde_weight=0.0, phase_field_weight=1.0, epsilon=0.07), 4),
        (lambda u: P.DiceBCELoss().to(dev).bce(u, t.to(dev)), dataclasses.replace(po.STAGE2, dice_weight=0, bce_weight=1.0, pde_weight=0.0, phase_field_weight=0.0), 2),
    ]
    for fn, p, idx in cases:
        u = u0.to(dev).requires_grad_(True)
        val = fn(u)
        (3.0 * val).backward()  # upstream gradient != 1 is read on the device
        comps, og = po.loss_and_grad(u0.numpy().astype(np.float64), t64, p, po.X_PROB)
        assert rel_scalar(val.item(), comps[idx]) < TOL
        assert rel_max(u.grad.cpu().numpy(), 3.0 * og) < TOL


def test_forward_logits_and_dice_bce_loss(P, po, dev):
    import dataclasses

    z, t = blob_inputs(2, 64, 96, seed=9)
    crit = P.DiceBCEPDELoss(**STAGE2_KW).to(dev)
    zl = z.to(dev).requires_grad_(True)
    loss = crit.forward_logits(zl, t.to(dev), activation="sigmoid")
    loss.backward()
    comps, og = po.loss_and_grad(z.numpy().astype(np.float64), t.numpy().astype(np.float64), po.STAGE2, po.X_LOGITS_SIGMOID)
    assert rel_scalar(loss.item(), comps[0]) < TOL and rel_max(zl.grad.cpu().numpy(), og) < TOL
    with pytest.raises(ValueError):
        crit.forward_logits(zl, t.to(dev), activation="relu")
    base = P.DiceBCELoss(dice_weight=0.3, bce_weight=0.7).to(dev)
    u = torch.sigmoid(z).to(dev).requires_grad_(True)
    lb = base(u, t.to(dev))
    lb.backward()
    p = dataclasses.replace(po.STAGE2, dice_weight=0.3, bce_weight=0.7, pde_weight=0.0, phase_field_weight=0.0)
    comps, og = po.loss_and_grad(torch.sigmoid(z).numpy().astype(np.float64), t.numpy().astype(np.float64), p, po.X_PROB)
    assert rel_scalar(lb.item(), comps[0]) < TOL and rel_max(u.grad.cpu().numpy(), og) < TOL


def test_no_grad_path_and_weights_read_at_call_time(P, po, dev):
    z, t = iid_inputs(2, 32, 32, seed=1)
    u, m = torch.sigmoid(z).to(dev), t.to(dev)
    crit = P.DiceBCEPDELoss(**STAGE2_KW).to(dev)
    with torch.no_grad():
        a = crit(u, m)
    assert not a.requires_grad
    crit.pde_weight = 0.0  # plain attribute like the reference; gate evaluated per call (src/loss.py:150)
    crit.phase_field_weight = 0.0
    crit.epsilon = -1.0    # not validated while the phase-field term is off (src/loss.py:155)
    b = crit(u, m)
    r = crit.components()
    assert rel_scalar(b.item(), 0.5 * r["dice_loss"].item() + 0.5 * r["bce_loss"].item()) < 1e-6
    crit.phase_field_weight = 1e-4
    with pytest.raises(ValueError, match="epsilon must be positive"):
        crit(u, m)


def test_input_contract_errors(P, dev):
    crit = P.DiceBCEPDELoss().to(dev)
    u = torch.rand(2, 1, 16, 16, device=dev)
    m = torch.zeros_like(u)
    with pytest.raises(RuntimeError, match="contiguous"):
        crit(u.transpose(2, 3), m)
    with pytest.raises(ValueError, match="target size"):
        crit(u, m[:1])
    with pytest.raises(RuntimeError, match="single-channel"):
        crit(torch.rand(2, 3, 16, 16, device=dev), torch.rand(2, 3, 16, 16, device=dev))
    with pytest.raises(RuntimeError, match="reflect"):
        crit(torch.rand(2, 1, 1, 16, device=dev), torch.rand(2, 1, 1, 16, device=dev))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit(u.cpu(), m.cpu())
    # out-of-range probabilities are counted (nn.BCELoss raises on them): report slot 5
    bad = u.clone()
    bad[0, 0, 0, 0] = 1.5
    bad[1, 0, 3, 3] = float("nan")
    with torch.no_grad():
        loss_bad = crit(bad, m)
    assert crit.last_report[5].item() == 2.0
    assert torch.isnan(loss_bad)          # no host sync, but not silent either
    with torch.no_grad():
        assert torch.isfinite(crit(u, m))
    crit.strict_inputs = True             # opt-in: the reference's exception (one sync per call)
    with pytest.raises(RuntimeError, match="between 0 and 1"):
        crit(bad, m)
    crit(u, m)


def test_pde_operators_and_their_autograd(P, po, dev):
    """compute_laplacian / reaction_term / compute_residual / compute_gradient_magnitude incl. the
    diffusion-only loss of src/ablation.py:77-86 built from them"""
    from oracle import torch_port

    z, _ = blob_inputs(2, 37, 53, seed=14)
    u_cpu = torch.sigmoid(z)
    reg = P.PDERegularization(diffusion_coeff=2.0, reaction_threshold=0.4).to(dev)
    u = u_cpu.to(dev).requires_grad_(True)
    lap = reg.compute_laplacian(u)
    assert rel_max(lap.detach().cpu().numpy(), po.laplacian(u_cpu.numpy())) < 1e-6
    assert rel_max(reg.reaction_term(u).detach().cpu().numpy(), po.reaction(u_cpu.numpy(), 0.4)) < 1e-6
    assert rel_max(reg.compute_gradient_magnitude(u).detach().cpu().numpy(), po.grad_mag_sq(u_cpu.numpy())) < 1e-6
    # gradients of scalar functionals of each operator vs the torch port's autograd (fp64 on CPU)
    w = torch.randn(u_cpu.shape, generator=torch.Generator().manual_seed(5))
    for mine, theirs in (
        (lambda v: (reg.compute_laplacian(v) * w.to(v.device)).sum(), lambda v: (torch_port.laplacian(v) * w.double()).sum()),
        (lambda v: (reg.compute_gradient_magnitude(v) * w.to(v.device)).sum(), lambda v: (torch_port.grad_mag_sq(v) * w.double()).sum()),
        (lambda v: torch.mean(reg.compute_residual(v) ** 2), lambda v: torch_port.rd_loss(v, 2.0, 0.4)),
        (lambda v: torch.mean((reg.diffusion_coeff * reg.compute_laplacian(v)) ** 2),
         lambda v: torch.mean((2.0 * torch_port.laplacian(v)) ** 2)),  # diffusion-only ablation loss
    ):
        ug = u_cpu.to(dev).requires_grad_(True)
        mine(ug).backward()
        ud = u_cpu.double().requires_grad_(True)
        theirs(ud).backward()
        assert rel_max(ug.grad.cpu().numpy(), ud.grad.numpy()) < TOL
        assert rel_l2(ug.grad.cpu().numpy(), ud.grad.numpy()) < TOL
    assert P.create_pde_regularization(3.0, 0.25).diffusion_coeff == 3.0


def test_host_session_end_to_end(P, po, dev):
    """pil_session_*: host buffers in, loss report + gradient in host memory out"""
    z, t = blob_inputs(20, 64, 128, seed=31)  # 20 images -> 16 uneven chunks
    params = P.LossParams(**STAGE2_KW)
    grad = torch.empty_like(z).pin_memory()
    with P.HostSession(32, 64, 128) as sess:
        rep = sess.run(z.pin_memory(), t.pin_memory(), params, grad_host=grad, activation="sigmoid")
        rep2 = sess.run(z.numpy(), t.numpy(), params, grad_host=None, activation="sigmoid")  # pageable, loss only
    comps, og = po.loss_and_grad(z.numpy().astype(np.float64), t.numpy().astype(np.float64), po.STAGE2, po.X_LOGITS_SIGMOID)
    for k in range(5):
        assert rel_scalar(rep[k], comps[k]) < TOL
        assert rep[k] == rep2[k]
    assert rel_max(grad.numpy(), og) < TOL and rel_l2(grad.numpy(), og) < TOL


def test_emulated_two_rank_data_parallel(P, po, dev):
    """Two 'ranks' as two batch slices on one GPU: K1 per slice, sums added (what the all-reduce does),
    finalize with the device-side pixel count, K2 per slice with grad_scale = world_size."""
    from physics_informed_image_segmentation_b200 import functional as Fn

    z, t = iid_inputs(5, 48, 64, seed=17)
    x, m, p = z.to(dev), t.to(dev), P.LossParams(**STAGE2_KW)
    bounds = [P.shard_bounds(5, r, 2) for r in range(2)]
    assert bounds == [(0, 3), (3, 5)]
    parts = [Fn.forward_sums(x[a:b], m[a:b], p, 1, finalize=False)[0] for a, b in bounds]
    gs = parts[0] + parts[1]
    rep = Fn.finalize_report(gs, -1, p)
    grads = torch.cat([Fn.backward_grad(x[a:b], m[a:b], p, 1, gs, -1, grad_scale=2.0) for a, b in bounds])
    comps, og = po.loss_and_grad(z.numpy().astype(np.float64), t.numpy().astype(np.float64), po.STAGE2, 1)
    assert rel_scalar(rep[0].item(), comps[0]) < TOL
    assert rel_max(grads.cpu().numpy(), 2.0 * og) < TOL
    host = P.loss_report_from_sums(gs.cpu().tolist(), None, p)
    assert rel_scalar(host["loss"], comps[0]) < TOL


def test_host_session_gradient_on_device(dev):
    """pil_session_run_ex(PIL_SESSION_GRAD_ON_DEVICE): same loss report and the same gradient as the
    host-gradient path, with only the report crossing back."""
    import physics_informed_image_segmentation_b200 as P
    from tests.helpers import iid_inputs

    z, t = iid_inputs(5, 48, 64, seed=9)
    p = P.LossParams(pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0)
    zh, th = z.pin_memory(), t.pin_memory()
    gh = torch.empty_like(zh).pin_memory()
    with P.HostSession(5, 48, 64, device=0) as sess:
        rep_h = sess.run(zh, th, p, grad_host=gh)
        rep_d = sess.run(zh, th, p, grad_on_device=True)
        gd = sess.device_gradient().clone()
    torch.cuda.synchronize()
    assert np.allclose(rep_d[:5], rep_h[:5], rtol=2e-6, atol=0)
    assert gd.shape == (5, 1, 48, 64)
    den = gh.abs().max().item()
    assert (gd.cpu() - gh).abs().max().item() / den < 2e-6


def _ref_dice_iou(u, t, thr=0.5, smooth=1e-6):
    """reference src/metrics.py:38-73 and src/evaluate.py:62-97, restated with torch on the CPU"""
    pb = (u > thr).float()
    B = u.shape[0]
    I = (pb * t).reshape(B, -1).double().sum(1)
    P = pb.reshape(B, -1).double().sum(1)
    T = t.reshape(B, -1).double().sum(1)
    return ((2 * I + smooth) / (P + T + smooth)).float(), ((I + smooth) / (P + T - I + smooth)).float()


@pytest.mark.parametrize("shape", [(6, 64, 96), (3, 33, 50), (2, 7, 9), (5, 256, 256)])
@pytest.mark.parametrize("entry", ["prob", "logits"])
def test_batch_metrics_ride_on_the_training_step(dev, shape, entry):
    """Per-image thresholded Dice / IoU (reference train_epoch, src/train.py:153-160) from the same pass as the
    loss: values match the reference formulas, and loss / gradient are unchanged by switching them on."""
    import physics_informed_image_segmentation_b200 as P
    from tests.helpers import blob_inputs

    B, H, W = shape
    z, t = blob_inputs(B, H, W, seed=31)
    u = torch.sigmoid(z)
    crit = P.DiceBCEPDELoss(pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0).to(dev)

    def run(metrics):
        crit.enable_batch_metrics(0.5 if metrics else None)
        x = (u if entry == "prob" else z).to(dev).requires_grad_(True)
        loss = crit(x, t.to(dev)) if entry == "prob" else crit.forward_logits(x, t.to(dev))
        loss.backward()
        return loss.detach(), x.grad

    l0, g0 = run(False)
    l1, g1 = run(True)
    m = crit.last_batch_metrics()
    assert abs(l0.item() - l1.item()) <= 2e-6 * abs(l0.item())
    assert (g0 - g1).abs().max().item() <= 2e-6 * g0.abs().max().item()
    # the threshold is applied to the kernel's own fp32 u: compare against u as the GPU computes it for logits
    u_ref = u if entry == "prob" else torch.sigmoid(z.to(dev)).cpu()
    dice_ref, iou_ref = _ref_dice_iou(u_ref, t)
    assert m["dice"].shape == (B,) and m["iou"].shape == (B,)
    assert torch.allclose(m["dice"].cpu(), dice_ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(m["iou"].cpu(), iou_ref, rtol=1e-5, atol=1e-6)
    # no-grad (validation) path
    with torch.no_grad():
        crit(u.to(dev), t.to(dev))
    mv = crit.last_batch_metrics()
    d2, i2 = _ref_dice_iou(u, t)
    assert torch.allclose(mv["dice"].cpu(), d2, rtol=1e-5, atol=1e-6) and torch.allclose(mv["iou"].cpu(), i2, rtol=1e-5, atol=1e-6)
    crit.enable_batch_metrics(None)


def test_metric_drop_ins_and_count_cache(dev):
    """compute_dice_score(_batch) / compute_iou(_batch) with the reference's signatures: values match the
    reference formulas; after a loss evaluation with batch metrics on the SAME tensors they launch no pass over
    the maps (only the tiny per-image finalisation)."""
    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200 import functional as Fn
    from tests.helpers import blob_inputs

    B, H, W = 5, 80, 112
    z, t = blob_inputs(B, H, W, seed=41)
    u = torch.sigmoid(z)
    ud, td = u.to(dev), t.to(dev)
    dice_ref, iou_ref = _ref_dice_iou(u, t)
    assert torch.allclose(P.compute_dice_score_batch(ud, td).cpu(), dice_ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(P.compute_iou_batch(ud, td, threshold=0.5).cpu(), iou_ref, rtol=1e-5, atol=1e-6)
    pb = (u > 0.5).float()
    I, Pn, T = (pb * t).sum().double(), pb.sum().double(), t.sum().double()
    assert abs(P.compute_dice_score(ud, td).item() - float((2 * I + 1e-6) / (Pn + T + 1e-6))) < 1e-6
    assert abs(P.compute_iou(ud, td).item() - float((I + 1e-6) / (Pn + T - I + 1e-6))) < 1e-6
    d3, _ = _ref_dice_iou(u, t, thr=0.3)
    assert torch.allclose(P.compute_dice_score_batch(ud, td, threshold=0.3).cpu(), d3, rtol=1e-5, atol=1e-6)

    crit = P.DiceBCEPDELoss(pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0).to(dev).enable_batch_metrics(0.5)
    x = ud.clone().requires_grad_(True)
    loss = crit(x, td)
    k0 = Fn.launch_info().kernels_launched
    d = P.compute_dice_score_batch(x, td, threshold=0.5)       # what src/train.py:154-155 calls right after the loss
    i = P.compute_iou_batch(x, td, threshold=0.5)
    assert Fn.launch_info().kernels_launched - k0 == 2          # two B-element finalisations, no pass over the maps
    assert torch.allclose(d.cpu(), dice_ref, rtol=1e-5, atol=1e-6) and torch.allclose(i.cpu(), iou_ref, rtol=1e-5, atol=1e-6)
    loss.backward()
    # a different threshold, or a modified tensor, is not served from the cache
    k0 = Fn.launch_info().kernels_launched
    P.compute_dice_score_batch(x, td, threshold=0.4)
    assert Fn.launch_info().kernels_launched - k0 == 2          # counts kernel + finalisation
    with pytest.raises(RuntimeError):
        P.compute_dice_score_batch(u, t)                          # CPU tensors: no fallback
