"""The training / validation step around the fused loss: drop-ins for the reference's `train_epoch` and `validate`
(reference src/train.py:84-185, :188-286) with the same signatures and the same result dictionaries.

What the reference's step does around `loss = criterion(outputs, masks)` (src/train.py:108-167):
  * the model applies its output activation, the loss reads probabilities           -> here: the model runs with a logits
    head (integration.use_logits_head) and the activation is fused into both loss kernels (forward_logits);
  * a second, no-grad evaluation of Dice, BCE, RD and phase-field terms for logging  -> here: the four components come
    (:120-150) with one `.item()` host sync each                                        out of the SAME kernels as the
                                                                                        loss (criterion.last_report);
  * per-image thresholded Dice / IoU in Python loops (:154-155), boundary-F1 through -> here: threshold counts ride on
    B `.cpu().numpy()` copies and OpenCV on the host (:156), `.cpu()` of all three      the loss's forward pass; the
    (:158-160)                                                                          boundary-F1 is counted by CUDA
                                                                                        kernels; nothing leaves the GPU;
  * `loss.item()` every step (:166)                                                  -> here: running sums stay on the
                                                                                        device; ONE host sync per epoch.
The returned numbers are the reference's: means over batches of the loss terms, means over samples of the metrics.

Data parallel (not in the reference): pass a criterion built with `process_group=...` and a DDP-wrapped model.  Every
rank then evaluates the loss of the GLOBAL batch (the kernels exchange their partial sums), the loss gradient carries
the factor world_size that DDP's gradient averaging removes again (ddp_average=True), and the per-sample metrics are
combined across ranks at the end of the epoch.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import functional as Fn
from .integration import use_logits_head
from .loss import DiceBCEPDELoss, _FusedLossBase


def _unwrap(model):
    """The module whose `activation_name` decides what forward() returns (DDP / DataParallel wrap it in `.module`)."""
    inner = model
    while hasattr(inner, "module") and not hasattr(inner, "activation_name"):
        inner = inner.module
    return inner


class _EpochAccumulator:
    """Running sums of one epoch, all on the device: report vectors per batch, metric sums per sample."""

    def __init__(self, device: torch.device):
        self.report = torch.zeros(Fn.PIL_NOUT, dtype=torch.float64, device=device)   # sum over batches of the loss report
        self.metrics = torch.zeros(4, dtype=torch.float64, device=device)           # sum dice, sum iou, sum boundary-F1, #samples
        self.batch_dice = torch.zeros(1, dtype=torch.float64, device=device)        # validate(): sum over batches of the batch-level Dice
        self.batches = 0

    def add_report(self, report: torch.Tensor) -> None:
        self.report += report.to(torch.float64)
        self.batches += 1

    def add_metrics(self, dice: torch.Tensor, iou: torch.Tensor, bf1: torch.Tensor) -> None:
        self.metrics[0] += dice.sum(dtype=torch.float64)
        self.metrics[1] += iou.sum(dtype=torch.float64)
        self.metrics[2] += bf1.sum(dtype=torch.float64)
        self.metrics[3] += dice.numel()


def _step_loss(model, criterion: _FusedLossBase, images: torch.Tensor, masks: torch.Tensor):
    """Forward of the model and the fused loss.  Returns (loss, x, kind): x is what the loss kernels read -- logits when
    the model's activation could be moved into the kernels, probabilities otherwise."""
    inner = _unwrap(model)
    name = getattr(inner, "activation_name", None)
    if name in ("sigmoid", "tanh"):
        with use_logits_head(inner):           # reference src/unet.py:208-214: any other name skips the activation
            logits = model(images)
        return criterion.forward_logits(logits, masks, activation=name), logits, Fn.activation_kind(name)
    outputs = model(images)                    # a model without the reference's activation switch: probabilities
    return criterion(outputs, masks), outputs, Fn.X_PROB


def _batch_metrics(criterion: _FusedLossBase, x: torch.Tensor, masks: torch.Tensor, kind: int, acc: _EpochAccumulator,
                   threshold: float = 0.5, tolerance: int = 2) -> None:
    m = criterion.last_batch_metrics()                                   # per-image Dice / IoU from the loss's own pass
    counts = Fn.boundary_counts(x.detach(), masks, kind, threshold, tolerance)
    acc.add_metrics(m["dice"], m["iou"], Fn.boundary_f1(counts, tolerance))


def _results(criterion, acc: _EpochAccumulator, return_components: bool, compute_metrics: bool, validation: bool,
             group) -> Dict[str, float]:
    """One device-to-host copy for the whole epoch, then the reference's result dictionary (src/train.py:169-185,
    :267-286)."""
    vec = torch.cat([acc.report, acc.metrics, acc.batch_dice])
    if group is not None:
        # the epoch's host sync is also where a peer-exchange time-out surfaces: a rank that lost its peers has been
        # writing zero gradients and NaN losses since (include/pil.h); stop instead of training on
        from .sharding import peer_exchange_for

        px = peer_exchange_for(group, vec.device)
        if px is not None:
            px.check_lockstep()
        if px is not None and px.timed_out():
            raise RuntimeError("a peer-memory exchange wait timed out during this epoch: a rank died or the ranks are not "
                               "evaluating the loss in lock step (PIL_XCHG_TIMEOUT_MS)")
    if group is not None and compute_metrics:
        import torch.distributed as dist

        # the loss report is already global on every rank; the per-sample metric sums are local to the rank's shard
        # (the batch-level Dice of validate() becomes the mean over ranks of the ranks' batch-level scores)
        local = vec[Fn.PIL_NOUT:].clone()
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
        local[4] /= dist.get_world_size(group)
        vec[Fn.PIL_NOUT:] = local
    host = vec.cpu().tolist()  # the epoch's only host sync
    rep, met, bdice = host[:Fn.PIL_NOUT], host[Fn.PIL_NOUT:Fn.PIL_NOUT + 4], host[-1]
    nb = max(acc.batches, 1)
    results = {"loss": rep[Fn.OUT_TOTAL] / nb}
    if validation:
        results["dice_score"] = bdice / nb
    if return_components:
        results["dice_loss"] = rep[Fn.OUT_DICE] / nb
        results["bce_loss"] = rep[Fn.OUT_BCE] / nb
        if isinstance(criterion, DiceBCEPDELoss):
            if criterion.pde_weight > 0:
                results["pde_loss"] = rep[Fn.OUT_RD] / nb
            if criterion.phase_field_weight > 0:
                results["phase_field_loss"] = rep[Fn.OUT_PF] / nb
    if compute_metrics:
        ns = met[3]
        if not validation:
            results["dice_score"] = met[0] / ns if ns else 0.0
        results["iou_score"] = met[1] / ns if ns else 0.0
        results["boundary_f1_score"] = met[2] / ns if ns else 0.0
    return results


def _check_criterion(criterion) -> _FusedLossBase:
    if not isinstance(criterion, _FusedLossBase):
        raise TypeError("train_epoch / validate of this package drive the fused loss modules (DiceBCELoss, DiceBCEPDELoss of "
                        "physics_informed_image_segmentation_b200); call install_into_reference() before the reference builds "
                        f"its criterion, got {type(criterion).__name__}")
    return criterion


def train_epoch(model, dataloader, criterion, optimizer, device, return_components: bool = False,
                compute_metrics: bool = True) -> Dict[str, float]:
    """Train for one epoch (reference src/train.py:84-185: same arguments, same keys in the result)."""
    criterion = _check_criterion(criterion)
    model.train()
    acc: Optional[_EpochAccumulator] = None
    saved_thr = criterion.batch_metrics_threshold
    criterion.enable_batch_metrics(0.5 if compute_metrics else None)
    try:
        for images, masks in dataloader:
            images = images.to(device, non_blocking=True)
            masks = masks.to(device, non_blocking=True)
            if acc is None:
                acc = _EpochAccumulator(images.device)
            optimizer.zero_grad(set_to_none=True)
            loss, x, kind = _step_loss(model, criterion, images, masks)
            acc.add_report(criterion.last_report)
            if compute_metrics:
                with torch.no_grad():
                    _batch_metrics(criterion, x, masks, kind, acc)
            loss.backward()
            optimizer.step()
    finally:
        criterion.enable_batch_metrics(saved_thr)
    if acc is None:
        raise ZeroDivisionError("division by zero")  # the reference divides by num_batches == 0 (src/train.py:169)
    return _results(criterion, acc, return_components, compute_metrics, False, criterion.process_group)


def validate(model, dataloader, criterion, device, return_components: bool = False,
             compute_metrics: bool = True) -> Dict[str, float]:
    """Validate the model (reference src/train.py:188-286: same arguments, same keys in the result; `dice_score` is the
    batch-level thresholded Dice averaged over batches, as there)."""
    criterion = _check_criterion(criterion)
    model.eval()
    acc: Optional[_EpochAccumulator] = None
    saved_thr = criterion.batch_metrics_threshold
    criterion.enable_batch_metrics(0.5)  # the batch-level Dice score is always reported (src/train.py:221-222)
    try:
        with torch.no_grad():
            for images, masks in dataloader:
                images = images.to(device, non_blocking=True)
                masks = masks.to(device, non_blocking=True)
                if acc is None:
                    acc = _EpochAccumulator(images.device)
                _, x, kind = _step_loss(model, criterion, images, masks)
                acc.add_report(criterion.last_report)
                counts = criterion._last_counts                          # float64[B, 4]: sum [u>thr] t, sum [u>thr], sum t
                tot = counts.sum(dim=0)
                acc.batch_dice += (2.0 * tot[0] + 1e-6) / (tot[1] + tot[2] + 1e-6)  # compute_dice_score, src/metrics.py:4-35
                if compute_metrics:
                    _batch_metrics(criterion, x, masks, kind, acc)
    finally:
        criterion.enable_batch_metrics(saved_thr)
    if acc is None:
        raise ZeroDivisionError("division by zero")
    return _results(criterion, acc, return_components, compute_metrics, True, criterion.process_group)
