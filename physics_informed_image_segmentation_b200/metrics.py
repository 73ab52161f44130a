"""Drop-ins for the reference's thresholded accuracy metrics with the same signatures and defaults:

    compute_dice_score, compute_dice_score_batch      reference src/metrics.py:4-35, :38-73
    compute_iou, compute_iou_batch                    reference src/evaluate.py:26-59, :62-97

The reference binarises the prediction, then loops over the batch in Python with ~6 small kernels per image
and metric (src/train.py:153-160 calls two of them on every training step).  Here all four come from three
counts per image -- sum [p>thr]*t, sum [p>thr], sum t -- produced by ONE pass over the maps
(pil_forward_pointwise_metrics), or by NO extra pass at all when the loss module has just been evaluated on
the same tensors with batch metrics enabled (the counts then ride on the loss's own forward).
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch

from . import functional as Fn
from .functional import LossParams

_NEUTRAL = LossParams(dice_weight=0.5, bce_weight=0.5, pde_weight=0.0, phase_field_weight=0.0)


def threshold_counts(predictions: torch.Tensor, targets: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    """float64[B, 4] per image: sum [p>thr]*t, sum [p>thr], sum t, 0.  Served from the cache of the last
    fused forward when it saw exactly these tensors (same objects, same version counters, same threshold)."""
    cached = Fn.cached_counts(predictions, targets, Fn.X_PROB, threshold)
    if cached is not None:
        return cached
    Fn.check_maps(predictions, targets)
    _, counts = Fn.forward_pointwise_metrics(predictions.detach(), targets.detach(), _NEUTRAL, Fn.X_PROB, threshold)
    Fn.remember_counts(predictions, targets, Fn.X_PROB, threshold, counts)
    return counts


def compute_dice_score_batch(predictions: torch.Tensor, targets: torch.Tensor, threshold: float = 0.5,
                             smooth: float = 1e-6) -> torch.Tensor:
    """Dice of the thresholded prediction for each sample, shape (B,) (reference src/metrics.py:38-73)."""
    dice, _ = Fn.image_metrics(threshold_counts(predictions, targets, threshold), smooth)
    return dice


def compute_iou_batch(predictions: torch.Tensor, targets: torch.Tensor, threshold: float = 0.5,
                      smooth: float = 1e-6) -> torch.Tensor:
    """IoU of the thresholded prediction for each sample, shape (B,) (reference src/evaluate.py:62-97)."""
    _, iou = Fn.image_metrics(threshold_counts(predictions, targets, threshold), smooth)
    return iou


def _global_counts(predictions, targets, threshold) -> torch.Tensor:
    return threshold_counts(predictions, targets, threshold).sum(dim=0, keepdim=True)


def compute_dice_score(predictions: torch.Tensor, targets: torch.Tensor, threshold: float = 0.5,
                       smooth: float = 1e-6) -> torch.Tensor:
    """Dice of the thresholded prediction over the whole batch, 0-dim (reference src/metrics.py:4-35)."""
    dice, _ = Fn.image_metrics(_global_counts(predictions, targets, threshold), smooth)
    return dice[0]


def compute_iou(predictions: torch.Tensor, targets: torch.Tensor, threshold: float = 0.5,
                smooth: float = 1e-6) -> torch.Tensor:
    """IoU of the thresholded prediction over the whole batch, 0-dim (reference src/evaluate.py:26-59)."""
    _, iou = Fn.image_metrics(_global_counts(predictions, targets, threshold), smooth)
    return iou[0]


def compute_boundary_f1_batch(predictions: torch.Tensor, targets: torch.Tensor, threshold: float = 0.5,
                              tolerance: int = 2, smooth: float = 1e-6) -> torch.Tensor:
    """Boundary-F1 within `tolerance` pixels for each sample, shape (B,) (reference src/evaluate.py:196-229).

    The reference pulls every image to the host and runs OpenCV on it (findContours + drawContours, two
    distanceTransform calls) -- B device-to-host syncs on every training step (src/train.py:156).  Here the same
    boundary pixels and chamfer neighbourhoods are counted by CUDA kernels (include/pil.h pil_boundary_counts); the
    result stays on the predictions' device (the reference returns a CPU tensor; `.cpu()` works on both)."""
    counts = Fn.boundary_counts(predictions.detach(), targets.detach(), Fn.X_PROB, threshold, int(tolerance))
    return Fn.boundary_f1(counts, int(tolerance), smooth)


def compute_boundary_f1(predictions: torch.Tensor, targets: torch.Tensor, threshold: float = 0.5,
                        tolerance: int = 2, smooth: float = 1e-6) -> torch.Tensor:
    """Boundary-F1 of the FIRST sample, 0-dim -- the reference looks at `predictions[0, 0]` only
    (src/evaluate.py:149-150)."""
    return compute_boundary_f1_batch(predictions[:1].contiguous(), targets[:1].contiguous(), threshold, tolerance, smooth)[0]
