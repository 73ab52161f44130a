"""Drop-in for the reference's src/loss.py: DiceBCELoss and DiceBCEPDELoss with the same constructor
signatures, attribute surface and call convention, backed by the two fused sm_100a kernels.

    criterion = DiceBCEPDELoss(dice_weight=0.5, bce_weight=0.5, pde_weight=1e-4, phase_field_weight=1e-4,
                               diffusion_coeff=5.0, reaction_threshold=0.5, epsilon=0.05).to(device)
    loss = criterion(outputs, masks)      # outputs: probabilities (B,1,H,W), exactly like src/train.py:117
    loss.backward()                       # one fused backward kernel writes dL/d(outputs)

Extra (not in the reference): `forward_logits(logits, masks, activation="sigmoid"|"tanh")` fuses the
model's output activation (src/unet.py:208-214) into both kernels and returns dL/dlogits, and
`process_group=` shards the batch over ranks with one 64-byte all-reduce of the partial sums.
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch
import torch.nn as nn

from . import functional as Fn
from .functional import LossParams
from .pde import PDERegularization


class _FusedBCE(nn.Module):
    """Stands where the reference keeps `self.bce = nn.BCELoss()` (src/loss.py:34, :112): callers invoke
    `criterion.bce(outputs, masks)` for logging (src/train.py:135, :238)."""

    def __init__(self, owner):
        super().__init__()
        self._owner = weakref.ref(owner)

    def forward(self, predictions: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        owner = self._owner()
        if owner is not None and not (torch.is_grad_enabled() and predictions.requires_grad):
            cached = owner._cached_component(predictions, Fn.OUT_BCE, None, targets)
            if cached is not None:
                return cached
        p = LossParams(dice_weight=0.0, bce_weight=1.0, pde_weight=0.0, phase_field_weight=0.0)
        out, _ = Fn.fused_loss(predictions, targets, p, Fn.X_PROB, Fn.OUT_BCE, eager=False)  # logging call: gradient on demand
        return out


class _FusedLossBase(nn.Module):
    # Threshold new instances start with (None = off).  integration.install_into_reference(track_metrics=True)
    # sets it to 0.5 so that criteria built INSIDE the unmodified reference code produce the per-image counts
    # its per-step compute_dice_score_batch / compute_iou_batch calls are then served from.
    default_batch_metrics_threshold: Optional[float] = None

    def _init_runtime(self, process_group, ddp_average: bool):
        self.process_group = process_group
        self.ddp_average = ddp_average
        self._last = None  # (key, report) of the most recent fused forward
        self.last_report: Optional[torch.Tensor] = None
        # Per-step accuracy metrics (reference src/train.py:153-160): set to a threshold (0.5 in the
        # reference) to have every forward also leave per-image threshold counts; None = off.
        self.batch_metrics_threshold: Optional[float] = type(self).default_batch_metrics_threshold
        # Probabilities outside [0,1] (or NaN): the reference's nn.BCELoss raises.  The kernels count them and
        # return a NaN loss without a host sync; with strict_inputs=True the module also raises the reference's
        # RuntimeError, at the price of one sync per call (the reference's loop syncs anyway at loss.item()).
        self.strict_inputs: bool = False
        # False (default): a grad-enabled forward also runs the backward kernel, so the stencils are evaluated once per
        # step and loss.backward() is free -- at the price of a full-size gradient buffer per forward even if
        # .backward() never comes.  True: forward only; the gradient is computed when autograd asks for it.
        self.lazy_backward: bool = False
        self._last_counts: Optional[torch.Tensor] = None

    def _params(self) -> LossParams:
        raise NotImplementedError

    def _run(self, x: torch.Tensor, t: torch.Tensor, kind: int) -> torch.Tensor:
        p = self._params()
        loss, report, counts = Fn.fused_loss_with_counts(x, t, p, kind, Fn.OUT_TOTAL, self.process_group, self.ddp_average,
                                                         self.batch_metrics_threshold, eager=not self.lazy_backward)
        self.last_report = report
        self._last_counts = counts
        if self.strict_inputs and kind == Fn.X_PROB and report[Fn.OUT_INVALID].item() > 0:
            raise RuntimeError("all elements of input should be between 0 and 1")  # nn.BCELoss's message
        # identity (weak) + version counter: a different tensor that merely reuses the address never hits
        self._last = (weakref.ref(x), x._version, weakref.ref(t), t._version, kind, p, report)
        return loss

    def _cached_component(self, x: torch.Tensor, which: int, epsilon, t: Optional[torch.Tensor] = None):
        """Component `which` of the last fused forward if it was evaluated on exactly this tensor
        (same storage, same version counter) with the same physics knobs; else None."""
        if self._last is None:
            return None
        wx, vx, wt, vt, kind, p, report = self._last
        if kind != Fn.X_PROB or wx() is not x or x._version != vx:
            return None
        if t is not None and (wt() is not t or t._version != vt):
            return None
        if which == Fn.OUT_PF and (epsilon is None or float(epsilon) != float(p.epsilon)):
            return None
        return report[which]

    def components(self) -> dict:
        """Loss terms of the most recent forward (device scalars, no sync): what src/train.py:120-150
        recomputes with a second pass."""
        if self.last_report is None:
            raise RuntimeError("no forward pass has run yet")
        r = self.last_report
        return {"loss": r[0], "dice_loss": r[1], "bce_loss": r[2], "pde_loss": r[3], "phase_field_loss": r[4]}

    def enable_batch_metrics(self, threshold: Optional[float] = 0.5) -> "_FusedLossBase":
        """Have every forward also produce the reference's per-image Dice / IoU of the thresholded prediction
        (compute_dice_score_batch, src/metrics.py:38-73; compute_iou_batch, src/evaluate.py:62-97) from
        the same read of the maps; `None` switches it off."""
        self.batch_metrics_threshold = threshold
        return self

    def last_batch_metrics(self, smooth: float = 1e-6) -> dict:
        """{'dice': float32[B], 'iou': float32[B]} of this rank's images in the most recent forward
        (device tensors, no sync) -- what src/train.py:155-156 computes with 2 x B Python iterations."""
        if self._last_counts is None:
            raise RuntimeError("no forward pass with batch metrics enabled has run yet (enable_batch_metrics())")
        dice, iou = Fn.image_metrics(self._last_counts, smooth)
        return {"dice": dice, "iou": iou}

    def forward(self, predictions: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        return self._run(predictions, targets, Fn.X_PROB)

    def forward_logits(self, logits: torch.Tensor, targets: torch.Tensor, activation: str = "sigmoid") -> torch.Tensor:
        """Same loss evaluated on activation(logits) with the activation fused into the kernels; the
        backward kernel then writes dL/dlogits directly (replaces src/unet.py:208-214 + SigmoidBackward)."""
        return self._run(logits, targets, Fn.activation_kind(activation))


    def forward_features(self, features: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                         targets: torch.Tensor, activation: str = "sigmoid") -> torch.Tensor:
        """The model tail fused in as well: `features` is the U-Net's C-channel full-resolution map
        (contiguous NCHW), `weight` / `bias` the parameters of its 1x1 output convolution (src/unet.py:157, :205).  One
        pass over the features computes the logits and the pointwise loss sums, the fused backward kernel the gradient
        with respect to the logits, and loss.backward() one more pass for dL/dfeatures, dL/dweight and dL/dbias --
        replacing out_conv, the activation, the loss passes over the logits and the convolution backward's three
        kernels.  `last_logits` holds the fp32 logits (detached) for metrics."""
        p = self._params()
        loss, report, logits = Fn.fused_tail_loss(features, weight, bias, targets, p, Fn.activation_kind(activation),
                                                  self.process_group, self.ddp_average)
        self.last_report, self.last_logits = report, logits
        self._last = None
        if self.batch_metrics_threshold is not None:
            _, self._last_counts = Fn.forward_pointwise_metrics(logits, targets.detach(), p, Fn.activation_kind(activation),
                                                                self.batch_metrics_threshold)
        return loss


class DiceBCELoss(_FusedLossBase):
    """Batch-global soft Dice + mean BCE (reference src/loss.py:7-68)."""

    def __init__(self, dice_weight: float = 0.5, bce_weight: float = 0.5, smooth: float = 1e-6,
                 process_group=None, ddp_average: bool = True):
        super().__init__()
        self.dice_weight = dice_weight
        self.bce_weight = bce_weight
        self.smooth = smooth
        self.bce = _FusedBCE(self)
        self._init_runtime(process_group, ddp_average)

    def _params(self) -> LossParams:
        return LossParams(dice_weight=self.dice_weight, bce_weight=self.bce_weight, pde_weight=0.0,
                          phase_field_weight=0.0, smooth=self.smooth)


class DiceBCEPDELoss(_FusedLossBase):
    """Dice + BCE + lambda_rd * RD residual + lambda_pf * phase-field energy (reference src/loss.py:71-162).

    The weights are read at call time (plain attributes, as in the reference), `pde_weight > 0` and
    `phase_field_weight > 0` gate their terms (src/loss.py:150, :155), epsilon is validated only when
    the phase-field term is active (src/pde.py:199-200)."""

    def __init__(self, dice_weight: float = 0.5, bce_weight: float = 0.5, pde_weight: float = 1e-3,
                 phase_field_weight: float = 0.0, smooth: float = 1e-6, diffusion_coeff: float = 1.0,
                 reaction_threshold: float = 0.5, epsilon: float = 0.05,
                 process_group=None, ddp_average: bool = True):
        super().__init__()
        self.dice_weight = dice_weight
        self.bce_weight = bce_weight
        self.pde_weight = pde_weight
        self.phase_field_weight = phase_field_weight
        self.smooth = smooth
        self.epsilon = epsilon
        self.pde_regularization = PDERegularization(diffusion_coeff=diffusion_coeff,
                                                    reaction_threshold=reaction_threshold)
        self.pde_regularization._owner = weakref.ref(self)
        self.bce = _FusedBCE(self)
        self._init_runtime(process_group, ddp_average)

    def _params(self) -> LossParams:
        reg = self.pde_regularization
        return LossParams(dice_weight=self.dice_weight, bce_weight=self.bce_weight, pde_weight=self.pde_weight,
                          phase_field_weight=self.phase_field_weight, smooth=self.smooth,
                          diffusion_coeff=reg.diffusion_coeff, reaction_threshold=reg.reaction_threshold,
                          epsilon=self.epsilon)
