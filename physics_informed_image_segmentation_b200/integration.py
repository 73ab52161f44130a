"""Glue for running the UNMODIFIED reference training code on top of this package.

`install_into_reference()` rebinds the four names the reference exports for this path
(`DiceBCELoss`, `DiceBCEPDELoss`, `PDERegularization`, `create_pde_regularization`; reference
src/__init__.py:3-4) in every already-imported module of the reference package, so that
src/train.py:656,:708 and src/ablation.py:102,:146 construct the fused modules and the
`isinstance(criterion, DiceBCEPDELoss)` checks at src/train.py:142,:174 keep working.

`use_logits_head(model)` makes the reference UNet return logits without touching its code or its
state dict: UNet.forward applies the activation under `if self.activation_name == ...` tests
(src/unet.py:208-214), so any other value skips it.
"""
from __future__ import annotations

import contextlib
import sys

from . import metrics as _metrics
from .loss import DiceBCELoss, DiceBCEPDELoss, _FusedLossBase
from .pde import PDERegularization, create_pde_regularization

_NAMES = {
    "DiceBCELoss": DiceBCELoss,
    "DiceBCEPDELoss": DiceBCEPDELoss,
    "PDERegularization": PDERegularization,
    "create_pde_regularization": create_pde_regularization,
}
# the thresholded accuracy metrics train_epoch / validate call on every step (src/train.py:154-155, :257-258)
_METRIC_NAMES = {
    "compute_dice_score": _metrics.compute_dice_score,
    "compute_dice_score_batch": _metrics.compute_dice_score_batch,
    "compute_iou": _metrics.compute_iou,
    "compute_iou_batch": _metrics.compute_iou_batch,
    "compute_boundary_f1": _metrics.compute_boundary_f1,
    "compute_boundary_f1_batch": _metrics.compute_boundary_f1_batch,
}


def install_into_reference(package: str = "src", track_metrics: bool = True, fused_step: bool = True) -> list:
    """Rebind the loss/PDE names inside the imported reference package.  Returns the list of
    (module, name) pairs that were replaced.  With track_metrics (default) the thresholded Dice/IoU and boundary-F1
    functions are rebound too and every criterion built afterwards leaves per-image threshold counts
    (threshold 0.5, the value the reference passes), so those per-step metric calls cost no pass over the maps.
    With fused_step (default) `train_epoch` and `validate` (src/train.py:84-286) are replaced by training.train_epoch /
    training.validate: logits head + fused activation, components and metrics from the loss's own kernels, one host
    sync per epoch; train_stage (src/train.py:289-391) picks them up through the module's globals."""
    replaced = []
    names = dict(_NAMES)
    if track_metrics:
        names.update(_METRIC_NAMES)
        _FusedLossBase.default_batch_metrics_threshold = 0.5
    if fused_step:
        from . import training as _training

        names.update({"train_epoch": _training.train_epoch, "validate": _training.validate})
    _names_backup = names
    for modname, mod in list(sys.modules.items()):
        if mod is None or not (modname == package or modname.startswith(package + ".")):
            continue
        for name, new in _names_backup.items():
            if hasattr(mod, name) and getattr(mod, name) is not new:
                setattr(mod, name, new)
                replaced.append((modname, name))
    return replaced


@contextlib.contextmanager
def use_logits_head(model):
    """Within the context `model(images)` returns logits; pair with criterion.forward_logits(...,
    activation=<the restored name>).  Yields the original activation name."""
    original = model.activation_name
    model.activation_name = "none"
    try:
        yield original
    finally:
        model.activation_name = original
