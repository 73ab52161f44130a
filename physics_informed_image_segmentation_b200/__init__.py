"""B200-native physics-prior segmentation loss: a drop-in for the hot path of
seemapoudel58/Physics_informed_image_segmentation (src/pde.py + src/loss.py + the loss call of
src/train.py), implemented as hand-written sm_100a CUDA kernels behind a C ABI (include/pil.h).

The names re-exported here are the ones the reference's `src/__init__.py:3-4` exports for this path.
"""
from .loss import DiceBCELoss, DiceBCEPDELoss
from .pde import PDERegularization, create_pde_regularization
from .functional import LossParams
from .sharding import shard_bounds, all_reduce_sums, loss_report_from_sums
from .session import HostSession
from .sweep import sweep_losses, s2_grid, s3_grid
from .metrics import (compute_dice_score, compute_dice_score_batch, compute_iou, compute_iou_batch, compute_boundary_f1,
                      compute_boundary_f1_batch)
from .integration import install_into_reference, use_logits_head
from .training import train_epoch, validate

__all__ = [
    "DiceBCELoss", "DiceBCEPDELoss", "PDERegularization", "create_pde_regularization", "LossParams",
    "shard_bounds", "all_reduce_sums", "loss_report_from_sums", "HostSession",
    "install_into_reference", "use_logits_head", "sweep_losses", "s2_grid", "s3_grid",
    "compute_dice_score", "compute_dice_score_batch", "compute_iou", "compute_iou_batch",
    "compute_boundary_f1", "compute_boundary_f1_batch", "train_epoch", "validate",
]
__version__ = "0.1.0"
