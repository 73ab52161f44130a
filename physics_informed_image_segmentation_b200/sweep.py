"""Parameter sweeps as ONE batched loss evaluation (BASELINE config 4).

The reference explores its knobs by training one model per setting, variants run one after the other
(reference run_ablation.py:159-224 define the S2 diffusion-coefficient grid D in {0.5, 1, 2, 5, 10, 100}
and the S3 interface-width grid eps in {0.001, 0.01, 0.05, 0.1, 0.2}; src/ablation.py:1302-1324 loops over
them).  Evaluating the LOSS of a batch under all settings needs one pass over the maps, not K: the
stencils, the cubic's factors and the Dice/BCE sums do not depend on the knobs (include/pil.h
pil_forward_moments), so the K losses come out of 13 sums.
"""
from __future__ import annotations

from dataclasses import replace
from typing import List, Sequence

import torch

from . import functional as Fn
from .functional import LossParams

S2_DIFFUSION_COEFFS = (0.5, 1.0, 2.0, 5.0, 10.0, 100.0)   # reference run_ablation.py:176-188
S3_EPSILONS = (0.001, 0.01, 0.05, 0.1, 0.2)                # reference run_ablation.py:210-224


def s2_grid(base: LossParams = LossParams()) -> List[LossParams]:
    """S2: reaction-diffusion only, pde_weight = 1e-3, D varies (reference run_ablation.py:159-188)."""
    return [replace(base, pde_weight=1e-3, phase_field_weight=0.0, diffusion_coeff=d) for d in S2_DIFFUSION_COEFFS]


def s3_grid(base: LossParams = LossParams()) -> List[LossParams]:
    """S3: RD + phase field, both weights 1e-4, D = 5, a = 0.5, eps varies (reference run_ablation.py:191-224)."""
    return [replace(base, pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0, reaction_threshold=0.5, epsilon=e)
            for e in S3_EPSILONS]


def sweep_losses(predictions: torch.Tensor, targets: torch.Tensor, grid: Sequence[LossParams], activation: str = "none",
                 process_group=None) -> torch.Tensor:
    """float32[K, 8] loss reports (total, dice_loss, bce, L_rd, L_pf, n_invalid, -, -), one row per
    setting of `grid`, from a single read of the maps.  `activation`: 'none' (probabilities, what the
    reference's loss receives), 'sigmoid' or 'tanh' (logits; src/unet.py:208-214 fused)."""
    return Fn.sweep_losses(predictions, targets, list(grid), Fn.activation_kind(activation), group=process_group)
