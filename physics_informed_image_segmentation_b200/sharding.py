"""Data-parallel plumbing for the loss: batch sharding and the one exchange step.

The batch shards by whole images (stencils stop at each image's mirror boundary, reference
src/pde.py:67, so there is no halo exchange).  The only coupling is the batch-global reductions of
reference src/loss.py:134-141 and src/pde.py:143,:210: 8 doubles, all-reduced (SUM) between the
forward and the backward kernel.  Works with any torch.distributed backend (NCCL on the GPUs, gloo
in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from .functional import LossParams


def shard_bounds(batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[b0, b1) of the images rank `rank` owns: contiguous, sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if batch < 0:
        raise ValueError("batch must be >= 0")
    base, rem = divmod(batch, world_size)
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def all_reduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of a sums vector (float64[8]) over the data-parallel group."""
    import torch.distributed as dist

    if sums.dtype != torch.float64 or sums.numel() != 8:
        raise ValueError("sums must be float64[8] (include/pil.h PIL_NSUMS)")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def loss_report_from_sums(sums: Sequence[float], n_global: Optional[int], p: LossParams) -> dict:
    """Host-side assembly of the scalar loss from (all-reduced) sums -- the same arithmetic as the
    device finalize (reference src/loss.py:134-160), for logging and for tests that have no GPU."""
    s = [float(v) for v in sums]
    n = float(n_global) if n_global and n_global > 0 else s[7]
    dice_loss = 1.0 - (2.0 * s[0] + p.smooth) / (s[1] + s[2] + p.smooth)
    bce, rd, pf = s[3] / n, s[4] / n, s[5] / n
    total = p.dice_weight * dice_loss + p.bce_weight * bce
    if p.pde_weight > 0:
        total += p.pde_weight * rd
    if p.phase_field_weight > 0:
        total += p.phase_field_weight * pf
    return {"loss": total, "dice_loss": dice_loss, "bce_loss": bce, "pde_loss": rd, "phase_field_loss": pf,
            "n_invalid": s[6], "n_pixels": n}
