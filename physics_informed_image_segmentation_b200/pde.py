"""Drop-in for the reference's src/pde.py: same class, same method names, same errors.

Every method runs CUDA kernels from libpil.so; the two scalar losses go through the fused
forward/backward kernels (with only their own term switched on), the map-valued operators through
the stand-alone stencil kernels.  CUDA tensors only.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as Fn
from .functional import LossParams


class PDERegularization(nn.Module):
    """Reaction-diffusion residual and phase-field energy operators (reference src/pde.py:6-212)."""

    def __init__(self, diffusion_coeff: float = 1.0, reaction_threshold: float = 0.5):
        super().__init__()
        # reference src/pde.py:14-17
        if diffusion_coeff <= 0:
            raise ValueError("diffusion_coeff must be positive")
        if not (0 < reaction_threshold < 1):
            raise ValueError("reaction_threshold must be in (0,1)")
        self.diffusion_coeff = diffusion_coeff
        self.reaction_threshold = reaction_threshold
        # The reference registers its three 3x3 stencils as buffers (src/pde.py:45-47), so they are part
        # of the criterion's state_dict; keep the same keys/shapes so state dicts interchange.  The
        # kernels have the stencils baked in and never read these.
        self.register_buffer("laplacian_kernel",
                             torch.tensor([[0.0, 1.0, 0.0], [1.0, -4.0, 1.0], [0.0, 1.0, 0.0]]).view(1, 1, 3, 3))
        self.register_buffer("grad_x_kernel",
                             torch.tensor([[0.0, 0.0, 0.0], [-0.5, 0.0, 0.5], [0.0, 0.0, 0.0]]).view(1, 1, 3, 3))
        self.register_buffer("grad_y_kernel",
                             torch.tensor([[0.0, -0.5, 0.0], [0.0, 0.0, 0.0], [0.0, 0.5, 0.0]]).view(1, 1, 3, 3))
        self._owner = None  # set by DiceBCEPDELoss so logging calls can be served from its last forward

    # -- map-valued operators ---------------------------------------------------------------------
    def compute_laplacian(self, u: torch.Tensor) -> torch.Tensor:
        """5-point Laplacian with mirror (Neumann) boundary; reference src/pde.py:49-79."""
        return Fn.laplacian(u)

    def reaction_term(self, u: torch.Tensor) -> torch.Tensor:
        """f(u) = u(1-u)(u-a); reference src/pde.py:81-99."""
        return Fn.reaction(u, self.reaction_threshold)

    def compute_residual(self, u: torch.Tensor) -> torch.Tensor:
        """r = D lap(u) + f(u); reference src/pde.py:101-122."""
        return self.diffusion_coeff * self.compute_laplacian(u) + self.reaction_term(u)

    def compute_gradient_magnitude(self, u: torch.Tensor) -> torch.Tensor:
        """|grad u|^2 by central differences on the mirror-padded map; reference src/pde.py:147-178."""
        return Fn.grad_mag_sq(u)

    # -- scalar losses (fused kernels) ------------------------------------------------------------
    def _params(self, **kw) -> LossParams:
        base = dict(dice_weight=0.0, bce_weight=0.0, pde_weight=0.0, phase_field_weight=0.0,
                    diffusion_coeff=self.diffusion_coeff, reaction_threshold=self.reaction_threshold)
        base.update(kw)
        return LossParams(**base)

    def compute_loss(self, u: torch.Tensor) -> torch.Tensor:
        """mean(r^2); reference src/pde.py:124-145."""
        cached = self._from_owner(u, Fn.OUT_RD, None)
        if cached is not None:
            return cached
        # the kernels read a target map; with every target-dependent weight at 0 the prediction map
        # itself is a valid stand-in (same size, already resident in cache)
        out, _ = Fn.fused_loss(u, u, self._params(pde_weight=1.0), Fn.X_PROB, Fn.OUT_RD, eager=False)
        return out

    def compute_phase_field_loss(self, u: torch.Tensor, epsilon: float = 0.05) -> torch.Tensor:
        """mean((eps/2)|grad u|^2 + u^2(1-u)^2/eps); reference src/pde.py:180-212."""
        if epsilon <= 0:
            raise ValueError("epsilon must be positive")
        cached = self._from_owner(u, Fn.OUT_PF, epsilon)
        if cached is not None:
            return cached
        out, _ = Fn.fused_loss(u, u, self._params(phase_field_weight=1.0, epsilon=epsilon), Fn.X_PROB, Fn.OUT_PF, eager=False)
        return out

    def _from_owner(self, u: torch.Tensor, which: int, epsilon):
        """Logging re-evaluation (reference src/train.py:142-149) served from the owner's last fused
        forward when it was on this very tensor -- zero extra passes.  Only without autograd."""
        owner = self._owner() if self._owner is not None else None
        if owner is None or (torch.is_grad_enabled() and u.requires_grad):
            return None
        return owner._cached_component(u, which, epsilon)


def create_pde_regularization(diffusion_coeff: float = 1.0, reaction_threshold: float = 0.5) -> PDERegularization:
    """Factory; reference src/pde.py:215-232."""
    return PDERegularization(diffusion_coeff=diffusion_coeff, reaction_threshold=reaction_threshold)
