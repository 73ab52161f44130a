"""Torch restatement of the reference loss, op for op (the "port" CPU baseline).

TEST INFRASTRUCTURE ONLY -- used by tests/ as a second checker (autograd gives the gradient)
and by bench.py's cpu_baseline / --impl reference legs as the stand-in for the reference's CPU
PyTorch loss on the GPU box, where /root/reference does not exist.  It issues the same ATen op
sequence as the reference (reflect pad -> 1-channel 3x3 conv2d, BCELoss, batch-global Dice), so
its cost profile on host cores is the reference's; tests/test_oracle.py checks that it is
bit-identical to the real reference in the build container.

Reference lines restated: src/pde.py:24-47 (stencils), :67-77 (laplacian), :99 (reaction),
:120,:143 (RD loss), :164-176 (|grad u|^2), :204-210 (PF loss); src/loss.py:130-160 (assembly);
src/unet.py:208-214 (output activation).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_STENCILS = {}


def _stencils(like: torch.Tensor):
    key = (like.dtype, like.device)
    if key not in _STENCILS:
        mk = lambda rows: torch.tensor(rows, dtype=torch.float32).to(like).reshape(1, 1, 3, 3)
        _STENCILS[key] = (
            mk([[0, 1, 0], [1, -4, 1], [0, 1, 0]]),
            mk([[0, 0, 0], [-0.5, 0, 0.5], [0, 0, 0]]),
            mk([[0, -0.5, 0], [0, 0, 0], [0, 0.5, 0]]),
        )
    return _STENCILS[key]


def _mirror(u):
    return F.pad(u, (1, 1, 1, 1), mode="reflect")


def laplacian(u):
    return F.conv2d(_mirror(u), _stencils(u)[0], padding=0)


def reaction(u, a):
    return u * (1.0 - u) * (u - a)


def rd_loss(u, D, a):
    res = D * laplacian(u) + reaction(u, a)
    return torch.mean(res ** 2)


def grad_mag_sq(u):
    up = _mirror(u)
    _, kx, ky = _stencils(u)
    return F.conv2d(up, kx, padding=0) ** 2 + F.conv2d(up, ky, padding=0) ** 2


def pf_loss(u, eps):
    if eps <= 0:
        raise ValueError("epsilon must be positive")
    return torch.mean((eps / 2.0) * grad_mag_sq(u) + (1.0 / eps) * (u ** 2) * ((1.0 - u) ** 2))


def activate(x, kind: int):
    if kind == 0:
        return x
    if kind == 1:
        return torch.sigmoid(x)
    return (torch.tanh(x) + 1.0) / 2.0


def loss_components(u, t, p):
    """p: oracle.pil_oracle.Params.  Returns (total, dice_loss, bce, rd|None, pf|None)."""
    uf, tf = u.reshape(-1), t.reshape(-1)
    inter = (uf * tf).sum()
    dice_loss = 1 - (2.0 * inter + p.smooth) / (uf.sum() + tf.sum() + p.smooth)
    bce = F.binary_cross_entropy(u, t)
    total = p.dice_weight * dice_loss + p.bce_weight * bce
    rd = pf = None
    if p.pde_weight > 0:
        rd = rd_loss(u, p.diffusion_coeff, p.reaction_threshold)
        total = total + p.pde_weight * rd
    if p.phase_field_weight > 0:
        pf = pf_loss(u, p.epsilon)
        total = total + p.phase_field_weight * pf
    return total, dice_loss, bce, rd, pf


def loss(u, t, p):
    return loss_components(u, t, p)[0]


def fwd_bwd(x, t, p, kind: int = 1):
    """One step of the hot path on CPU: activation -> loss -> backward.  Returns (loss, dL/dx)."""
    x = x.detach().requires_grad_(True)
    L = loss(activate(x, kind), t, p)
    L.backward()
    return L.detach(), x.grad
