"""Generate tests/golden/ref_cases.npz by RUNNING THE REAL REFERENCE (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/src/{pde,loss}.py by path (oracle/ref_loader.py), evaluates
DiceBCEPDELoss / DiceBCELoss / PDERegularization forward and autograd backward in fp32 and in fp64
(`.double()` modules on the same fp32-representable inputs) and stores inputs + outputs.  The
reference ships no golden vectors of its own, so these pin the oracle and,
through it, the CUDA kernels.  torch version at generation time is recorded in the file.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_cases.npz")

STAGE2 = dict(dice_weight=0.5, bce_weight=0.5, pde_weight=1e-4, phase_field_weight=1e-4, smooth=1e-6,
              diffusion_coeff=5.0, reaction_threshold=0.5, epsilon=0.05)
CTOR_DEFAULT = dict(dice_weight=0.5, bce_weight=0.5, pde_weight=1e-3, phase_field_weight=0.0, smooth=1e-6,
                    diffusion_coeff=1.0, reaction_threshold=0.5, epsilon=0.05)


def iid(g, B, H, W):
    z = 2.0 * torch.randn(B, 1, H, W, generator=g)
    t = (torch.rand(B, 1, H, W, generator=g) > 0.5).float()
    return z, t


def blob(g, B, H, W):
    lo = torch.randn(B, 1, max(H // 8, 2), max(W // 8, 2), generator=g)
    sm = torch.nn.functional.interpolate(lo, size=(H, W), mode="bicubic", align_corners=False)
    t = (sm > 0.3).float()
    z = 4.0 * sm - 1.2 + 0.3 * torch.randn(B, 1, H, W, generator=g)
    return z, t


def saturated(g, B, H, W):
    z, t = iid(g, B, H, W)
    vals = torch.tensor([200.0, -200.0, 17.0, -17.0, 30.0, -30.0, 90.0, -90.0, 16.5, -104.0])
    flat = z.view(-1)
    idx = torch.randperm(flat.numel(), generator=g)[: 4 * vals.numel()]
    flat[idx] = vals.repeat(4)
    return z, t


def soft(g, B, H, W):
    z, _ = iid(g, B, H, W)
    return z, torch.rand(B, 1, H, W, generator=g)


CASES = [
    # name, generator, (B,H,W), params, activation
    ("iid_stage2", iid, (2, 16, 20), STAGE2, "sigmoid"),
    ("tiny_2x2", iid, (1, 2, 2), STAGE2, "sigmoid"),
    ("odd_3x5_ctor_default", iid, (2, 3, 5), CTOR_DEFAULT, "sigmoid"),
    ("blob_rd_only", blob, (1, 33, 31), dict(STAGE2, phase_field_weight=0.0, pde_weight=1e-3), "sigmoid"),
    ("blob_pf_only", blob, (1, 17, 19), dict(STAGE2, pde_weight=0.0, epsilon=0.01), "sigmoid"),
    ("saturated", saturated, (2, 8, 8), STAGE2, "sigmoid"),
    ("soft_targets", soft, (3, 24, 40), STAGE2, "sigmoid"),
    ("tanh_head", iid, (1, 12, 12), STAGE2, "tanh"),
    ("dice_bce_only", iid, (2, 9, 7), dict(STAGE2, pde_weight=0.0, phase_field_weight=0.0), "sigmoid"),
    ("big_D_small_eps", blob, (1, 40, 24), dict(STAGE2, diffusion_coeff=100.0, epsilon=0.001, pde_weight=1e-3), "sigmoid"),
    ("wide_2x130", iid, (1, 2, 130), STAGE2, "sigmoid"),
    ("tall_130x2", iid, (1, 130, 2), STAGE2, "sigmoid"),
]


def act(z, name):
    if name == "sigmoid":
        return torch.sigmoid(z)
    return (torch.tanh(z) + 1.0) / 2.0


def run_case(ref_loss, z, t, params, activation, dtype):
    """Reference evaluation: returns dict of numpy arrays."""
    crit = ref_loss.DiceBCEPDELoss(**params)
    if dtype == torch.float64:
        crit = crit.double()
    out = {}
    # logits entry: activation (src/unet.py:208-214) -> loss -> autograd
    zz = z.to(dtype).clone().requires_grad_(True)
    u = act(zz, activation)
    L = crit(u, t.to(dtype))
    L.backward()
    out["loss"] = L.detach().numpy()
    out["dz"] = zz.grad.numpy()
    # probability entry on the fp32-rounded probabilities (what train.py hands the criterion)
    u32 = act(z, activation).detach()
    uu = u32.to(dtype).clone().requires_grad_(True)
    L2 = crit(uu, t.to(dtype))
    L2.backward()
    out["loss_p"] = L2.detach().numpy()
    out["du"] = uu.grad.numpy()
    with torch.no_grad():
        uf, tf = uu.detach().view(-1), t.to(dtype).view(-1)
        inter = (uf * tf).sum()
        dice = 1 - (2.0 * inter + crit.smooth) / (uf.sum() + tf.sum() + crit.smooth)
        comps = [dice, crit.bce(uu.detach(), t.to(dtype)),
                 crit.pde_regularization.compute_loss(uu.detach()),
                 crit.pde_regularization.compute_phase_field_loss(uu.detach(), epsilon=crit.epsilon)]
        out["comps_p"] = np.array([float(c) for c in comps], dtype=np.float64)
        out["lap_p"] = crit.pde_regularization.compute_laplacian(uu.detach()).numpy()
        out["gms_p"] = crit.pde_regularization.compute_gradient_magnitude(uu.detach()).numpy()
    return out


def main():
    assert ref_loader.available(), "needs the reference checkout (build container)"
    ref_loss = ref_loader.loss()
    ref_pde = ref_loader.pde()
    torch.manual_seed(0)
    torch.set_num_threads(1)
    blob_out = {}
    meta = {"torch": torch.__version__, "cases": []}

    # known-answer stencil vectors, re-derived from the reference here
    ka = torch.tensor([[1.0, 2, 4], [3, 5, 9], [7, 8, 6]]).view(1, 1, 3, 3)
    reg = ref_pde.PDERegularization(1.0, 0.5)
    blob_out["ka_u"] = ka.numpy()
    blob_out["ka_lap"] = reg.compute_laplacian(ka).numpy()
    blob_out["ka_gms"] = reg.compute_gradient_magnitude(ka).numpy()
    blob_out["ka_pad1d"] = torch.nn.functional.pad(torch.tensor([[[1.0, 2, 3, 4]]]), (1, 1), mode="reflect").numpy()

    for k, (name, gen, (B, H, W), params, activation) in enumerate(CASES):
        g = torch.Generator().manual_seed(1234 + k)
        z, t = gen(g, B, H, W)
        blob_out[f"{name}.z"] = z.numpy()
        blob_out[f"{name}.t"] = t.numpy()
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            for key, val in run_case(ref_loss, z, t, params, activation, dt).items():
                blob_out[f"{name}.{tag}.{key}"] = val
        # plain DiceBCELoss on the same probabilities (src/loss.py:7-68)
        u32 = act(z, activation)
        uu = u32.clone().requires_grad_(True)
        Lb = ref_loss.DiceBCELoss(params["dice_weight"], params["bce_weight"], params["smooth"])(uu, t)
        Lb.backward()
        blob_out[f"{name}.f32.dicebce_loss"] = Lb.detach().numpy()
        blob_out[f"{name}.f32.dicebce_du"] = uu.grad.numpy()
        meta["cases"].append({"name": name, "shape": [B, H, W], "params": params, "activation": activation})

    blob_out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **blob_out)
    print(f"wrote {OUT}: {os.path.getsize(OUT)} bytes, {len(CASES)} cases, torch {torch.__version__}")


if __name__ == "__main__":
    main()
