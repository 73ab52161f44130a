"""Shared helpers for the parity tests: synthetic inputs and error metrics."""
from __future__ import annotations

import numpy as np
import torch


def iid_inputs(B, H, W, seed=1234):
    g = torch.Generator().manual_seed(seed)
    z = 2.0 * torch.randn(B, 1, H, W, generator=g)
    t = (torch.rand(B, 1, H, W, generator=g) > 0.5).float()
    return z, t


def blob_inputs(B, H, W, seed=1234):
    g = torch.Generator().manual_seed(seed)
    lo = torch.randn(B, 1, max(H // 16, 2), max(W // 16, 2), generator=g)
    sm = torch.nn.functional.interpolate(lo, size=(H, W), mode="bicubic", align_corners=False)
    t = (sm > 0.3).float()
    z = 4.0 * sm - 1.2 + 0.3 * torch.randn(B, 1, H, W, generator=g)
    return z, t


def rel_max(a, ref):
    """max|a-ref| / max|ref|  -- the headline gradient metric of SURVEY.md 8c."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.max(np.abs(ref))
    return float(np.max(np.abs(a - ref)) / (den if den > 0 else 1.0))


def rel_l2(a, ref):
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.linalg.norm(ref)
    return float(np.linalg.norm(a - ref) / (den if den > 0 else 1.0))


def rel_scalar(a, ref):
    a, ref = float(a), float(ref)
    return abs(a - ref) / max(abs(ref), 1e-30)
