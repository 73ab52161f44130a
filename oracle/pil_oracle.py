"""ctypes wrapper around oracle/libpil_oracle.so (the plain-C CPU oracle).

TEST INFRASTRUCTURE ONLY -- may be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py, never by the product package.
See oracle/pil_oracle.c for what each function restates (reference file:line).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpil_oracle.so")

X_PROB, X_LOGITS_SIGMOID, X_LOGITS_TANH = 0, 1, 2


class PiloParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in (
        "dice_weight", "bce_weight", "pde_weight", "phase_field_weight",
        "diffusion_coeff", "reaction_threshold", "epsilon", "smooth")]


@dataclass
class Params:
    """Same knobs and defaults as DiceBCEPDELoss.__init__ (src/loss.py:86-96)."""
    dice_weight: float = 0.5
    bce_weight: float = 0.5
    pde_weight: float = 1e-3
    phase_field_weight: float = 0.0
    diffusion_coeff: float = 1.0
    reaction_threshold: float = 0.5
    epsilon: float = 0.05
    smooth: float = 1e-6

    def c(self) -> PiloParams:
        return PiloParams(self.dice_weight, self.bce_weight, self.pde_weight, self.phase_field_weight,
                          self.diffusion_coeff, self.reaction_threshold, self.epsilon, self.smooth)


STAGE2 = Params(pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0,
                reaction_threshold=0.5, epsilon=0.05)  # main.py:14-43 defaults


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("pil_oracle.c", "pil_oracle_impl.inc")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64, dbl, vp, ci = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int
        pp = ctypes.POINTER(PiloParams)
        for suf in ("f32", "f64"):
            getattr(L, f"pilo_laplacian_{suf}").argtypes = [vp, vp, i64, i64, i64]
            getattr(L, f"pilo_grad_mag_sq_{suf}").argtypes = [vp, vp, i64, i64, i64]
            getattr(L, f"pilo_reaction_{suf}").argtypes = [vp, vp, i64, dbl]
            getattr(L, f"pilo_sums_{suf}").argtypes = [vp, vp, i64, i64, i64, ci, pp, vp]
            getattr(L, f"pilo_backward_{suf}").argtypes = [vp, vp, vp, i64, i64, i64, ci, pp, vp, i64, dbl]
            for n in ("laplacian", "grad_mag_sq", "reaction", "sums", "backward"):
                getattr(L, f"pilo_{n}_{suf}").restype = None
        L.pilo_finalize.argtypes = [vp, i64, pp, vp]
        L.pilo_finalize.restype = None
        _lib = L
    return _lib


def _suf(a: np.ndarray) -> str:
    if a.dtype == np.float32:
        return "f32"
    if a.dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle works on float32/float64 arrays, got {a.dtype}")


def _bhw(a: np.ndarray):
    if a.ndim == 4:
        assert a.shape[1] == 1, "single-channel maps only (src/pde.py:45-47 kernels are (1,1,3,3))"
        return a.shape[0], a.shape[2], a.shape[3]
    assert a.ndim == 3
    return a.shape


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def laplacian(u: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(u)
    out = np.empty_like(u)
    B, H, W = _bhw(u)
    getattr(lib(), f"pilo_laplacian_{_suf(u)}")(_ptr(u), _ptr(out), B, H, W)
    return out


def grad_mag_sq(u: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(u)
    out = np.empty_like(u)
    B, H, W = _bhw(u)
    getattr(lib(), f"pilo_grad_mag_sq_{_suf(u)}")(_ptr(u), _ptr(out), B, H, W)
    return out


def reaction(u: np.ndarray, a: float) -> np.ndarray:
    u = np.ascontiguousarray(u)
    out = np.empty_like(u)
    getattr(lib(), f"pilo_reaction_{_suf(u)}")(_ptr(u), _ptr(out), u.size, float(a))
    return out


def sums(x: np.ndarray, t: np.ndarray, p: Params, x_kind: int = X_PROB) -> np.ndarray:
    """Raw sums [I, P, T, sum_bce, sum_r2, sum_pf, n_invalid, n_pixels] of one shard (float64[8])."""
    x = np.ascontiguousarray(x)
    t = np.ascontiguousarray(t, dtype=x.dtype)
    B, H, W = _bhw(x)
    out = np.zeros(8, dtype=np.float64)
    cp = p.c()
    getattr(lib(), f"pilo_sums_{_suf(x)}")(_ptr(x), _ptr(t), B, H, W, x_kind, ctypes.byref(cp), _ptr(out))
    return out


def finalize(s: np.ndarray, n_global: int, p: Params) -> np.ndarray:
    """[total, dice_loss, bce, L_rd, L_pf] (float64[5]) from global sums."""
    s = np.ascontiguousarray(s, dtype=np.float64)
    out = np.zeros(5, dtype=np.float64)
    cp = p.c()
    lib().pilo_finalize(_ptr(s), int(n_global), ctypes.byref(cp), _ptr(out))
    return out


def backward(x: np.ndarray, t: np.ndarray, p: Params, gsums: np.ndarray, n_global: int,
             x_kind: int = X_PROB, grad_scale: float = 1.0) -> np.ndarray:
    x = np.ascontiguousarray(x)
    t = np.ascontiguousarray(t, dtype=x.dtype)
    gsums = np.ascontiguousarray(gsums, dtype=np.float64)
    B, H, W = _bhw(x)
    g = np.empty_like(x)
    cp = p.c()
    getattr(lib(), f"pilo_backward_{_suf(x)}")(_ptr(x), _ptr(t), _ptr(g), B, H, W, x_kind, ctypes.byref(cp),
                                                _ptr(gsums), int(n_global), float(grad_scale))
    return g


def loss_and_grad(x: np.ndarray, t: np.ndarray, p: Params, x_kind: int = X_PROB):
    """Single-shard convenience: (components float64[5], grad like x)."""
    s = sums(x, t, p, x_kind)
    n = int(s[7])
    return finalize(s, n, p), backward(x, t, p, s, n, x_kind)
