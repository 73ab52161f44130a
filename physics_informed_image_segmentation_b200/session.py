"""Host-buffer entry point (pil_session_* of include/pil.h): maps in host memory in, loss report and
gradient in host memory out, with the H2D / kernels / D2H pipeline inside the library."""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib
from .functional import LossParams, activation_kind

_NP_X = {np.dtype(np.float32): _lib.F32}
_NP_T = {np.dtype(np.float32): _lib.F32, np.dtype(np.uint8): _lib.U8}


def _host_ptr(a, what: str = "buffer") -> int:
    """Address of a HOST buffer the library reads or writes in place.  Non-contiguous buffers are rejected, never
    copied: a temporary copy would dangle once this function returns (inputs) or swallow the result (gradient)."""
    if isinstance(a, torch.Tensor):
        if a.is_cuda:
            raise ValueError("HostSession takes host tensors; use DiceBCEPDELoss for device tensors")
        if not a.is_contiguous():
            raise ValueError(f"host {what} must be contiguous")
        return a.data_ptr()
    if not isinstance(a, np.ndarray):
        raise TypeError(f"host {what} must be a torch tensor or a numpy array, got {type(a).__name__}")
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError(f"host {what} must be C-contiguous")
    return a.ctypes.data


def _check_buffers(x_host, t_host, grad_host, shape, x_dtype: int, t_dtype: int) -> int:
    """Everything pil_session_run_ex takes on trust: shapes, dtypes, contiguity and writability.  Returns B."""
    shp = tuple(x_host.shape)
    if len(shp) not in (3, 4) or (len(shp) == 4 and shp[1] != 1):
        raise ValueError(f"expected (B,1,H,W) or (B,H,W) maps, got {shp}")
    B = shp[0]
    if shp[-2:] != tuple(shape[1:]) or B > shape[0] or B < 1:
        raise ValueError(f"maps {shp} do not fit the session {tuple(shape)}")
    if tuple(t_host.shape) != shp:
        raise ValueError(f"targets {tuple(t_host.shape)} differ from the maps {shp}")
    if _dtype_code(x_host, False) != x_dtype or _dtype_code(t_host, True) != t_dtype:
        raise TypeError("dtype differs from the session's")
    if grad_host is not None:
        if tuple(grad_host.shape) != shp:
            raise ValueError(f"gradient buffer {tuple(grad_host.shape)} differs from the maps {shp}")
        if _dtype_code(grad_host, False) != x_dtype:
            raise TypeError("the gradient buffer must have the maps' dtype")
        if isinstance(grad_host, np.ndarray) and not grad_host.flags["WRITEABLE"]:
            raise ValueError("the gradient buffer is read-only")
    return int(B)


def _dtype_code(a, is_target: bool) -> int:
    if isinstance(a, torch.Tensor):
        if a.dtype == torch.float32:
            return _lib.F32
        if a.dtype == torch.bfloat16:
            return _lib.BF16
        if is_target and a.dtype in (torch.uint8, torch.bool):
            return _lib.U8
        raise TypeError(f"unsupported dtype {a.dtype}")
    table = _NP_T if is_target else _NP_X
    try:
        return table[np.asarray(a).dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {np.asarray(a).dtype}") from None


class HostSession:
    """Owns the device buffers, streams and events for maps of a fixed (max_B, H, W)."""

    def __init__(self, max_batch: int, height: int, width: int, device: int = 0, x_dtype: int = _lib.F32,
                 t_dtype: int = _lib.F32):
        self._h = ctypes.c_void_p()
        self.shape = (int(max_batch), int(height), int(width))
        self.x_dtype, self.t_dtype = x_dtype, t_dtype
        self._device = int(device)
        st = _lib.lib().pil_session_create(ctypes.byref(self._h), int(device), *self.shape, x_dtype, t_dtype)
        _lib.check(st, "pil_session_create")

    def run(self, x_host, t_host, params: LossParams, grad_host=None, activation: str = "sigmoid",
            grad_on_device: bool = False) -> np.ndarray:
        """One forward+backward.  Returns the loss report (float32[8]: total, dice, bce, rd, pf, ...);
        the gradient w.r.t. x is written into grad_host when given.  With grad_on_device=True (and no
        grad_host) the backward runs too but the gradient stays on the device -- `device_gradient()` views
        it -- the way a training step consumes it; only the report crosses back."""
        params.validate()
        if self._h is None:
            raise RuntimeError("session closed")
        B = _check_buffers(x_host, t_host, grad_host, self.shape, self.x_dtype, self.t_dtype)
        out = np.zeros(_lib.PIL_NOUT, dtype=np.float32)
        cp = params.c()
        flags = _lib.PIL_SESSION_GRAD_ON_DEVICE if (grad_on_device and grad_host is None) else 0
        st = _lib.lib().pil_session_run_ex(self._h, _host_ptr(x_host, "maps"), _host_ptr(t_host, "targets"),
                                           _host_ptr(grad_host, "gradient buffer") if grad_host is not None else None, B,
                                           activation_kind(activation), ctypes.byref(cp), out.ctypes.data, flags)
        _lib.check(st, "pil_session_run_ex")
        self._last_B = B
        return out

    def run_sharded(self, x_host, t_host, params: LossParams, exchange, n_global: int = -1, grad_scale: float = 1.0,
                    activation: str = "sigmoid") -> np.ndarray:
        """Data-parallel step from host buffers: x_host / t_host are THIS rank's images of the global batch and
        `exchange` is the group's sharding.PeerExchange.  Returns the GLOBAL loss report (identical on every rank);
        the gradient of the global loss w.r.t. this rank's maps stays on the device (`device_gradient()`).
        All ranks must call it in lock step."""
        params.validate()
        if self._h is None:
            raise RuntimeError("session closed")
        B = _check_buffers(x_host, t_host, None, self.shape, self.x_dtype, self.t_dtype)
        out = np.zeros(_lib.PIL_NOUT, dtype=np.float32)
        cp = params.c()
        ex = exchange.next_step()
        st = _lib.lib().pil_session_run_xchg(self._h, _host_ptr(x_host, "maps"), _host_ptr(t_host, "targets"), B,
                                             activation_kind(activation), ctypes.byref(cp), ctypes.byref(ex), int(n_global),
                                             float(grad_scale), out.ctypes.data)
        _lib.check(st, "pil_session_run_xchg")
        self._last_B = B
        return out

    def device_gradient(self) -> torch.Tensor:
        """The gradient the last run(grad_on_device=True) left on the device, as a (B,1,H,W) tensor view
        (valid until the next run; clone it to keep it)."""
        if self.x_dtype != _lib.F32:
            raise TypeError("device_gradient() views float32 sessions only")
        ptr = _lib.lib().pil_session_grad_ptr(self._h)
        B, (H, W) = getattr(self, "_last_B", self.shape[0]), self.shape[1:]
        n = B * H * W

        class _Arr:  # __cuda_array_interface__ producer over the session's buffer
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 3}

        return torch.as_tensor(_Arr(), device=torch.device("cuda", self._device)).view(B, 1, H, W)

    def close(self) -> None:
        if self._h is not None and self._h.value:
            _lib.lib().pil_session_destroy(self._h)
        self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
