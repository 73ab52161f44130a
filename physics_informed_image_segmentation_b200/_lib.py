"""Loader and thin ctypes binding of the C-ABI library (include/pil.h).

There is no CPU fallback and no alternate backend: if libpil.so is missing (and cannot be built
because nvcc is absent) importing the compute entry points fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
SO_PATH = os.path.join(CSRC, "libpil.so")
SOURCES = [os.path.join(CSRC, "pil_kernels.cu"), os.path.join(CSRC, "pil_session.cu")]
HEADERS = [os.path.join(INCLUDE, "pil.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]

PIL_NSUMS = 8
PIL_NOUT = 8
PIL_NMOMENTS = 16
F32, BF16, U8 = 0, 1, 2
X_PROB, X_LOGITS_SIGMOID, X_LOGITS_TANH = 0, 1, 2


class PilParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in (
        "dice_weight", "bce_weight", "pde_weight", "phase_field_weight",
        "diffusion_coeff", "reaction_threshold", "epsilon", "smooth")]


class PilLaunchInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "fwd_blocks", "fwd_threads", "fwd_rows_per_segment", "fwd_aligned",
        "bwd_blocks", "bwd_threads", "bwd_rows_per_segment", "bwd_aligned")] + [("kernels_launched", ctypes.c_int64)]


PIL_MAX_RANKS = 8
PIL_IPC_HANDLE_BYTES = 64
PIL_XCHG_DEFER_FINALIZE = 1
PIL_SESSION_GRAD_ON_DEVICE = 1


class PilExchange(ctypes.Structure):
    """include/pil.h PilExchange: the peer-memory mailboxes of a data-parallel group."""
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("epoch", ctypes.c_uint64),
                ("flags", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
                ("mailbox", ctypes.c_void_p * PIL_MAX_RANKS)]


class PilError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = _lib.pil_status_string(status).decode() if _lib is not None else str(status)
        super().__init__(f"{where}: {msg} (status {status})")


def _stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into csrc/libpil.so (in-tree, so it travels with the repo)."""
    if not force and not _stale():
        return SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("libpil.so is missing/stale and nvcc was not found; this package has no CPU fallback")
    srcs = [s for s in SOURCES if os.path.exists(s)]
    cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC, "-o", SO_PATH, *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return SO_PATH


_lib = None

_SIGS = {
    # name: (restype, argtypes)
    "pil_version": (ctypes.c_int, []),
    "pil_status_string": (ctypes.c_char_p, [ctypes.c_int]),
    "pil_validate_params": (ctypes.c_int, [ctypes.POINTER(PilParams)]),
    "pil_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64] * 3),
    "pil_workspace_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "pil_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(PilParams),
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "pil_forward_pointwise_metrics": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(PilParams),
                                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p,
                                                     ctypes.c_size_t, ctypes.POINTER(PilExchange), ctypes.c_void_p]),
    "pil_image_metrics": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "pil_forward_moments": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_size_t, ctypes.c_void_p]),
    "pil_sweep_finalize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(PilParams), ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "pil_finalize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(PilParams), ctypes.c_void_p,
                                    ctypes.c_void_p]),
    "pil_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(PilParams),
                                    ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p]),
    "pil_forward_pointwise": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(PilParams),
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "pil_backward_accumulate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                               ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(PilParams),
                                               ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_float,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "pil_loss_fwd_bwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(PilParams),
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "pil_exchange_bytes": (ctypes.c_size_t, []),
    "pil_exchange_alloc": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]),
    "pil_exchange_open": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "pil_exchange_close": (ctypes.c_int, [ctypes.c_void_p]),
    "pil_exchange_free": (ctypes.c_int, [ctypes.c_void_p]),
    "pil_exchange_status": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]),
    "pil_exchange_finalize": (ctypes.c_int, [ctypes.POINTER(PilExchange), ctypes.c_int64, ctypes.POINTER(PilParams),
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pil_forward_pointwise_xchg": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(PilParams),
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                  ctypes.POINTER(PilExchange), ctypes.c_void_p]),
    "pil_backward_accumulate_xchg": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                    ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                    ctypes.POINTER(PilParams), ctypes.POINTER(PilExchange), ctypes.c_int64,
                                                    ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "pil_scale_gradient": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "pil_laplacian": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_void_p]),
    "pil_laplacian_adjoint": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                             ctypes.c_int64, ctypes.c_void_p]),
    "pil_reaction": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p]),
    "pil_grad_mag_sq": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_void_p]),
    "pil_grad_mag_sq_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    "pil_session_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "pil_session_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int64, ctypes.c_int, ctypes.POINTER(PilParams), ctypes.c_void_p]),
    "pil_session_run_ex": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int64, ctypes.c_int, ctypes.POINTER(PilParams), ctypes.c_void_p, ctypes.c_int]),
    "pil_session_grad_ptr": (ctypes.c_void_p, [ctypes.c_void_p]),
    "pil_session_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pil_last_launch_info": (ctypes.c_int, [ctypes.POINTER(PilLaunchInfo)]),
    "pil_set_tuning": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "pil_set_l2_keep_mb": (ctypes.c_int, [ctypes.c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


def lib() -> ctypes.CDLL:
    """The loaded library; builds it first when sources are newer and nvcc is available."""
    global _lib
    if _lib is None:
        override = os.environ.get("PIL_LIB")  # development: an instrumented build of the same sources
        if override:
            L = ctypes.CDLL(override)
        else:
            try:
                build()
            except RuntimeError:
                if not os.path.exists(SO_PATH):
                    raise
            L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError here == the library does not export what pil.h declares
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int, where: str) -> None:
    if status != 0:
        lib()
        raise PilError(status, where)
