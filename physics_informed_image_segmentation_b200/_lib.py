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
OBJ_DIR = os.path.join(CSRC, "obj")
HEADERS = [os.path.join(INCLUDE, "pil.h"), os.path.join(CSRC, "pil_common.cuh"), os.path.join(CSRC, "pil_fwdrow.cuh")]

# Translation units of libpil.so: (source, extra defines, object name).  The fused kernels are compiled once per
# input kind (-DPIL_KIND) so that their template instantiations build in parallel.
_KIND_SOURCES = ("pil_fwd.cu", "pil_point.cu", "pil_bwd.cu")
_PLAIN_SOURCES = ("pil_api.cu", "pil_session.cu", "pil_graph.cu", "pil_tail.cu", "pil_boundary.cu")


def _units():
    units = []
    for src in _KIND_SOURCES:
        for k in (0, 1, 2):
            units.append((os.path.join(CSRC, src), [f"-DPIL_KIND={k}"], f"{src[:-3]}_k{k}.o"))
    for src in _PLAIN_SOURCES:
        if os.path.exists(os.path.join(CSRC, src)):
            units.append((os.path.join(CSRC, src), [], f"{src[:-3]}.o"))
    return units


SOURCES = sorted({u[0] for u in _units()})

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
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
PIL_XCHG_DEVICE_EPOCH = 2
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


def build(force: bool = False, verbose: bool = False, extra_flags=(), so_path: str = None) -> str:
    """Compile csrc/*.cu for sm_100a (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo) into csrc/libpil.so
    (in-tree, so it travels with the repo).  Objects are compiled in parallel, one nvcc process per unit;
    only units older than their source or any header are recompiled."""
    so_path = so_path or SO_PATH
    if not force and so_path == SO_PATH and not _stale():
        return SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("libpil.so is missing/stale and nvcc was not found; this package has no CPU fallback")
    from concurrent.futures import ThreadPoolExecutor

    obj_dir = OBJ_DIR if so_path == SO_PATH and not extra_flags else so_path + ".obj"
    os.makedirs(obj_dir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in HEADERS if os.path.exists(h))

    def compile_unit(unit):
        src, defs, obj = unit
        obj = os.path.join(obj_dir, obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, *defs, "-I", INCLUDE, "-I", CSRC, "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {os.path.basename(src)} {' '.join(defs)}:\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=max(1, min(os.cpu_count() or 1, 12))) as pool:
        results = list(pool.map(compile_unit, _units()))
    objs = [r[0] for r in results]
    if verbose:
        print("".join(r[1] for r in results))
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so_path, *objs],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return so_path


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
    "pil_forward_moments_xchg": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_size_t, ctypes.POINTER(PilExchange), ctypes.c_void_p]),
    "pil_sweep_finalize_xchg": (ctypes.c_int, [ctypes.POINTER(PilExchange), ctypes.c_int64, ctypes.POINTER(PilParams), ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pil_sweep_finalize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(PilParams), ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "pil_finalize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(PilParams), ctypes.c_void_p,
                                    ctypes.c_void_p]),
    "pil_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(PilParams),
                                    ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p]),
    "pil_backward_if_scaled": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
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
    "pil_session_run_xchg": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                            ctypes.POINTER(PilParams), ctypes.POINTER(PilExchange), ctypes.c_int64, ctypes.c_float,
                                            ctypes.c_void_p]),
    "pil_exchange_push": (ctypes.c_int, [ctypes.POINTER(PilExchange), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pil_session_grad_ptr": (ctypes.c_void_p, [ctypes.c_void_p]),
    "pil_session_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pil_step_graph_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.POINTER(PilParams), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(PilExchange), ctypes.c_int64,
                                             ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "pil_sweep_graph_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(PilExchange),
                                              ctypes.c_int64, ctypes.POINTER(PilParams), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p]),
    "pil_step_graph_launch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "pil_step_graph_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pil_boundary_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64] * 3),
    "pil_boundary_tolerance_offsets": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "pil_boundary_counts": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "pil_boundary_f1": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "pil_tail_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64]),
    "pil_tail_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                        ctypes.POINTER(PilParams), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.POINTER(PilExchange), ctypes.c_void_p]),
    "pil_tail_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "pil_last_launch_info": (ctypes.c_int, [ctypes.POINTER(PilLaunchInfo)]),
    "pil_set_tuning": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "pil_set_l2_keep_mb": (ctypes.c_int, [ctypes.c_int]),
    "pil_set_bwd_staging": (ctypes.c_int, [ctypes.c_int]),
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
