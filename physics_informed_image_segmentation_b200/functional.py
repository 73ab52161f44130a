"""Torch-tensor level wrappers over the C ABI (include/pil.h) and the autograd glue.

PyTorch is plumbing here (device memory, streams, torch.distributed); every arithmetic pass over the
maps happens inside libpil.so.  CUDA tensors only -- there is no CPU path.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, replace
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import F32, BF16, U8, X_PROB, X_LOGITS_SIGMOID, X_LOGITS_TANH, PIL_NMOMENTS, PIL_NOUT, PIL_NSUMS, PilParams

OUT_TOTAL, OUT_DICE, OUT_BCE, OUT_RD, OUT_PF, OUT_INVALID = 0, 1, 2, 3, 4, 5

_ACTIVATION_KIND = {"none": X_PROB, "prob": X_PROB, "sigmoid": X_LOGITS_SIGMOID, "tanh": X_LOGITS_TANH}


@dataclass(frozen=True)
class LossParams:
    """Knobs of DiceBCEPDELoss.__init__ (reference src/loss.py:86-96), same names and defaults."""
    dice_weight: float = 0.5
    bce_weight: float = 0.5
    pde_weight: float = 1e-3
    phase_field_weight: float = 0.0
    smooth: float = 1e-6
    diffusion_coeff: float = 1.0
    reaction_threshold: float = 0.5
    epsilon: float = 0.05

    def c(self) -> PilParams:
        return PilParams(float(self.dice_weight), float(self.bce_weight), float(self.pde_weight),
                         float(self.phase_field_weight), float(self.diffusion_coeff),
                         float(self.reaction_threshold), float(self.epsilon), float(self.smooth))

    def validate(self) -> None:
        """Same exceptions, same messages, same order as the reference (src/pde.py:14-17, :199-200)."""
        if self.diffusion_coeff <= 0:
            raise ValueError("diffusion_coeff must be positive")
        if not (0 < self.reaction_threshold < 1):
            raise ValueError("reaction_threshold must be in (0,1)")
        if self.phase_field_weight > 0 and self.epsilon <= 0:
            raise ValueError("epsilon must be positive")


def activation_kind(name: str) -> int:
    try:
        return _ACTIVATION_KIND[name.lower()]
    except KeyError:
        raise ValueError(f"Unsupported output_activation: {name}. Must be 'sigmoid' or 'tanh'") from None


def _x_dtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"predictions must be float32 or bfloat16, got {t.dtype}")


def _t_dtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype in (torch.uint8, torch.bool):
        return U8
    raise TypeError(f"targets must be float32, bfloat16, uint8 or bool, got {t.dtype}")


def _bhw(x: torch.Tensor) -> Tuple[int, int, int]:
    if x.dim() == 4:
        if x.shape[1] != 1:
            # the reference's stencils are (1,1,3,3) conv kernels (src/pde.py:45-47): C != 1 raises there too
            raise RuntimeError(f"expected single-channel maps (B,1,H,W), got {tuple(x.shape)}")
        return x.shape[0], x.shape[2], x.shape[3]
    if x.dim() == 3:
        return x.shape[0], x.shape[1], x.shape[2]
    raise RuntimeError(f"expected (B,1,H,W) or (B,H,W) maps, got {tuple(x.shape)}")


def check_maps(x: torch.Tensor, t: Optional[torch.Tensor]) -> Tuple[int, int, int]:
    if not x.is_cuda:
        raise RuntimeError("physics_informed_image_segmentation_b200 runs on CUDA tensors only (no CPU fallback); "
                           f"got a tensor on {x.device}")
    if not x.is_contiguous():
        # reference: predictions.view(-1) (src/loss.py:130) raises on non-contiguous input
        raise RuntimeError("view size is not compatible with input tensor's size and stride: "
                           "predictions must be contiguous (reference src/loss.py:130 uses .view(-1))")
    B, H, W = _bhw(x)
    if H < 2 or W < 2:
        raise RuntimeError("reflect padding needs H >= 2 and W >= 2 (reference src/pde.py:67)")
    if t is not None:
        if t.shape != x.shape:
            raise ValueError(f"Using a target size ({tuple(t.shape)}) that is different to the input size "
                             f"({tuple(x.shape)}) is deprecated. Please ensure they have the same size.")
        if t.device != x.device:
            raise RuntimeError(f"targets on {t.device}, predictions on {x.device}")
        if not t.is_contiguous():
            raise RuntimeError("targets must be contiguous (reference src/loss.py:131 uses .view(-1))")
    return B, H, W


# ---------------------------------------------------------------------------------------------
# workspace cache: one zero-initialised scratch buffer per (device, stream)
# ---------------------------------------------------------------------------------------------
_WORKSPACES = {}


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def workspace(device: torch.device, B: int, H: int, W: int) -> torch.Tensor:
    need = _lib.lib().pil_workspace_bytes(B, H, W)
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream_ptr(device))
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 16), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


# ---------------------------------------------------------------------------------------------
# raw calls
# ---------------------------------------------------------------------------------------------
def forward_sums(x: torch.Tensor, t: torch.Tensor, p: LossParams, kind: int, sums: Optional[torch.Tensor] = None,
                 report: Optional[torch.Tensor] = None, finalize: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """K1.  Returns (sums float64[8], report float32[8] or None).  With finalize=True the kernel's last
    block also assembles the loss as if this shard were the whole batch (single-GPU fast path)."""
    B, H, W = check_maps(x, t)
    L = _lib.lib()
    dev = x.device
    if sums is None:
        sums = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
    if finalize and report is None:
        report = torch.empty(PIL_NOUT, dtype=torch.float32, device=dev)
    ws = workspace(dev, B, H, W)
    cp = p.c()
    with torch.cuda.device(dev):
        st = L.pil_forward(x.data_ptr(), t.data_ptr(), B, H, W, _x_dtype(x), _t_dtype(t), kind, ctypes.byref(cp),
                           sums.data_ptr(), report.data_ptr() if finalize else None, ws.data_ptr(), ws.numel(),
                           _stream_ptr(dev))
    _lib.check(st, "pil_forward")
    return sums, (report if finalize else None)


def forward_moments(x: torch.Tensor, t: torch.Tensor, kind: int, moments: Optional[torch.Tensor] = None,
                    ex: Optional["_lib.PilExchange"] = None) -> torch.Tensor:
    """One pass over the maps -> the 13 parameter-independent sums (float64[16]) every loss setting is a
    closed form of (include/pil.h pil_forward_moments).  All-reduce (SUM) across ranks when sharded -- or pass `ex`
    (a PeerExchange step descriptor): the kernel's last block then stores the shard's sums into every rank's mailbox
    and sweep_finalize_xchg assembles the global losses without a collective call."""
    B, H, W = check_maps(x, t)
    dev = x.device
    if moments is None:
        moments = torch.empty(PIL_NMOMENTS, dtype=torch.float64, device=dev)
    ws = workspace(dev, B, H, W)
    with torch.cuda.device(dev):
        if ex is not None:
            st = _lib.lib().pil_forward_moments_xchg(x.data_ptr(), t.data_ptr(), B, H, W, _x_dtype(x), _t_dtype(t), kind,
                                                     moments.data_ptr(), ws.data_ptr(), ws.numel(), ctypes.byref(ex), _stream_ptr(dev))
        else:
            st = _lib.lib().pil_forward_moments(x.data_ptr(), t.data_ptr(), B, H, W, _x_dtype(x), _t_dtype(t), kind,
                                                moments.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _lib.check(st, "pil_forward_moments")
    return moments


def sweep_finalize_xchg(ex: "_lib.PilExchange", n_global: int, params: Sequence[LossParams], device: torch.device,
                        reports: Optional[torch.Tensor] = None, moments: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Loss reports (float32[K, 8], K <= 32) of the GLOBAL batch from the moment sums all ranks pushed into the mailbox
    (forward_moments(..., ex=ex) with the same descriptor); identical on every rank."""
    K = len(params)
    if not (1 <= K <= 32):
        raise ValueError("the exchanged sweep evaluates 1..32 settings per call")
    for p in params:
        p.validate()
    dev = torch.device(device)
    if reports is None:
        reports = torch.empty(K, PIL_NOUT, dtype=torch.float32, device=dev)
    arr = (PilParams * K)(*[p.c() for p in params])
    with torch.cuda.device(dev):
        st = _lib.lib().pil_sweep_finalize_xchg(ctypes.byref(ex), int(n_global), arr, K, reports.data_ptr(),
                                                moments.data_ptr() if moments is not None else None, _stream_ptr(dev))
    _lib.check(st, "pil_sweep_finalize_xchg")
    return reports


def sweep_finalize(moments: torch.Tensor, n_global: int, params: Sequence[LossParams],
                   reports: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Loss reports (float32[K, 8], rows laid out like finalize_report) of K parameter settings from one
    moments vector; no pass over the maps."""
    K = len(params)
    if K < 1:
        raise ValueError("need at least one parameter setting")
    for p in params:
        p.validate()
    dev = moments.device
    if reports is None:
        reports = torch.empty(K, PIL_NOUT, dtype=torch.float32, device=dev)
    arr = (PilParams * K)(*[p.c() for p in params])
    with torch.cuda.device(dev):
        st = _lib.lib().pil_sweep_finalize(moments.data_ptr(), int(n_global), arr, K, reports.data_ptr(), _stream_ptr(dev))
    _lib.check(st, "pil_sweep_finalize")
    return reports


def sweep_losses(x: torch.Tensor, t: torch.Tensor, params: Sequence[LossParams], kind: int = X_PROB, group=None) -> torch.Tensor:
    """Batched loss evaluation of a parameter grid (reference run_ablation.py:159-224 S2/S3 grids; BASELINE
    config 4): ONE read of x and t for all K settings.  With `group`, x and t are this rank's shard of the
    batch and the moments are all-reduced, so every rank returns the losses of the global batch."""
    if group is not None:
        import torch.distributed as dist
        from .sharding import peer_exchange_for

        px = peer_exchange_for(group, x.device)
        if px is not None and not px.device_epoch and len(params) <= 32:
            ex = px.next_step()  # the kernels swap the moment sums over peer memory: no collective call
            forward_moments(x.detach(), t.detach(), kind, ex=ex)
            return sweep_finalize_xchg(ex, -1, params, x.device)
        moments = forward_moments(x.detach(), t.detach(), kind)
        dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=group)
        return sweep_finalize(moments, -1, params)
    moments = forward_moments(x.detach(), t.detach(), kind)
    return sweep_finalize(moments, -1, params)


def finalize_report(sums: torch.Tensor, n_global: int, p: LossParams, report: Optional[torch.Tensor] = None) -> torch.Tensor:
    L = _lib.lib()
    dev = sums.device
    if report is None:
        report = torch.empty(PIL_NOUT, dtype=torch.float32, device=dev)
    cp = p.c()
    with torch.cuda.device(dev):
        st = L.pil_finalize(sums.data_ptr(), int(n_global), ctypes.byref(cp), report.data_ptr(), _stream_ptr(dev))
    _lib.check(st, "pil_finalize")
    return report


def backward_grad(x: torch.Tensor, t: torch.Tensor, p: LossParams, kind: int, gsums: torch.Tensor, n_global: int,
                  upstream: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
                  out: Optional[torch.Tensor] = None, only_if_scaled: bool = False) -> torch.Tensor:
    """K2.  grad = upstream * grad_scale * dL/dx, same dtype/shape as x.  only_if_scaled: `out` already holds the
    gradient for upstream == 1; the kernel returns at once in that case (pil_backward_if_scaled)."""
    B, H, W = check_maps(x, t)
    L = _lib.lib()
    dev = x.device
    if out is None:
        out = torch.empty_like(x)
    up_ptr = None
    if upstream is not None:
        if upstream.dtype != torch.float32 or upstream.numel() != 1 or upstream.device != dev:
            upstream = upstream.to(device=dev, dtype=torch.float32).reshape(1)
        up_ptr = upstream.data_ptr()
    cp = p.c()
    fn = L.pil_backward_if_scaled if only_if_scaled else L.pil_backward
    with torch.cuda.device(dev):
        st = fn(x.data_ptr(), t.data_ptr(), out.data_ptr(), B, H, W, _x_dtype(x), _t_dtype(t), kind,
                ctypes.byref(cp), gsums.data_ptr(), int(n_global), up_ptr, float(grad_scale), _stream_ptr(dev))
    _lib.check(st, "pil_backward_if_scaled" if only_if_scaled else "pil_backward")
    return out


def forward_pointwise(x: torch.Tensor, t: torch.Tensor, p: LossParams, kind: int,
                      sums: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K1L: the neighbour-free sums only (I, P, T, BCE, double well); flat streaming kernel."""
    B, H, W = check_maps(x, t)
    dev = x.device
    if sums is None:
        sums = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
    ws = workspace(dev, B, H, W)
    cp = p.c()
    with torch.cuda.device(dev):
        st = _lib.lib().pil_forward_pointwise(x.data_ptr(), t.data_ptr(), B, H, W, _x_dtype(x), _t_dtype(t), kind,
                                              ctypes.byref(cp), sums.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _lib.check(st, "pil_forward_pointwise")
    return sums


def backward_accumulate(x: torch.Tensor, t: torch.Tensor, p: LossParams, kind: int, gsums: torch.Tensor, n_global: int,
                        upstream: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
                        out: Optional[torch.Tensor] = None, stencil_sums: Optional[torch.Tensor] = None,
                        report: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """K2 + the stencil sums the pointwise forward skipped.  Returns (grad, stencil_sums); when
    `report` is given the kernel's last block also assembles the loss (single shard)."""
    B, H, W = check_maps(x, t)
    dev = x.device
    if out is None:
        out = torch.empty_like(x)
    if stencil_sums is None:
        stencil_sums = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
    up_ptr = None
    if upstream is not None:
        if upstream.dtype != torch.float32 or upstream.numel() != 1 or upstream.device != dev:
            upstream = upstream.to(device=dev, dtype=torch.float32).reshape(1)
        up_ptr = upstream.data_ptr()
    ws = workspace(dev, B, H, W)
    cp = p.c()
    with torch.cuda.device(dev):
        st = _lib.lib().pil_backward_accumulate(x.data_ptr(), t.data_ptr(), out.data_ptr(), B, H, W, _x_dtype(x), _t_dtype(t),
                                                kind, ctypes.byref(cp), gsums.data_ptr(), int(n_global), up_ptr,
                                                float(grad_scale), stencil_sums.data_ptr(),
                                                report.data_ptr() if report is not None else None,
                                                ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _lib.check(st, "pil_backward_accumulate")
    return out, stencil_sums


# most recent per-image threshold counts, keyed on the exact tensors they were computed from
_LAST_COUNTS = None


def remember_counts(x: torch.Tensor, t: torch.Tensor, kind: int, threshold: float, counts: torch.Tensor) -> None:
    import weakref

    global _LAST_COUNTS
    _LAST_COUNTS = (weakref.ref(x), x._version, weakref.ref(t), t._version, int(kind), float(threshold), counts)


def cached_counts(x: torch.Tensor, t: torch.Tensor, kind: int, threshold: float) -> Optional[torch.Tensor]:
    """The counts of the last fused evaluation if it saw exactly these tensors (same objects, same version
    counters -- a tensor that merely reuses the address never hits), same input kind and threshold."""
    if _LAST_COUNTS is None:
        return None
    wx, vx, wt, vt, k, thr, counts = _LAST_COUNTS
    if wx() is x and x._version == vx and wt() is t and t._version == vt and k == int(kind) and thr == float(threshold):
        return counts
    return None


def forward_pointwise_metrics(x: torch.Tensor, t: torch.Tensor, p: LossParams, kind: int, threshold: float = 0.5,
                              sums: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None,
                              ex: Optional["_lib.PilExchange"] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """K1M: the pointwise forward plus, on the same read of x and t, three counts per image
    (float64[B, 4]: sum [u>thr]*t, sum [u>thr], sum t, 0) -- what the reference's per-step Dice/IoU metrics
    need (src/metrics.py:38-73, src/evaluate.py:62-97).  Returns (sums, counts)."""
    B, H, W = check_maps(x, t)
    dev = x.device
    if sums is None:
        sums = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
    if counts is None:
        counts = torch.empty(B, 4, dtype=torch.float64, device=dev)
    ws = workspace(dev, B, H, W)
    cp = p.c()
    with torch.cuda.device(dev):
        st = _lib.lib().pil_forward_pointwise_metrics(x.data_ptr(), t.data_ptr(), B, H, W, _x_dtype(x), _t_dtype(t), kind,
                                                      ctypes.byref(cp), sums.data_ptr(), counts.data_ptr(), float(threshold),
                                                      ws.data_ptr(), ws.numel(), ctypes.byref(ex) if ex is not None else None,
                                                      _stream_ptr(dev))
    _lib.check(st, "pil_forward_pointwise_metrics")
    return sums, counts


def image_metrics(counts: torch.Tensor, smooth: float = 1e-6) -> Tuple[torch.Tensor, torch.Tensor]:
    """(dice[B], iou[B]) float32 from the per-image counts, on the device (src/metrics.py:66-70,
    src/evaluate.py:90-94)."""
    B = counts.shape[0]
    dev = counts.device
    dice = torch.empty(B, dtype=torch.float32, device=dev)
    iou = torch.empty(B, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().pil_image_metrics(counts.data_ptr(), B, float(smooth), dice.data_ptr(), iou.data_ptr(), _stream_ptr(dev))
    _lib.check(st, "pil_image_metrics")
    return dice, iou


_BOUNDARY_WS = {}


def boundary_counts(x: torch.Tensor, t: torch.Tensor, kind: int = X_PROB, threshold: float = 0.5, tolerance: int = 2,
                    counts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-image boundary counts (int64[B, 4]: |Bp|, |Bt|, #Bp within tolerance of Bt, #Bt within tolerance of Bp) of
    the thresholded prediction against the mask -- the integers behind the reference's OpenCV boundary-F1
    (src/evaluate.py:102-193), counted on the device without a host sync (include/pil.h pil_boundary_counts)."""
    if not x.is_cuda:
        raise RuntimeError("CUDA tensors only (no CPU fallback)")
    if t.shape != x.shape or t.device != x.device or not x.is_contiguous() or not t.is_contiguous():
        raise RuntimeError("predictions and targets must be contiguous tensors of one shape on one device")
    if not (0 <= int(tolerance) <= 6):
        raise NotImplementedError("boundary tolerance must be an integer in [0, 6]")
    B, H, W = _bhw(x)
    dev = x.device
    if counts is None:
        counts = torch.empty(B, 4, dtype=torch.int64, device=dev)
    L = _lib.lib()
    need = L.pil_boundary_workspace_bytes(B, H, W)
    key = (dev.index, _stream_ptr(dev))
    ws = _BOUNDARY_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _BOUNDARY_WS[key] = ws
    with torch.cuda.device(dev):
        st = L.pil_boundary_counts(x.data_ptr(), t.data_ptr(), B, H, W, _x_dtype(x), _t_dtype(t), kind, float(threshold),
                                   int(tolerance), counts.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _lib.check(st, "pil_boundary_counts")
    return counts


def boundary_f1(counts: torch.Tensor, tolerance: int = 2, smooth: float = 1e-6) -> torch.Tensor:
    """float32[B] boundary-F1 from the counts, with the reference's float32 arithmetic (src/evaluate.py:171-191)."""
    B = counts.shape[0]
    dev = counts.device
    out = torch.empty(B, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().pil_boundary_f1(counts.data_ptr(), B, int(tolerance), float(smooth), out.data_ptr(), _stream_ptr(dev))
    _lib.check(st, "pil_boundary_f1")
    return out


def forward_pointwise_xchg(x: torch.Tensor, t: torch.Tensor, p: LossParams, kind: int, ex: "_lib.PilExchange",
                           sums: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K1L + push of the shard's sums into every rank's mailbox (include/pil.h pil_forward_pointwise_xchg)."""
    B, H, W = check_maps(x, t)
    dev = x.device
    if sums is None:
        sums = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
    ws = workspace(dev, B, H, W)
    cp = p.c()
    with torch.cuda.device(dev):
        st = _lib.lib().pil_forward_pointwise_xchg(x.data_ptr(), t.data_ptr(), B, H, W, _x_dtype(x), _t_dtype(t), kind,
                                                   ctypes.byref(cp), sums.data_ptr(), ws.data_ptr(), ws.numel(),
                                                   ctypes.byref(ex), _stream_ptr(dev))
    _lib.check(st, "pil_forward_pointwise_xchg")
    return sums


def backward_accumulate_xchg(x: torch.Tensor, t: torch.Tensor, p: LossParams, kind: int, ex: "_lib.PilExchange",
                             n_global: int = -1, upstream: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
                             out: Optional[torch.Tensor] = None, stencil_sums: Optional[torch.Tensor] = None,
                             report: Optional[torch.Tensor] = None, total_sums: Optional[torch.Tensor] = None):
    """K2 fed from the mailbox; its last block swaps the stencil sums and finalises the GLOBAL loss.
    Returns (grad of this shard, report float32[8] of the global batch, total_sums float64[8])."""
    B, H, W = check_maps(x, t)
    dev = x.device
    if out is None:
        out = torch.empty_like(x)
    if stencil_sums is None:
        stencil_sums = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
    if report is None:
        report = torch.empty(PIL_NOUT, dtype=torch.float32, device=dev)
    if total_sums is None:
        total_sums = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
    up_ptr = None
    if upstream is not None:
        if upstream.dtype != torch.float32 or upstream.numel() != 1 or upstream.device != dev:
            upstream = upstream.to(device=dev, dtype=torch.float32).reshape(1)
        up_ptr = upstream.data_ptr()
    ws = workspace(dev, B, H, W)
    cp = p.c()
    with torch.cuda.device(dev):
        st = _lib.lib().pil_backward_accumulate_xchg(x.data_ptr(), t.data_ptr(), out.data_ptr(), B, H, W, _x_dtype(x),
                                                     _t_dtype(t), kind, ctypes.byref(cp), ctypes.byref(ex), int(n_global),
                                                     up_ptr, float(grad_scale), stencil_sums.data_ptr(), report.data_ptr(),
                                                     total_sums.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _lib.check(st, "pil_backward_accumulate_xchg")
    return out, report, total_sums


def exchange_finalize(ex: "_lib.PilExchange", n_global: int, p: LossParams, device: torch.device,
                      report: Optional[torch.Tensor] = None, total_sums: Optional[torch.Tensor] = None):
    """Deferred finalisation (PIL_XCHG_DEFER_FINALIZE): global report + complete global sums from the mailbox."""
    dev = torch.device(device)
    if report is None:
        report = torch.empty(PIL_NOUT, dtype=torch.float32, device=dev)
    if total_sums is None:
        total_sums = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
    cp = p.c()
    with torch.cuda.device(dev):
        st = _lib.lib().pil_exchange_finalize(ctypes.byref(ex), int(n_global), ctypes.byref(cp), report.data_ptr(),
                                              total_sums.data_ptr(), _stream_ptr(dev))
    _lib.check(st, "pil_exchange_finalize")
    return report, total_sums


def loss_fwd_bwd(x: torch.Tensor, t: torch.Tensor, p: LossParams, kind: int, grad: Optional[torch.Tensor] = None,
                 sums: Optional[torch.Tensor] = None, report: Optional[torch.Tensor] = None):
    """One training-step evaluation on a single shard: pointwise forward -> backward (+ stencil sums).
    Returns (report float32[8], sums float64[8] -- the same vector pil_forward produces, grad like x)."""
    B, H, W = check_maps(x, t)
    dev = x.device
    if grad is None:
        grad = torch.empty_like(x)
    if sums is None:
        sums = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
    if report is None:
        report = torch.empty(PIL_NOUT, dtype=torch.float32, device=dev)
    ws = workspace(dev, B, H, W)
    cp = p.c()
    with torch.cuda.device(dev):
        st = _lib.lib().pil_loss_fwd_bwd(x.data_ptr(), t.data_ptr(), grad.data_ptr(), B, H, W, _x_dtype(x), _t_dtype(t), kind,
                                         ctypes.byref(cp), sums.data_ptr(), report.data_ptr(), ws.data_ptr(), ws.numel(),
                                         _stream_ptr(dev))
    _lib.check(st, "pil_loss_fwd_bwd")
    return report, sums, grad


class StepGraph:
    """One training-step evaluation (pointwise forward -> backward with stencil sums and loss report) as ONE
    CUDA-graph launch (include/pil.h pil_step_graph_*).  The graph is bound to `x`, `t` and `grad`: refill x and t in
    place (copy_) between launches.  With `exchange` (a sharding.PeerExchange created with device_epoch=True) x and t
    are this rank's shard and `report` / `sums` describe the GLOBAL batch; construction is then collective and
    performs one real step.  `launch()` enqueues the step on the current stream and returns `report` (float32[8],
    device, no sync)."""

    def __init__(self, x: torch.Tensor, t: torch.Tensor, p: LossParams, kind: int, grad: Optional[torch.Tensor] = None,
                 exchange=None, n_global: int = -1, grad_scale: float = 1.0, upstream: Optional[torch.Tensor] = None):
        p.validate()
        B, H, W = check_maps(x, t)
        dev = x.device
        self.x, self.t = x, t
        self.grad = torch.empty_like(x) if grad is None else grad
        self.sums = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)      # single shard: complete; sharded: this shard's pointwise sums
        self.total = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev) if exchange is not None else self.sums
        self.stencil = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
        self.report = torch.empty(PIL_NOUT, dtype=torch.float32, device=dev)
        self.upstream = None
        if upstream is not None:
            self.upstream = upstream.to(device=dev, dtype=torch.float32).reshape(1)
        # a private workspace: the graph keeps its address
        self._ws = torch.zeros(max(_lib.lib().pil_workspace_bytes(B, H, W), 1 << 12), dtype=torch.uint8, device=dev)
        self._ex = None
        if exchange is not None:
            if not getattr(exchange, "device_epoch", False):
                raise ValueError("StepGraph needs a PeerExchange created with device_epoch=True")
            self._ex = exchange.next_step()
            self._exchange = exchange  # keeps the mailboxes mapped
        single = exchange is None and upstream is None and grad_scale == 1.0
        cp = p.c()
        self._h = ctypes.c_void_p()
        with torch.cuda.device(dev):
            st = _lib.lib().pil_step_graph_create(
                ctypes.byref(self._h), x.data_ptr(), t.data_ptr(), self.grad.data_ptr(), B, H, W, _x_dtype(x), _t_dtype(t), kind,
                ctypes.byref(cp), self.sums.data_ptr(), self.report.data_ptr(), self.stencil.data_ptr(), self._ws.data_ptr(),
                self._ws.numel(), ctypes.byref(self._ex) if self._ex is not None else None,
                int(B * H * W if exchange is None else n_global), self.upstream.data_ptr() if self.upstream is not None else None,
                float(grad_scale), self.total.data_ptr() if (exchange is not None or single) else None, _stream_ptr(dev))
        _lib.check(st, "pil_step_graph_create")
        self._dev = dev

    def launch(self) -> torch.Tensor:
        with torch.cuda.device(self._dev):
            st = _lib.lib().pil_step_graph_launch(self._h, _stream_ptr(self._dev))
        _lib.check(st, "pil_step_graph_launch")
        return self.report

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().pil_step_graph_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SweepGraph:
    """One sweep step (forward_moments -> sweep_finalize[_xchg]: K loss reports from one pass over the maps) as ONE
    CUDA-graph launch (include/pil.h pil_sweep_graph_create).  Bound to `x` and `t`: refill them in place between
    launches.  With `exchange` (a sharding.PeerExchange created with device_epoch=True) x and t are this rank's shard
    and the reports describe the GLOBAL batch; construction is then collective and performs one real sweep step.
    `launch()` returns `reports` (float32[K, 8], device, no sync)."""

    def __init__(self, x: torch.Tensor, t: torch.Tensor, kind: int, params: Sequence[LossParams], exchange=None, n_global: int = -1):
        K = len(params)
        if K < 1 or (exchange is not None and K > 32):
            raise ValueError("need 1..32 parameter settings (any number >= 1 without an exchange)")
        for p in params:
            p.validate()
        B, H, W = check_maps(x, t)
        dev = x.device
        self.x, self.t = x, t
        self.moments = torch.empty(PIL_NMOMENTS, dtype=torch.float64, device=dev)          # this shard's
        self.global_moments = torch.empty(PIL_NMOMENTS, dtype=torch.float64, device=dev) if exchange is not None else self.moments
        self.reports = torch.empty(K, PIL_NOUT, dtype=torch.float32, device=dev)
        self._ws = torch.zeros(max(_lib.lib().pil_workspace_bytes(B, H, W), 1 << 12), dtype=torch.uint8, device=dev)
        self._ex = None
        if exchange is not None:
            if not getattr(exchange, "device_epoch", False):
                raise ValueError("SweepGraph needs a PeerExchange created with device_epoch=True")
            self._ex = exchange.next_step()
            self._exchange = exchange  # keeps the mailboxes mapped
        arr = (PilParams * K)(*[p.c() for p in params])
        self._h = ctypes.c_void_p()
        with torch.cuda.device(dev):
            st = _lib.lib().pil_sweep_graph_create(
                ctypes.byref(self._h), x.data_ptr(), t.data_ptr(), B, H, W, _x_dtype(x), _t_dtype(t), kind, self.moments.data_ptr(),
                self._ws.data_ptr(), self._ws.numel(), ctypes.byref(self._ex) if self._ex is not None else None,
                int(B * H * W if exchange is None else n_global), arr, K, self.reports.data_ptr(),
                self.global_moments.data_ptr() if exchange is not None else None, _stream_ptr(dev))
        _lib.check(st, "pil_sweep_graph_create")
        self._dev = dev

    def launch(self) -> torch.Tensor:
        with torch.cuda.device(self._dev):
            st = _lib.lib().pil_step_graph_launch(self._h, _stream_ptr(self._dev))
        _lib.check(st, "pil_step_graph_launch")
        return self.reports

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().pil_step_graph_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def scale_gradient(grad: torch.Tensor, upstream: torch.Tensor) -> torch.Tensor:
    """grad *= upstream on the device, in place; free when upstream == 1 (plain loss.backward())."""
    dev = grad.device
    if upstream.dtype != torch.float32 or upstream.numel() != 1 or upstream.device != dev:
        upstream = upstream.to(device=dev, dtype=torch.float32).reshape(1)
    with torch.cuda.device(dev):
        st = _lib.lib().pil_scale_gradient(grad.data_ptr(), _x_dtype(grad), grad.numel(), upstream.data_ptr(), _stream_ptr(dev))
    _lib.check(st, "pil_scale_gradient")
    return grad


def launch_info() -> _lib.PilLaunchInfo:
    info = _lib.PilLaunchInfo()
    _lib.check(_lib.lib().pil_last_launch_info(ctypes.byref(info)), "pil_last_launch_info")
    return info


# ---------------------------------------------------------------------------------------------
# autograd glue
# ---------------------------------------------------------------------------------------------
def _params_for(p: LossParams, which: int) -> LossParams:
    """Weights that make the fused gradient the gradient of ONE report entry."""
    if which == OUT_DICE:
        return replace(p, dice_weight=1.0, bce_weight=0.0, pde_weight=0.0, phase_field_weight=0.0)
    if which == OUT_BCE:
        return replace(p, dice_weight=0.0, bce_weight=1.0, pde_weight=0.0, phase_field_weight=0.0)
    if which == OUT_RD:
        return replace(p, dice_weight=0.0, bce_weight=0.0, pde_weight=1.0, phase_field_weight=0.0)
    if which == OUT_PF:
        return replace(p, dice_weight=0.0, bce_weight=0.0, pde_weight=0.0, phase_field_weight=1.0)
    return p


class _FusedLossFn(torch.autograd.Function):
    """Training path.  Both halves of `loss = criterion(x, t); loss.backward()` are known to run, so (eager=True) the
    forward launches the pointwise sums kernel AND the backward kernel (which also accumulates the two stencil sums):
    every 5-point stencil is evaluated once per step, and the loss report is complete when forward returns.
    backward() then only applies the upstream scalar (a no-op kernel when it is 1).  The price is a full-size gradient
    buffer parked on the graph whether or not .backward() is ever called; eager=False (component evaluations such as
    criterion.bce(...) or pde_regularization.compute_loss(...), and criterion.lazy_backward = True) runs the full
    forward kernel only and computes the gradient on demand from the saved sums.

    x is saved through save_for_backward WITHOUT detaching, so an in-place change of x between forward and backward
    raises autograd's usual version-counter error instead of returning a stale gradient.

    Returns the requested report entry plus the whole report vector (float32[8]), so one evaluation serves the total
    loss AND the four logged components (reference src/train.py:120-150)."""

    @staticmethod
    def forward(ctx, x, t, p: LossParams, kind: int, which: int, group, ddp_average: bool, metrics_thr=None, eager: bool = True):
        x_d = x.detach()
        t_d = t.detach()
        pg = _params_for(p, which)  # the gradient is that of the requested entry
        counts = None
        dev = x_d.device
        scale, n_global = 1.0, x_d.numel()
        grad = None
        if group is not None:
            import torch.distributed as dist

            scale = float(dist.get_world_size(group)) if ddp_average else 1.0
            n_global = -1
        if not eager:
            # forward only: all six sums from the full forward kernel, no gradient buffer
            if metrics_thr is not None:
                _, counts = forward_pointwise_metrics(x_d, t_d, pg, kind, metrics_thr)
            sums, _ = forward_sums(x_d, t_d, pg, kind, finalize=False)
            if group is not None:
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
            report = finalize_report(sums, n_global, p)
        elif group is not None:
            from .sharding import peer_exchange_for

            px = peer_exchange_for(group, dev)
            if px is not None and pg is p:
                # peer-memory path: the two kernels swap their sums over NVLink themselves (no NCCL call)
                ex = px.next_step()
                if metrics_thr is not None:
                    _, counts = forward_pointwise_metrics(x_d, t_d, pg, kind, metrics_thr, ex=ex)
                else:
                    forward_pointwise_xchg(x_d, t_d, pg, kind, ex)
                grad, report, sums = backward_accumulate_xchg(x_d, t_d, pg, kind, ex, -1, grad_scale=scale)
            else:
                if metrics_thr is not None:
                    sums, counts = forward_pointwise_metrics(x_d, t_d, pg, kind, metrics_thr)
                else:
                    sums = forward_pointwise(x_d, t_d, pg, kind)
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)           # the gradient needs global I, P, T
                grad, stencil = backward_accumulate(x_d, t_d, pg, kind, sums, -1, grad_scale=scale)
                dist.all_reduce(stencil, op=dist.ReduceOp.SUM, group=group)        # only the loss VALUE needs these
                sums = sums + stencil
                report = finalize_report(sums, -1, p)
        elif metrics_thr is not None:
            # per-image threshold counts ride on the pointwise forward (K1M), then the usual backward
            sums, counts = forward_pointwise_metrics(x_d, t_d, pg, kind, metrics_thr)
            report = torch.empty(PIL_NOUT, dtype=torch.float32, device=dev)
            stencil = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
            grad, _ = backward_accumulate(x_d, t_d, pg, kind, sums, x_d.numel(), stencil_sums=stencil, report=report)
            sums = sums + stencil
            if pg is not p:
                report = finalize_report(sums, x_d.numel(), p)
        else:
            report, sums, grad = loss_fwd_bwd(x_d, t_d, pg, kind)
            if pg is not p:
                report = finalize_report(sums, x_d.numel(), p)
        ctx.save_for_backward(x, t, sums)  # x itself, not a detached alias: keeps autograd's in-place check alive
        ctx.grad = grad
        ctx.p, ctx.kind, ctx.n_global, ctx.scale = pg, kind, n_global, scale
        if counts is None:
            counts = torch.empty(0, 4, dtype=torch.float64, device=dev)
        ctx.mark_non_differentiable(report, counts)
        return report[which], report, counts

    @staticmethod
    def backward(ctx, g_loss, _g_report, _g_counts):
        x, t, sums = ctx.saved_tensors  # raises if x or t was modified in place since forward
        grad = ctx.grad
        none8 = (None,) * 8
        if grad is not None:
            ctx.grad = None  # the buffer is handed out (and possibly scaled in place) once
            if grad.dtype == torch.bfloat16:
                # the eager buffer is already rounded to bf16: rescaling it would round twice.  The kernel decides on
                # the device: upstream == 1 (plain loss.backward()) -> nothing to do; otherwise recomputed in fp32.
                return (backward_grad(x.detach(), t.detach(), ctx.p, ctx.kind, sums, ctx.n_global, upstream=g_loss,
                                      grad_scale=ctx.scale, out=grad, only_if_scaled=True),) + none8
            return (scale_gradient(grad, g_loss),) + none8
        # lazy evaluation, or a second backward through a retained graph: from the saved global sums
        grad = backward_grad(x.detach(), t.detach(), ctx.p, ctx.kind, sums, ctx.n_global, upstream=g_loss, grad_scale=ctx.scale)
        return (grad,) + none8


_TAIL_WS = {}


def _tail_workspace(dev: torch.device, C: int) -> torch.Tensor:
    key = (dev.index, _stream_ptr(dev), C)
    ws = _TAIL_WS.get(key)
    if ws is None:
        ws = torch.zeros(_lib.lib().pil_tail_workspace_bytes(C), dtype=torch.uint8, device=dev)
        _TAIL_WS[key] = ws
    return ws


def _check_tail(feat: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], t: torch.Tensor):
    if not feat.is_cuda:
        raise RuntimeError("CUDA tensors only (no CPU fallback)")
    if feat.dim() != 4 or not feat.is_contiguous():
        raise RuntimeError("features must be a contiguous NCHW tensor (B, C, H, W); call .contiguous() on channels_last maps")
    B, C, H, W = feat.shape
    if not (1 <= C <= 128):
        raise RuntimeError(f"the fused model tail supports 1..128 feature channels, got {C}")
    if weight.numel() != C or weight.dtype != torch.float32 or not weight.is_contiguous() or weight.device != feat.device:
        raise RuntimeError(f"weight must hold the {C} float32 values of the (1, {C}, 1, 1) output-convolution kernel on {feat.device}")
    if bias is not None and (bias.numel() != 1 or bias.dtype != torch.float32 or bias.device != feat.device):
        raise RuntimeError("bias must be one float32 value on the features' device")
    if tuple(t.shape) not in ((B, 1, H, W), (B, H, W)) or t.device != feat.device or not t.is_contiguous():
        raise ValueError(f"targets must be contiguous (B,1,H,W) maps on {feat.device}, got {tuple(t.shape)}")
    if H < 2 or W < 2:
        raise RuntimeError("reflect padding needs H >= 2 and W >= 2 (reference src/pde.py:67)")
    return B, C, H, W


class _TailLossFn(torch.autograd.Function):
    """1x1 output convolution + activation + loss as ONE differentiable op over (features, weight, bias):
    pil_tail_forward (convolution fused with the pointwise sums) -> the fused backward kernel on the logits ->
    pil_tail_backward in backward() (dL/dfeatures, dL/dweight, dL/dbias from one pass over the features)."""

    @staticmethod
    def forward(ctx, feat, weight, bias, t, p: LossParams, kind: int, group, ddp_average: bool):
        fd, td = feat.detach(), t.detach()
        B, C, H, W = _check_tail(fd, weight.detach(), bias.detach() if bias is not None else None, td)
        dev = fd.device
        L = _lib.lib()
        logits = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
        sums = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
        ws = _tail_workspace(dev, C)
        cp = p.c()
        ex = None
        scale, n_global = 1.0, B * H * W
        if group is not None:
            import torch.distributed as dist
            from .sharding import peer_exchange_for

            scale = float(dist.get_world_size(group)) if ddp_average else 1.0
            n_global = -1
            px = peer_exchange_for(group, dev)
            ex = px.next_step() if px is not None else None
        wd, bd = weight.detach().reshape(-1), (bias.detach().reshape(-1) if bias is not None else None)
        with torch.cuda.device(dev):
            st = L.pil_tail_forward(fd.data_ptr(), _x_dtype(fd), wd.data_ptr(), bd.data_ptr() if bd is not None else None,
                                    td.data_ptr(), _t_dtype(td), B, C, H, W, kind, ctypes.byref(cp), logits.data_ptr(),
                                    sums.data_ptr(), ws.data_ptr(), ws.numel(), ctypes.byref(ex) if ex is not None else None,
                                    _stream_ptr(dev))
        _lib.check(st, "pil_tail_forward")
        if ex is not None:
            gl, report, sums = backward_accumulate_xchg(logits, td, p, kind, ex, -1, grad_scale=scale)
        elif group is not None:
            import torch.distributed as dist

            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
            gl, stencil = backward_accumulate(logits, td, p, kind, sums, -1, grad_scale=scale)
            dist.all_reduce(stencil, op=dist.ReduceOp.SUM, group=group)
            sums = sums + stencil
            report = finalize_report(sums, -1, p)
        else:
            report = torch.empty(PIL_NOUT, dtype=torch.float32, device=dev)
            stencil = torch.empty(PIL_NSUMS, dtype=torch.float64, device=dev)
            gl, _ = backward_accumulate(logits, td, p, kind, sums, n_global, stencil_sums=stencil, report=report)
        ctx.save_for_backward(feat, weight)
        ctx.gl, ctx.has_bias, ctx.wshape = gl, bias is not None, weight.shape
        ctx.mark_non_differentiable(report, logits)
        return report[OUT_TOTAL], report, logits

    @staticmethod
    def backward(ctx, g_loss, _g_report, _g_logits):
        feat, weight = ctx.saved_tensors
        gl = ctx.gl
        if gl is None:
            raise RuntimeError("the fused model tail keeps its logits gradient for one backward pass only")
        ctx.gl = None
        gl = scale_gradient(gl, g_loss)
        fd = feat.detach()
        B, C, H, W = fd.shape
        dev = fd.device
        dfeat = torch.empty_like(fd)
        dw = torch.empty(C, dtype=torch.float32, device=dev)
        db = torch.empty(1, dtype=torch.float32, device=dev) if ctx.has_bias else None
        ws = _tail_workspace(dev, C)
        with torch.cuda.device(dev):
            st = _lib.lib().pil_tail_backward(fd.data_ptr(), _x_dtype(fd), weight.detach().reshape(-1).data_ptr(), gl.data_ptr(),
                                              B, C, H, W, dfeat.data_ptr(), dw.data_ptr(), db.data_ptr() if db is not None else None,
                                              ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _lib.check(st, "pil_tail_backward")
        return dfeat, dw.view(ctx.wshape), db, None, None, None, None, None


def fused_tail_loss(feat: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], t: torch.Tensor, p: LossParams,
                    kind: int = X_LOGITS_SIGMOID, group=None, ddp_average: bool = True):
    """(loss with grad_fn over feat / weight / bias, detached report float32[8], detached fp32 logits (B,1,H,W))."""
    p.validate()
    if kind not in (X_LOGITS_SIGMOID, X_LOGITS_TANH):
        raise ValueError("the fused model tail produces logits: activation must be 'sigmoid' or 'tanh'")
    return _TailLossFn.apply(feat, weight, bias, t, p, kind, group, ddp_average)


def fused_loss_with_counts(x: torch.Tensor, t: torch.Tensor, p: LossParams, kind: int = X_PROB, which: int = OUT_TOTAL,
                           group=None, ddp_average: bool = True, metrics_threshold: Optional[float] = None, eager: bool = True):
    """(requested scalar with grad_fn, detached report float32[8], per-image counts float64[B,4] or None).
    With metrics_threshold the per-image threshold counts of this rank's images come from the same pass
    over the maps in the training path; the no-grad path spends one extra pointwise pass on them."""
    p.validate()
    check_maps(x, t)
    if x.requires_grad and torch.is_grad_enabled():
        loss, report, counts = _FusedLossFn.apply(x, t, p, kind, which, group, ddp_average, metrics_threshold, eager)
        if metrics_threshold is None:
            return loss, report, None
        remember_counts(x, t, kind, metrics_threshold, counts)
        return loss, report, counts
    loss, report = fused_loss(x, t, p, kind, which, group, ddp_average)
    counts = None
    if metrics_threshold is not None:
        _, counts = forward_pointwise_metrics(x.detach(), t.detach(), p, kind, metrics_threshold)
        remember_counts(x, t, kind, metrics_threshold, counts)
    return loss, report, counts


def fused_loss(x: torch.Tensor, t: torch.Tensor, p: LossParams, kind: int = X_PROB, which: int = OUT_TOTAL,
               group=None, ddp_average: bool = True, eager: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """(requested scalar with grad_fn, detached report float32[8]).  eager=False: no gradient buffer is computed unless
    .backward() is actually called (component / logging evaluations)."""
    p.validate()
    check_maps(x, t)
    if x.requires_grad and torch.is_grad_enabled():
        loss, report, _ = _FusedLossFn.apply(x, t, p, kind, which, group, ddp_average, None, eager)
        return loss, report
    # no graph needed (validation / logging): skip the autograd.Function overhead
    if group is not None:
        import torch.distributed as dist

        sums, _ = forward_sums(x.detach(), t.detach(), p, kind, finalize=False)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        report = finalize_report(sums, -1, p)
    else:
        _, report = forward_sums(x.detach(), t.detach(), p, kind, finalize=True)
    return report[which], report


# ---------------------------------------------------------------------------------------------
# stand-alone stencil operators (PDERegularization methods)
# ---------------------------------------------------------------------------------------------
def _stencil_call(name: str, *tensors: torch.Tensor) -> torch.Tensor:
    u = tensors[0]
    B, H, W = check_maps(u, None)
    if u.dtype != torch.float32:
        raise TypeError(f"{name} works on float32 maps, got {u.dtype}")
    out = torch.empty_like(u)
    fn = getattr(_lib.lib(), name)
    with torch.cuda.device(u.device):
        st = fn(*[t.data_ptr() for t in tensors], out.data_ptr(), B, H, W, _stream_ptr(u.device))
    _lib.check(st, name)
    return out


class _LaplacianFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u):
        return _stencil_call("pil_laplacian", u.detach())

    @staticmethod
    def backward(ctx, g):
        return _stencil_call("pil_laplacian_adjoint", g.contiguous())


class _GradMagSqFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u):
        ctx.save_for_backward(u.detach())
        return _stencil_call("pil_grad_mag_sq", u.detach())

    @staticmethod
    def backward(ctx, g):
        (u,) = ctx.saved_tensors
        return _stencil_call("pil_grad_mag_sq_backward", u, g.contiguous())


class _ReactionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, a: float):
        ud = u.detach()
        if not ud.is_cuda:
            raise RuntimeError("CUDA tensors only (no CPU fallback)")
        if ud.dtype != torch.float32 or not ud.is_contiguous():
            raise TypeError("reaction_term works on contiguous float32 maps")
        out = torch.empty_like(ud)
        with torch.cuda.device(ud.device):
            st = _lib.lib().pil_reaction(ud.data_ptr(), out.data_ptr(), ud.numel(), float(a), _stream_ptr(ud.device))
        _lib.check(st, "pil_reaction")
        ctx.save_for_backward(ud)
        ctx.a = float(a)
        return out

    @staticmethod
    def backward(ctx, g):
        (u,) = ctx.saved_tensors
        a = ctx.a
        return g * (u * (2.0 * (1.0 + a) - 3.0 * u) - a), None  # f'(u) = -3u^2 + 2(1+a)u - a


def laplacian(u: torch.Tensor) -> torch.Tensor:
    return _LaplacianFn.apply(u)


def grad_mag_sq(u: torch.Tensor) -> torch.Tensor:
    return _GradMagSqFn.apply(u)


def reaction(u: torch.Tensor, a: float) -> torch.Tensor:
    return _ReactionFn.apply(u, a)
