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


class PeerExchange:
    """Peer-memory mailboxes of one data-parallel group (include/pil.h PilExchange).

    One process per GPU of one NVLink/NVSwitch node.  Construction allocates this rank's mailbox,
    all-gathers the CUDA IPC handles through torch.distributed (plumbing) and maps every peer's
    mailbox.  `next_step()` returns the descriptor for the next training step (epoch + 1): all
    ranks must call it in lock step, which a data-parallel loop does by construction.
    """

    def __init__(self, device: torch.device, group=None, device_epoch: bool = False):
        """device_epoch: the step tag lives in a counter in each rank's own mailbox (PIL_XCHG_DEVICE_EPOCH) instead of
        coming from the host with every call: the arguments of a step are then identical from step to step, which is
        what lets functional.StepGraph replay a captured step.  The mode is fixed for the lifetime of the mailboxes
        and must be the same on every rank."""
        import ctypes

        import torch.distributed as dist

        from . import _lib

        L = _lib.lib()
        self._lib = L
        self.device = torch.device(device)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > _lib.PIL_MAX_RANKS:
            raise ValueError(f"peer exchange supports up to {_lib.PIL_MAX_RANKS} ranks, got {self.world}")
        own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * _lib.PIL_IPC_HANDLE_BYTES)()
        with torch.cuda.device(self.device):
            _lib.check(L.pil_exchange_alloc(ctypes.byref(own), handle), "pil_exchange_alloc")
        self._own = own.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self._peers = {}
        self._ex = _lib.PilExchange()
        self._ex.rank, self._ex.world, self._ex.epoch = self.rank, self.world, 0
        for r in range(self.world):
            if r == self.rank:
                self._ex.mailbox[r] = self._own
                continue
            ptr = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * _lib.PIL_IPC_HANDLE_BYTES).from_buffer_copy(handles[r])
            with torch.cuda.device(self.device):
                _lib.check(L.pil_exchange_open(buf, ctypes.byref(ptr)), "pil_exchange_open")
            self._peers[r] = ptr.value
            self._ex.mailbox[r] = ptr.value
        self._epoch = 0
        self.device_epoch = bool(device_epoch)
        if self.device_epoch:
            self._ex.flags = _lib.PIL_XCHG_DEVICE_EPOCH
        dist.barrier(group=group)  # nobody pushes before every mailbox is mapped everywhere

    def next_step(self):
        """Descriptor of the next step; the caller runs forward_pointwise_xchg then backward_accumulate_xchg with it."""
        from . import _lib

        ex = _lib.PilExchange()
        ex.rank, ex.world, ex.epoch = self.rank, self.world, self._epoch
        ex.flags = self._ex.flags
        for r in range(self.world):
            ex.mailbox[r] = self._ex.mailbox[r]
        self._epoch += 1  # device-epoch mode: informational only, the kernels count the steps themselves
        return ex

    def check_lockstep(self) -> None:
        """Collective, one small all-reduce: every rank must have taken the same number of exchange steps.  The step tag
        comes from this host-side counter (host-epoch mode), so one extra or missing grad-mode forward on one rank
        desynchronises every later step -- the kernels then time out (NaN loss, zero gradient, `timed_out()`).  Call it
        where a host sync happens anyway; training.train_epoch does at the end of every epoch."""
        import torch.distributed as dist

        e = torch.tensor([self._epoch, -self._epoch], dtype=torch.int64, device=self.device)
        dist.all_reduce(e, op=dist.ReduceOp.MAX, group=self.group)
        lo, hi = -int(e[1].item()), int(e[0].item())
        if lo != hi:
            raise RuntimeError(f"peer exchange out of lock step: ranks have taken between {lo} and {hi} exchange steps "
                               f"(this rank: {self._epoch}); every rank must evaluate the loss the same number of times")

    def timed_out(self) -> bool:
        """Host sync: has any exchange wait on this rank given up (peer dead / not in lock step)?"""
        import ctypes

        st = ctypes.c_int(0)
        from . import _lib

        with torch.cuda.device(self.device):
            _lib.check(self._lib.pil_exchange_status(self._own, ctypes.byref(st),
                                                     torch.cuda.current_stream(self.device).cuda_stream), "pil_exchange_status")
        return st.value != 0

    def close(self):
        import torch.distributed as dist

        if self._own is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier(group=self.group)  # no peer may still be writing into a mailbox that goes away
            for ptr in self._peers.values():
                self._lib.pil_exchange_close(ptr)
            self._lib.pil_exchange_free(self._own)
        self._peers, self._own = {}, None


_EXCHANGES = {}


def _group_key(group, device) -> tuple:
    import torch.distributed as dist

    g = None if (group is None or group is dist.group.WORLD) else group
    return (id(g) if g is not None else 0, torch.device(device).index)


def enable_peer_exchange(device: torch.device, group=None, device_epoch: bool = False) -> PeerExchange:
    """Collective: create (once) the peer-memory exchange the fused loss uses for `group`."""
    key = _group_key(group, device)
    if key not in _EXCHANGES:
        _EXCHANGES[key] = PeerExchange(device, group, device_epoch=device_epoch)
    elif _EXCHANGES[key].device_epoch != bool(device_epoch):
        raise RuntimeError("the peer exchange of this group already exists with a different epoch mode")
    return _EXCHANGES[key]


def peer_exchange_for(group, device: torch.device) -> Optional[PeerExchange]:
    return _EXCHANGES.get(_group_key(group, device))


def disable_peer_exchange() -> None:
    for px in list(_EXCHANGES.values()):
        px.close()
    _EXCHANGES.clear()


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
    if s[6] > 0:  # probabilities outside [0,1]: the reference's nn.BCELoss raises; the device finalize returns NaN
        total = float("nan")
    return {"loss": total, "dice_loss": dice_loss, "bce_loss": bce, "pde_loss": rd, "phase_field_loss": pf,
            "n_invalid": s[6], "n_pixels": n}
