"""Peer-memory exchange (include/pil.h PilExchange) with several ranks EMULATED on one GPU.

Every emulated rank owns a mailbox on the same device and "maps" the others by plain pointer.  A
kernel must never wait for a flag that a LATER launch writes (one GPU gives no co-scheduling
guarantee), so per step all pointwise forwards are enqueued first (they push phase 0), then all
backwards with PIL_XCHG_DEFER_FINALIZE (they find every phase-0 flag already set and only push phase
1), then the deferred finalizes.  The result must equal the unsharded single-shard evaluation and
the CPU oracle: global Dice sums and means over the global pixel count (reference
src/loss.py:134-141, src/pde.py:143,:210)."""
import ctypes

import numpy as np
import pytest
import torch

from tests.helpers import blob_inputs, iid_inputs, rel_l2, rel_max, rel_scalar

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


class EmulatedRanks:
    def __init__(self, world):
        from physics_informed_image_segmentation_b200 import _lib

        self._lib = _lib
        self.L = _lib.lib()
        self.world = world
        self.boxes = []
        for _ in range(world):
            p = ctypes.c_void_p()
            _lib.check(self.L.pil_exchange_alloc(ctypes.byref(p), None), "pil_exchange_alloc")
            self.boxes.append(p.value)
        self.epoch = 0

    def descriptors(self, flags):
        out = []
        for r in range(self.world):
            ex = self._lib.PilExchange()
            ex.rank, ex.world, ex.epoch, ex.flags = r, self.world, self.epoch, flags
            for q in range(self.world):
                ex.mailbox[q] = self.boxes[q]
            out.append(ex)
        self.epoch += 1
        return out

    def status(self, r):
        st = ctypes.c_int(-1)
        self._lib.check(self.L.pil_exchange_status(self.boxes[r], ctypes.byref(st), None), "pil_exchange_status")
        return st.value

    def close(self):
        torch.cuda.synchronize()
        for b in self.boxes:
            self.L.pil_exchange_free(b)


@pytest.mark.parametrize("world,B,H,W,maker", [(2, 4, 64, 128, iid_inputs), (3, 7, 33, 52, blob_inputs), (8, 8, 16, 24, iid_inputs)])
def test_emulated_ranks_match_unsharded_and_oracle(dev, world, B, H, W, maker):
    from oracle import pil_oracle as po
    from physics_informed_image_segmentation_b200 import _lib, functional as Fn
    from physics_informed_image_segmentation_b200.sharding import shard_bounds

    z, t = maker(B, H, W, seed=77)
    p = Fn.LossParams(dice_weight=0.5, bce_weight=0.5, pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0,
                      reaction_threshold=0.5, epsilon=0.05)
    kind = Fn.X_LOGITS_SIGMOID
    x, tt = z.to(dev), t.to(dev)
    rep1, sums1, g1 = Fn.loss_fwd_bwd(x, tt, p, kind)
    torch.cuda.synchronize()

    ranks = EmulatedRanks(world)
    try:
        shards = [shard_bounds(B, r, world) for r in range(world)]
        xs = [x[b0:b1].contiguous() for b0, b1 in shards]
        ts = [tt[b0:b1].contiguous() for b0, b1 in shards]
        for step in range(3):  # three steps: both slot parities and one slot reuse
            exs = ranks.descriptors(_lib.PIL_XCHG_DEFER_FINALIZE)
            for r in range(world):
                Fn.forward_pointwise_xchg(xs[r], ts[r], p, kind, exs[r])
            grads = []
            for r in range(world):
                g, _, _ = Fn.backward_accumulate_xchg(xs[r], ts[r], p, kind, exs[r], n_global=-1)
                grads.append(g)
            reports = [Fn.exchange_finalize(exs[r], -1, p, dev) for r in range(world)]
            torch.cuda.synchronize()
            for r in range(world):
                assert ranks.status(r) == 0, "an exchange wait timed out"
            g = torch.cat(grads, dim=0)
            # every rank assembles the same global report, bit for bit (rank-ordered sums)
            for r in range(1, world):
                assert torch.equal(reports[r][0], reports[0][0])
                assert torch.equal(reports[r][1], reports[0][1])
            rep, tot = reports[0]
            for k in range(5):
                assert rel_scalar(rep[k].item(), rep1[k].item()) < 2e-6, (k, rep[k].item(), rep1[k].item())
            assert rel_max(g.cpu().numpy(), g1.cpu().numpy()) < 2e-6
            assert tot[7].item() == B * H * W

        z64, t64 = z.numpy().astype(np.float64), t.numpy().astype(np.float64)
        po_p = po.Params(dice_weight=0.5, bce_weight=0.5, pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0,
                         reaction_threshold=0.5, epsilon=0.05, smooth=1e-6)
        comps, og = po.loss_and_grad(z64, t64, po_p, po.X_LOGITS_SIGMOID)
        assert rel_scalar(rep[0].item(), comps[0]) < 1e-5
        assert rel_max(g.cpu().numpy(), og) < 1e-5 and rel_l2(g.cpu().numpy(), og) < 1e-5
    finally:
        ranks.close()


def test_world_one_in_kernel_finalize(dev):
    """world == 1 through the exchange path: the backward's last block waits on its own flags (already
    set by its own pushes) and finalises in the kernel -- the code path real multi-GPU runs take."""
    from physics_informed_image_segmentation_b200 import functional as Fn

    z, t = iid_inputs(3, 40, 72, seed=5)
    p = Fn.LossParams(pde_weight=1e-3, phase_field_weight=1e-3, diffusion_coeff=2.0)
    x, tt = z.to(dev), t.to(dev)
    rep1, sums1, g1 = Fn.loss_fwd_bwd(x, tt, p, Fn.X_LOGITS_SIGMOID)
    ranks = EmulatedRanks(1)
    try:
        for _ in range(3):
            (ex,) = ranks.descriptors(0)
            Fn.forward_pointwise_xchg(x, tt, p, Fn.X_LOGITS_SIGMOID, ex)
            g, rep, tot = Fn.backward_accumulate_xchg(x, tt, p, Fn.X_LOGITS_SIGMOID, ex, n_global=-1)
            torch.cuda.synchronize()
            assert ranks.status(0) == 0
            assert torch.equal(rep[:5], rep1[:5])
            assert torch.equal(g, g1)
            assert torch.equal(tot, sums1)
    finally:
        ranks.close()


def test_bad_descriptor_is_rejected(dev):
    from physics_informed_image_segmentation_b200 import _lib, functional as Fn

    z, t = iid_inputs(1, 8, 8)
    ex = _lib.PilExchange()
    ex.rank, ex.world, ex.epoch = 0, 2, 0  # mailbox pointers left NULL
    with pytest.raises(_lib.PilError):
        Fn.forward_pointwise_xchg(z.to(dev), t.to(dev), Fn.LossParams(), Fn.X_LOGITS_SIGMOID, ex)
