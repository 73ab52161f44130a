"""world_size-2 test of the data-parallel exchange on CPU (gloo): each rank evaluates the raw sums of
its batch shard (with the CPU oracle, since there is no GPU here), the package's all-reduce adds
them, and the assembled loss / the per-shard gradient with global sums equal the unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import physics_informed_image_segmentation_b200 as P
        from oracle import pil_oracle as po
        from tests.helpers import iid_inputs, rel_max

        z, t = iid_inputs(5, 20, 28, seed=42)  # same global batch on every rank, uneven shards (3 + 2)
        z, t = z.numpy().astype(np.float64), t.numpy().astype(np.float64)
        a, b = P.shard_bounds(5, rank, world)
        local = torch.from_numpy(po.sums(z[a:b], t[a:b], po.STAGE2, 1))
        gs = P.all_reduce_sums(local.clone(), group=None)
        params = P.LossParams(pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0, epsilon=0.05)
        rep = P.loss_report_from_sums(gs.tolist(), None, params)
        full = po.sums(z, t, po.STAGE2, 1)
        comps = po.finalize(full, int(full[7]), po.STAGE2)
        g_full = po.backward(z, t, po.STAGE2, full, int(full[7]), 1)
        g_loc = po.backward(z[a:b], t[a:b], po.STAGE2, gs.numpy(), int(gs[7].item()), 1)
        q.put((rank, abs(rep["loss"] - comps[0]) / comps[0], rel_max(g_loc, g_full[a:b]), float(gs[7]), b - a))
    finally:
        dist.destroy_process_group()


def test_two_rank_sum_exchange_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[4] for r in res] == [3, 2]
    for _, loss_err, grad_err, npx, _ in res:
        assert loss_err < 1e-13 and grad_err < 1e-12 and npx == 5 * 20 * 28


def test_all_reduce_sums_is_identity_without_process_group():
    import physics_informed_image_segmentation_b200 as P

    s = torch.arange(8, dtype=torch.float64)
    assert torch.equal(P.all_reduce_sums(s.clone()), s)
    with pytest.raises(ValueError):
        P.all_reduce_sums(torch.zeros(8))
