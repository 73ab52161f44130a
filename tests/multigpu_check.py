"""Data-parallel check on REAL GPUs (one process per GPU); run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py

Every rank holds a shard of one global batch (generated identically on every rank), evaluates the Stage II
loss through the module API -- first with the NCCL all-reduce of the sums, then with the peer-memory
exchange (sharding.enable_peer_exchange) -- and compares loss, components and its slice of the gradient
with (a) the single-shard evaluation of the whole batch on its own GPU and (b) the CPU oracle.
tests/test_multigpu_spawn.py launches this when at least two GPUs are visible."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200 import functional as Fn, sharding
    from oracle import pil_oracle as po
    from tests.helpers import blob_inputs, rel_max, rel_scalar

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    B, H, W = 2 * world + 1, 96, 160  # unequal shards on purpose
    z, t = blob_inputs(B, H, W, seed=5)
    b0, b1 = sharding.shard_bounds(B, rank, world)
    kw = dict(dice_weight=0.5, bce_weight=0.5, pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0,
              reaction_threshold=0.5, epsilon=0.05)

    # (a) whole batch as a single shard on this GPU
    crit1 = P.DiceBCEPDELoss(**kw).to(dev)
    x_all = z.to(dev).requires_grad_(True)
    l_all = crit1.forward_logits(x_all, t.to(dev))
    l_all.backward()
    rep_all = crit1.last_report.clone()
    # (b) oracle
    comps, og = po.loss_and_grad(z.numpy().astype(np.float64), t.numpy().astype(np.float64), po.STAGE2, po.X_LOGITS_SIGMOID)

    results = {}
    for mode in ("nccl", "peer"):
        if mode == "peer":
            sharding.enable_peer_exchange(dev, None)
        crit = P.DiceBCEPDELoss(**kw, process_group=dist.group.WORLD, ddp_average=False).to(dev)
        crit.enable_batch_metrics(0.5)
        for step in range(3):  # several steps: both mailbox parities, slot reuse
            x = z[b0:b1].to(dev).requires_grad_(True)
            loss = crit.forward_logits(x, t[b0:b1].to(dev))
            loss.backward()
        torch.cuda.synchronize()
        rep = crit.last_report
        for k in range(5):
            assert rel_scalar(rep[k].item(), rep_all[k].item()) < 2e-6, (mode, k, rep[k].item(), rep_all[k].item())
            assert rel_scalar(rep[k].item(), comps[k]) < 1e-5, (mode, k)
        assert rel_max(x.grad.cpu().numpy(), x_all.grad[b0:b1].cpu().numpy()) < 2e-6, mode
        assert rel_max(x.grad.cpu().numpy(), og[b0:b1]) < 1e-5, mode
        m = crit.last_batch_metrics()
        assert m["dice"].shape == (b1 - b0,)
        # every rank holds the same global report
        gathered = [torch.empty_like(rep) for _ in range(world)]
        dist.all_gather(gathered, rep)
        for g in gathered:
            assert torch.equal(g[:5], rep[:5]), mode
        results[mode] = rep[:5].clone()
    assert torch.equal(results["nccl"], results["peer"]), "peer exchange and all-reduce must agree bit for bit"
    # parameter sweep over the sharded batch (BASELINE config 4): moments all-reduced, K losses of the GLOBAL batch
    grid = P.s2_grid() + P.s3_grid()
    sw = P.sweep_losses(z[b0:b1].to(dev), t[b0:b1].to(dev), grid, activation="sigmoid", process_group=dist.group.WORLD)
    sw_all = P.sweep_losses(z.to(dev), t.to(dev), grid, activation="sigmoid")
    assert torch.allclose(sw[:, :5], sw_all[:, :5], rtol=2e-6, atol=0)
    for k, gp in enumerate(grid):
        pp = po.Params(dice_weight=gp.dice_weight, bce_weight=gp.bce_weight, pde_weight=gp.pde_weight,
                       phase_field_weight=gp.phase_field_weight, diffusion_coeff=gp.diffusion_coeff,
                       reaction_threshold=gp.reaction_threshold, epsilon=gp.epsilon, smooth=gp.smooth)
        s = po.sums(z.numpy().astype(np.float64), t.numpy().astype(np.float64), pp, po.X_LOGITS_SIGMOID)
        assert rel_scalar(sw[k, 0].item(), po.finalize(s, int(s[7]), pp)[0]) < 1e-5, k
    px = sharding.peer_exchange_for(None, dev)
    assert px is not None and not px.timed_out()

    # ---- DDP: the averaged PARAMETER gradient of a conv model under ddp_average=True equals the single-process gradient
    # on the concatenated batch, through the fused train step and the peer exchange
    from tests.step_model import TinySegNet, make_batches
    from torch.nn.parallel import DistributedDataParallel as DDP

    Bg = 4 * world
    (images, masks), = make_batches(1, Bg, 64, 96, seed=21)
    g0, g1 = sharding.shard_bounds(Bg, rank, world)
    class F64Net(TinySegNet):
        """fp64 convolutions, fp32 logits: the parameter gradients then differ between the two runs only by what the
        LOSS gradient differs (the fp32 conv backward's own summation noise would dominate the comparison otherwise)"""

        def forward(self, x):
            return super().forward(x.double()).float()

    torch.manual_seed(7)
    ref_model = F64Net(4, "sigmoid").double().to(dev)
    ddp_model = DDP(F64Net(4, "sigmoid").double().to(dev), device_ids=[local])
    ddp_model.module.load_state_dict(ref_model.state_dict())
    crit_ref = P.DiceBCEPDELoss(**kw).to(dev)
    crit_ddp = P.DiceBCEPDELoss(**kw, process_group=dist.group.WORLD, ddp_average=True).to(dev)
    opt_ref = torch.optim.SGD(ref_model.parameters(), lr=0.0)
    opt_ddp = torch.optim.SGD(ddp_model.parameters(), lr=0.0)
    r_ref = P.train_epoch(ref_model, [(images, masks)], crit_ref, opt_ref, dev, return_components=True, compute_metrics=True)
    r_ddp = P.train_epoch(ddp_model, [(images[g0:g1], masks[g0:g1])], crit_ddp, opt_ddp, dev, return_components=True, compute_metrics=True)
    for (n1, p1), (n2, p2) in zip(ref_model.named_parameters(), ddp_model.module.named_parameters()):
        assert rel_max(p2.grad.cpu().numpy(), p1.grad.cpu().numpy()) < 1e-5, ("ddp parameter gradient", n1)
    for k in r_ref:
        assert rel_scalar(r_ddp[k], r_ref[k]) < (1e-5 if not k.endswith("_score") else 1e-6), ("ddp epoch result", k, r_ddp[k], r_ref[k])
    sharding.disable_peer_exchange()
    dist.barrier()

    # ---- one CUDA-graph launch per step over the device-epoch exchange == the direct exchange path (which ran the
    # metrics variant of the pointwise kernel: another summation order, so equal to fp32 rounding, not bit for bit)
    pxd = sharding.enable_peer_exchange(dev, None, device_epoch=True)
    xs, ts = z[b0:b1].to(dev).contiguous(), t[b0:b1].to(dev).contiguous()
    lp = P.LossParams(**kw)
    graph = Fn.StepGraph(xs, ts, lp, Fn.X_LOGITS_SIGMOID, exchange=pxd, n_global=B * H * W, grad_scale=1.0)
    for step in range(4):
        rep_g = graph.launch().clone()
    torch.cuda.synchronize()
    for k in range(5):
        assert rel_scalar(rep_g[k].item(), results["peer"][k].item()) < 2e-6, ("graph", k, rep_g, results["peer"])
        assert rel_scalar(rep_g[k].item(), comps[k]) < 1e-5, ("graph vs oracle", k)
    gathered = [torch.empty_like(rep_g) for _ in range(world)]
    dist.all_gather(gathered, rep_g)
    for gr in gathered:
        assert torch.equal(gr[:5], rep_g[:5]), "every rank holds the same global report"
    assert rel_max(graph.grad.cpu().numpy(), og[b0:b1]) < 1e-5
    # the sharded parameter sweep as one graph launch per step over the same device-epoch exchange (training steps and
    # sweep steps may share a mailbox: each completes the epoch it used)
    sgw = Fn.SweepGraph(xs, ts, Fn.X_LOGITS_SIGMOID, grid, exchange=pxd, n_global=B * H * W)
    for step in range(3):
        rep_sw = sgw.launch().clone()
    graph.launch()  # and a training step after it still sees consistent epochs
    torch.cuda.synchronize()
    assert torch.allclose(rep_sw[:, :5], sw_all[:, :5], rtol=2e-6, atol=0), ("sweep graph", rep_sw[:, 0], sw_all[:, 0])
    gathered = [torch.empty_like(rep_sw) for _ in range(world)]
    dist.all_gather(gathered, rep_sw)
    for gr in gathered:
        assert torch.equal(gr[:, :5], rep_sw[:, :5]), "every rank holds the same sweep reports"
    sgw.close()
    # the host-buffer session over the same exchange: global report, gradient of the shard on the device
    with P.HostSession(b1 - b0, H, W, device=local) as sess:
        for step in range(2):
            rep_h = sess.run_sharded(z[b0:b1].numpy(), t[b0:b1].numpy(), lp, pxd, n_global=B * H * W)
        for k in range(5):
            assert rel_scalar(float(rep_h[k]), comps[k]) < 1e-5, ("session", k, rep_h)
        assert rel_max(sess.device_gradient().cpu().numpy(), og[b0:b1]) < 1e-5
    assert not pxd.timed_out()
    graph.close()
    sharding.disable_peer_exchange()
    dist.barrier()
    if rank == 0:
        print(f"multigpu_check ok: world {world}, loss {results['peer'][0].item():.7f} (oracle {comps[0]:.7f})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
