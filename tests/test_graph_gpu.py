"""The captured-step path (pil_step_graph_*): one CUDA-graph launch must produce exactly what the two direct calls do,
also after the bound input buffers were refilled in place, for fp32 / bf16 maps and the aligned / scalar kernels."""
import numpy as np
import pytest
import torch

from tests.helpers import blob_inputs, iid_inputs, rel_max, rel_scalar

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200 import functional as Fn
    from oracle import pil_oracle as po

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return P, Fn, po, torch.device("cuda:0")


@pytest.mark.parametrize("shape", [(8, 128, 128), (8, 256, 256), (3, 37, 53), (2, 64, 1000)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_step_graph_matches_direct_calls(env, shape, dtype):
    P, Fn, po, dev = env
    p = P.LossParams(dice_weight=0.5, bce_weight=0.5, pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0,
                     reaction_threshold=0.5, epsilon=0.05)
    z, t = blob_inputs(*shape, seed=3)
    x, tt = z.to(dev).to(dtype), t.to(dev).to(dtype)
    g = Fn.StepGraph(x, tt, p, Fn.X_LOGITS_SIGMOID)
    for step in range(3):
        if step:  # refill the bound buffers in place
            z2, t2 = iid_inputs(*shape, seed=10 + step)
            x.copy_(z2.to(dev).to(dtype))
            tt.copy_(t2.to(dev).to(dtype))
        rep = g.launch().clone()
        grad = g.grad.clone()
        sums = g.sums.clone()
        rep_d, sums_d, grad_d = Fn.loss_fwd_bwd(x, tt, p, Fn.X_LOGITS_SIGMOID)
        torch.cuda.synchronize()
        assert torch.equal(rep[:6], rep_d[:6]), (step, rep, rep_d)
        assert torch.equal(grad, grad_d)
        assert torch.equal(sums, sums_d)
    if dtype == torch.float32:  # and against the fp64 oracle on the last inputs
        comps, og = po.loss_and_grad(x.cpu().numpy().astype(np.float64), tt.cpu().numpy().astype(np.float64), po.STAGE2, po.X_LOGITS_SIGMOID)
        assert rel_scalar(rep[0].item(), comps[0]) < 1e-5
        assert rel_max(grad.cpu().numpy(), og) < 1e-5
    g.close()


def test_step_graph_with_upstream_and_scale(env):
    P, Fn, po, dev = env
    p = P.LossParams(pde_weight=1e-3, phase_field_weight=1e-3, diffusion_coeff=2.0)
    z, t = blob_inputs(4, 96, 160, seed=5)
    x, tt = z.to(dev), t.to(dev)
    up = torch.tensor([0.25], device=dev)
    g = Fn.StepGraph(x, tt, p, Fn.X_LOGITS_SIGMOID, upstream=up, grad_scale=4.0)
    g.launch()
    _, _, grad_d = Fn.loss_fwd_bwd(x, tt, p, Fn.X_LOGITS_SIGMOID)
    torch.cuda.synchronize()
    assert rel_max(g.grad.cpu().numpy(), grad_d.cpu().numpy()) < 1e-6  # 0.25 * 4 == 1
    up.fill_(0.5)  # read on the device at run time: no re-capture
    g.launch()
    torch.cuda.synchronize()
    assert rel_max(g.grad.cpu().numpy(), 2.0 * grad_d.cpu().numpy()) < 1e-6
    g.close()


@pytest.mark.parametrize("shape", [(8, 128, 128), (3, 37, 53), (4, 96, 512)])
def test_sweep_graph_matches_direct_calls(env, shape):
    """pil_sweep_graph_create: moments pass + K loss reports as one graph launch == the two direct calls, also after the
    bound maps were refilled in place; and == the fused loss of each setting evaluated on its own."""
    P, Fn, po, dev = env
    grid = P.s2_grid() + P.s3_grid()
    z, t = blob_inputs(*shape, seed=21)
    x, tt = z.to(dev), t.to(dev)
    g = Fn.SweepGraph(x, tt, Fn.X_LOGITS_SIGMOID, grid)
    for step in range(3):
        if step:
            z2, t2 = iid_inputs(*shape, seed=30 + step)
            x.copy_(z2.to(dev))
            tt.copy_(t2.to(dev))
        rep = g.launch().clone()
        rep_d = Fn.sweep_finalize(Fn.forward_moments(x, tt, Fn.X_LOGITS_SIGMOID), -1, grid)
        torch.cuda.synchronize()
        assert torch.equal(rep[:, :6], rep_d[:, :6]), (step, rep[:, 0], rep_d[:, 0])
    for k in (0, 5, 8):  # against the one-setting-at-a-time evaluation (another kernel, other summation order)
        s1, r1 = Fn.forward_sums(x, tt, grid[k], Fn.X_LOGITS_SIGMOID)
        for c in range(4):  # total, dice, bce, rd (the sweep reports pf only where its weight is > 0)
            assert rel_scalar(rep[k, c].item(), r1[c].item()) < 1e-5, (k, c)
    g.close()
