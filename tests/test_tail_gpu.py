"""The fused model tail (1x1 output convolution + activation + loss; pil_tail_forward / pil_tail_backward through the
module API) against an fp64 PyTorch evaluation of the same graph on the CPU -- conv2d, the activation, the torch
restatement of the reference loss (oracle/torch_port.py, pinned bit-identical to the real reference) and autograd:
loss, dL/dfeatures, dL/dweight, dL/dbias within 1e-5 (fp32 features) / 1e-2 (bf16 features)."""
import numpy as np
import pytest
import torch

from tests.helpers import rel_max, rel_scalar

pytestmark = pytest.mark.gpu


def _reference(feat, w, b, t, params, kind):
    from oracle import torch_port

    f = feat.double().requires_grad_(True)
    ww = w.double().requires_grad_(True)
    bb = b.double().requires_grad_(True) if b is not None else None
    z = torch.nn.functional.conv2d(f, ww.view(1, -1, 1, 1), bb)
    loss = torch_port.loss(torch_port.activate(z, kind), t.double(), params)
    loss.backward()
    return loss.item(), f.grad, ww.grad, (bb.grad if bb is not None else None), z.detach()


@pytest.mark.parametrize("shape,act", [((2, 64, 64, 96), "sigmoid"), ((3, 16, 37, 53), "sigmoid"), ((2, 64, 32, 32), "tanh"), ((1, 5, 2, 2), "sigmoid")])
@pytest.mark.parametrize("use_bias", [True, False])
def test_tail_matches_fp64_autograd(shape, act, use_bias):
    import physics_informed_image_segmentation_b200 as P
    from oracle import pil_oracle as po

    dev = torch.device("cuda:0")
    B, C, H, W = shape
    g = torch.Generator().manual_seed(C * 7 + H)
    feat = torch.randn(B, C, H, W, generator=g)
    w = 0.1 * torch.randn(C, generator=g)   # logits of std < 1: no fp32 saturation, where the fp64 evaluation would differ by design
    b = torch.tensor([0.1]) if use_bias else None
    lo = torch.randn(B, 1, max(H // 8, 1), max(W // 8, 1), generator=g)
    t = (torch.nn.functional.interpolate(lo, size=(H, W), mode="nearest") > 0).float()
    kw = dict(dice_weight=0.5, bce_weight=0.5, pde_weight=1e-2, phase_field_weight=1e-2, diffusion_coeff=2.0, reaction_threshold=0.4, epsilon=0.1)
    kind = 1 if act == "sigmoid" else 2
    l_ref, df_ref, dw_ref, db_ref, z_ref = _reference(feat, w, b, t, po.Params(**kw), kind)

    crit = P.DiceBCEPDELoss(**kw).to(dev)
    fd = feat.to(dev).requires_grad_(True)
    wd = w.to(dev).view(1, C, 1, 1).clone().requires_grad_(True)   # the shape nn.Conv2d keeps its kernel in
    bd = b.to(dev).requires_grad_(True) if use_bias else None
    loss = crit.forward_features(fd, wd, bd, t.to(dev), activation=act)
    (2.0 * loss).backward()   # a non-unit upstream gradient
    assert rel_scalar(loss.item(), l_ref) < 1e-5
    assert rel_max(crit.last_logits.cpu().numpy(), z_ref.numpy()) < 1e-5
    assert rel_max(fd.grad.cpu().numpy(), 2.0 * df_ref.numpy()) < 1e-5
    assert wd.grad.shape == wd.shape and rel_max(wd.grad.cpu().numpy().reshape(-1), 2.0 * dw_ref.numpy()) < 1e-5
    if use_bias:
        assert rel_max(bd.grad.cpu().numpy(), 2.0 * db_ref.numpy()) < 1e-5
    comps = crit.components()
    assert rel_scalar(comps["loss"].item(), l_ref) < 1e-5


def test_tail_bf16_features_and_same_result_as_unfused_modules():
    import physics_informed_image_segmentation_b200 as P
    from oracle import pil_oracle as po

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(2, 64, 48, 64, generator=g)
    conv = torch.nn.Conv2d(64, 1, 1).to(dev)
    t = (torch.rand(2, 1, 48, 64, generator=g) > 0.5).float().to(dev)
    kw = dict(pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0)
    crit = P.DiceBCEPDELoss(**kw).to(dev)
    # unfused: torch's convolution + the fused loss on its logits
    f1 = feat.to(dev).requires_grad_(True)
    l1 = crit.forward_logits(conv(f1), t)
    l1.backward()
    gw1, gb1 = conv.weight.grad.clone(), conv.bias.grad.clone()
    conv.zero_grad()
    f2 = feat.to(dev).requires_grad_(True)
    l2 = crit.forward_features(f2, conv.weight, conv.bias, t)
    l2.backward()
    assert rel_scalar(l2.item(), l1.item()) < 2e-6
    assert rel_max(f2.grad.cpu().numpy(), f1.grad.cpu().numpy()) < 1e-5
    assert rel_max(conv.weight.grad.cpu().numpy(), gw1.cpu().numpy()) < 1e-4   # torch's fp32 weight-gradient reduction order differs
    assert rel_max(conv.bias.grad.cpu().numpy(), gb1.cpu().numpy()) < 1e-4
    # bf16 features (what autocast hands over): against the fp64 evaluation of the up-cast values, 1e-2
    fb = feat.bfloat16()
    l_ref, df_ref, dw_ref, db_ref, _ = _reference(fb.float(), conv.weight.detach().cpu().reshape(-1), conv.bias.detach().cpu(), t.cpu(),
                                                  po.Params(dice_weight=0.5, bce_weight=0.5, **kw), 1)
    conv.zero_grad()
    f3 = fb.to(dev).requires_grad_(True)
    l3 = crit.forward_features(f3, conv.weight, conv.bias, t)
    l3.backward()
    assert f3.grad.dtype == torch.bfloat16
    assert rel_scalar(l3.item(), l_ref) < 1e-5
    assert rel_max(f3.grad.float().cpu().numpy(), df_ref.numpy()) < 1e-2
    assert rel_max(conv.weight.grad.cpu().numpy().reshape(-1), dw_ref.numpy()) < 1e-4
    with pytest.raises(RuntimeError):
        crit.forward_features(feat.to(dev).to(memory_format=torch.channels_last), conv.weight, conv.bias, t)
