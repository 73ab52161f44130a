"""GPU parity of the boundary-F1 kernels (pil_boundary_counts / pil_boundary_f1, through the C ABI) with the golden
vectors of the real reference (OpenCV on the host, tests/golden/ref_bf1.npz) and with the CPU oracle on fresh inputs:
bit-exact on the four integer counts per image, float32-exact on the F1."""
import os

import numpy as np
import pytest
import torch

from oracle import boundary_oracle as bo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def env():
    import physics_informed_image_segmentation_b200 as P
    from physics_informed_image_segmentation_b200 import functional as Fn

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return P, Fn, torch.device("cuda:0"), np.load(os.path.join(ROOT, "tests", "golden", "ref_bf1.npz"))


def oracle_counts(p, t, thr, tol):
    return np.array([bo.boundary_counts((p[i, 0] > np.float32(thr)).astype(np.float32), t[i, 0], tol) for i in range(p.shape[0])], dtype=np.int64)


@pytest.mark.parametrize("tol", [0, 1, 2, 3])
def test_golden_reference_vectors(env, tol):
    P, Fn, dev, gold = env
    p, t = torch.from_numpy(gold["predictions"]).to(dev), torch.from_numpy(gold["targets"]).to(dev)
    counts = Fn.boundary_counts(p, t, Fn.X_PROB, 0.5, tol)
    assert np.array_equal(counts.cpu().numpy(), oracle_counts(gold["predictions"], gold["targets"], 0.5, tol))
    f1 = P.compute_boundary_f1_batch(p, t, threshold=0.5, tolerance=tol)
    assert f1.dtype == torch.float32 and f1.shape == (p.shape[0],)
    assert np.array_equal(f1.cpu().numpy(), gold[f"f1_tol{tol}"]), (f1.cpu().numpy(), gold[f"f1_tol{tol}"])
    if tol == 2:
        assert P.compute_boundary_f1(p, t).item() == gold["f1_first"][0]
        assert np.array_equal(P.compute_boundary_f1_batch(p, t, threshold=0.3).cpu().numpy(), gold["f1_thr03"])


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 1, 9), (2, 7, 1), (3, 2, 2), (2, 33, 47), (2, 128, 128), (1, 257, 300)])
def test_random_masks_against_oracle(env, shape):
    P, Fn, dev, _ = env
    B, H, W = shape
    rng = np.random.default_rng(B * 1000 + H * 10 + W)
    g = torch.Generator().manual_seed(H + W)
    lo = torch.randn(B, 1, max(H // 12, 2), max(W // 12, 2), generator=g)
    sm = torch.nn.functional.interpolate(lo, size=(H, W), mode="bicubic", align_corners=False)
    t = (sm > 0.1).float()
    p = torch.sigmoid(2.5 * sm + 0.7 * torch.randn(B, 1, H, W, generator=g))
    if H * W > 1:
        t[0, 0].view(-1)[rng.integers(0, H * W, size=max(1, H * W // 50))] = 1.0  # speckles: 1-pixel objects and holes
    for tol in (0, 2):
        c = Fn.boundary_counts(p.to(dev), t.to(dev), Fn.X_PROB, 0.5, tol).cpu().numpy()
        assert np.array_equal(c, oracle_counts(p.numpy(), t.numpy(), 0.5, tol)), (shape, tol)
    f1 = P.compute_boundary_f1_batch(p.to(dev), t.to(dev)).cpu().numpy()
    assert np.array_equal(f1, bo.boundary_f1_batch(p.numpy(), t.numpy()))


def test_logits_entry_and_other_dtypes(env):
    P, Fn, dev, gold = env
    g = torch.Generator().manual_seed(9)
    z = 3.0 * torch.randn(4, 1, 48, 80, generator=g)
    z = torch.nn.functional.avg_pool2d(z, 5, 1, 2) * 3
    t = (torch.rand(4, 1, 48, 80, generator=g) > 0.5).float()
    t = (torch.nn.functional.avg_pool2d(t, 7, 1, 3) > 0.5).float()
    want = oracle_counts(torch.sigmoid(z).numpy(), t.numpy(), 0.5, 2)
    assert np.array_equal(Fn.boundary_counts(z.to(dev), t.to(dev), Fn.X_LOGITS_SIGMOID, 0.5, 2).cpu().numpy(), want)
    assert np.array_equal(Fn.boundary_counts(z.to(dev), t.to(dev).to(torch.uint8), Fn.X_LOGITS_SIGMOID, 0.5, 2).cpu().numpy(), want)
    want_t = oracle_counts(((torch.tanh(z) + 1) / 2).numpy(), t.numpy(), 0.5, 2)
    assert np.array_equal(Fn.boundary_counts(z.to(dev), t.to(dev), Fn.X_LOGITS_TANH, 0.5, 2).cpu().numpy(), want_t)
    zb = z.bfloat16()
    want_b = oracle_counts(torch.sigmoid(zb.float()).numpy(), t.numpy(), 0.5, 2)
    assert np.array_equal(Fn.boundary_counts(zb.to(dev), t.bfloat16().to(dev), Fn.X_LOGITS_SIGMOID, 0.5, 2).cpu().numpy(), want_b)
    with pytest.raises(NotImplementedError):
        Fn.boundary_counts(z.to(dev), t.to(dev), Fn.X_LOGITS_SIGMOID, 0.5, 7)
