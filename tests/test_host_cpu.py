"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/pil.h declares, the
host-side mirror of the reference interface (constructors, attributes, errors), argument validation
of the C ABI that needs no device, and the sharding helpers."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import physics_informed_image_segmentation_b200 as P
from physics_informed_image_segmentation_b200 import _lib, functional as Fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "pil.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pil_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"libpil.so does not export {n}"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)
    assert L.pil_version() == 100
    assert b"diffusion_coeff must be positive" == L.pil_status_string(-6)


def test_c_abi_validation_needs_no_device():
    L = _lib.lib()
    ok = P.LossParams().c()
    assert L.pil_validate_params(ctypes.byref(ok)) == 0
    assert L.pil_validate_params(ctypes.byref(P.LossParams(diffusion_coeff=0.0).c())) == -6
    assert L.pil_validate_params(ctypes.byref(P.LossParams(reaction_threshold=1.0).c())) == -7
    assert L.pil_validate_params(ctypes.byref(P.LossParams(phase_field_weight=1.0, epsilon=0.0).c())) == -8
    assert L.pil_validate_params(ctypes.byref(P.LossParams(phase_field_weight=0.0, epsilon=0.0).c())) == 0
    assert L.pil_workspace_bytes(0, 8, 8) == 0 and L.pil_workspace_bytes(1, 1, 8) == 0
    assert L.pil_workspace_bytes(64, 1024, 1024) >= 256 + 8 * 8
    # argument checks come before any CUDA call
    buf = (ctypes.c_char * 64)()
    a = ctypes.addressof(buf)
    assert L.pil_forward(None, a, 1, 8, 8, 0, 0, 0, ctypes.byref(ok), a, None, a, 64, None) == -1
    assert L.pil_forward(a, a, 1, 1, 8, 0, 0, 0, ctypes.byref(ok), a, None, a, 64, None) == -2
    assert L.pil_forward(a, a, 1, 8, 8, 7, 0, 0, ctypes.byref(ok), a, None, a, 64, None) == -3
    assert L.pil_forward(a, a, 1, 8, 8, 0, 0, 9, ctypes.byref(ok), a, None, a, 64, None) == -4
    assert L.pil_forward(a, a, 1, 8, 8, 0, 0, 0, ctypes.byref(ok), a, None, a, 8, None) == -5
    assert L.pil_forward(a + 2, a, 1, 8, 8, 0, 0, 0, ctypes.byref(ok), a, None, a, 64, None) == -9
    assert L.pil_backward(a, a, None, 1, 8, 8, 0, 0, 0, ctypes.byref(ok), a, 64, None, 1.0, None) == -1


def test_constructor_signatures_and_defaults_match_reference():
    """src/loss.py:24-29, :86-96; src/pde.py:7-11"""
    import inspect

    sig = inspect.signature(P.DiceBCEPDELoss.__init__)
    want = [("dice_weight", 0.5), ("bce_weight", 0.5), ("pde_weight", 1e-3), ("phase_field_weight", 0.0),
            ("smooth", 1e-6), ("diffusion_coeff", 1.0), ("reaction_threshold", 0.5), ("epsilon", 0.05)]
    got = [(k, v.default) for k, v in list(sig.parameters.items())[1:9]]
    assert got == want
    sig = inspect.signature(P.DiceBCELoss.__init__)
    assert [(k, v.default) for k, v in list(sig.parameters.items())[1:4]] == [("dice_weight", 0.5), ("bce_weight", 0.5), ("smooth", 1e-6)]
    sig = inspect.signature(P.PDERegularization.__init__)
    assert [(k, v.default) for k, v in list(sig.parameters.items())[1:]] == [("diffusion_coeff", 1.0), ("reaction_threshold", 0.5)]


def test_attribute_surface_and_errors():
    c = P.DiceBCEPDELoss(0.5, 0.5, 1e-4, 1e-4, 1e-6, 5.0, 0.5, 0.05)  # positional, like a caller may
    assert (c.dice_weight, c.bce_weight, c.pde_weight, c.phase_field_weight, c.smooth, c.epsilon) == (0.5, 0.5, 1e-4, 1e-4, 1e-6, 0.05)
    assert c.pde_regularization.diffusion_coeff == 5.0 and c.pde_regularization.reaction_threshold == 0.5
    assert callable(c.bce) and isinstance(c, torch.nn.Module) and isinstance(c.pde_regularization, P.PDERegularization)
    assert sorted(c.state_dict()) == ["pde_regularization.grad_x_kernel", "pde_regularization.grad_y_kernel",
                                      "pde_regularization.laplacian_kernel"]
    assert c.state_dict()["pde_regularization.laplacian_kernel"].shape == (1, 1, 3, 3)
    assert c.to("cpu") is c
    with pytest.raises(ValueError, match="diffusion_coeff must be positive"):
        P.DiceBCEPDELoss(diffusion_coeff=0.0)
    with pytest.raises(ValueError, match=r"reaction_threshold must be in \(0,1\)"):
        P.PDERegularization(1.0, 0.0)
    with pytest.raises(ValueError, match="epsilon must be positive"):
        P.PDERegularization().compute_phase_field_loss(torch.rand(1, 1, 4, 4), epsilon=0.0)
    # epsilon <= 0 is accepted at construction, like the reference (only checked when the PF term runs)
    P.DiceBCEPDELoss(epsilon=-1.0)
    b = P.DiceBCELoss()
    assert (b.dice_weight, b.bce_weight, b.smooth) == (0.5, 0.5, 1e-6) and callable(b.bce)


def test_no_cpu_fallback():
    c = P.DiceBCEPDELoss()
    u = torch.rand(2, 1, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c(u, torch.zeros_like(u))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.PDERegularization().compute_laplacian(u)
    with pytest.raises(RuntimeError, match="CUDA"):
        P.PDERegularization().reaction_term(u)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "physics_informed_image_segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("# oracle", ""), f"{f} mentions the oracle"


def test_shard_bounds():
    for B in (0, 1, 5, 8, 64, 67):
        for W in (1, 2, 3, 8):
            spans = [P.shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.shard_bounds(4, 2, 2)


def test_loss_report_from_sums_matches_oracle_finalize():
    from oracle import pil_oracle as po

    s = np.array([10.0, 30.0, 25.0, 40.0, 3.0, 7.0, 0.0, 100.0])
    for kw in (dict(), dict(pde_weight=0.0), dict(phase_field_weight=2.0, pde_weight=3.0)):
        p = P.LossParams(**kw)
        rep = P.loss_report_from_sums(s, None, p)
        want = po.finalize(s, 100, po.Params(dice_weight=p.dice_weight, bce_weight=p.bce_weight, pde_weight=p.pde_weight,
                                             phase_field_weight=p.phase_field_weight, smooth=p.smooth,
                                             diffusion_coeff=p.diffusion_coeff, reaction_threshold=p.reaction_threshold,
                                             epsilon=p.epsilon))
        got = [rep[k] for k in ("loss", "dice_loss", "bce_loss", "pde_loss", "phase_field_loss")]
        assert np.allclose(got, want, rtol=1e-15, atol=0)


def test_install_into_reference_rebinds_names():
    import sys
    import types

    pkg = types.ModuleType("fakeref")
    sub = types.ModuleType("fakeref.train")
    class Old:  # noqa: E306
        pass
    sub.DiceBCEPDELoss = Old
    sub.DiceBCELoss = Old
    pkg.PDERegularization = Old
    sys.modules["fakeref"], sys.modules["fakeref.train"] = pkg, sub
    sub.compute_iou_batch = Old
    from physics_informed_image_segmentation_b200.loss import _FusedLossBase
    try:
        rep = P.install_into_reference("fakeref")
        assert sub.DiceBCEPDELoss is P.DiceBCEPDELoss and sub.DiceBCELoss is P.DiceBCELoss
        assert pkg.PDERegularization is P.PDERegularization and len(rep) == 4
        # the per-step metric calls are rebound too, and criteria built from now on leave the counts they need
        assert sub.compute_iou_batch is P.compute_iou_batch
        assert P.DiceBCEPDELoss().batch_metrics_threshold == 0.5
        _FusedLossBase.default_batch_metrics_threshold = None
        sub.compute_iou_batch = Old
        rep = P.install_into_reference("fakeref", track_metrics=False)
        assert sub.compute_iou_batch is Old and rep == [] and P.DiceBCEPDELoss().batch_metrics_threshold is None
    finally:
        _FusedLossBase.default_batch_metrics_threshold = None
        del sys.modules["fakeref"], sys.modules["fakeref.train"]
    # same signatures as the reference's functions (checked live when the checkout is present)
    import inspect
    sig = inspect.signature(P.compute_dice_score_batch)
    assert list(sig.parameters) == ["predictions", "targets", "threshold", "smooth"]
    assert sig.parameters["threshold"].default == 0.5 and sig.parameters["smooth"].default == 1e-6
    for f in (P.compute_dice_score, P.compute_iou, P.compute_iou_batch):
        assert inspect.signature(f) == sig
    if os.path.exists("/root/reference/src/metrics.py"):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_metrics_only", "/root/reference/src/metrics.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        for name in ("compute_dice_score", "compute_dice_score_batch"):
            rs = inspect.signature(getattr(ref, name))
            assert list(rs.parameters) == list(sig.parameters)
            assert all(rs.parameters[k].default == sig.parameters[k].default for k in ("threshold", "smooth"))

    class M:
        activation_name = "sigmoid"
    m = M()
    with P.use_logits_head(m) as name:
        assert name == "sigmoid" and m.activation_name == "none"
    assert m.activation_name == "sigmoid"


def test_exchange_descriptor_layout_and_validation():
    """PilExchange is a plain struct shared with C: its layout must match include/pil.h, and a bad
    descriptor is rejected before any CUDA call."""
    L = _lib.lib()
    assert ctypes.sizeof(_lib.PilExchange) == 4 + 4 + 8 + 4 + 4 + 8 * _lib.PIL_MAX_RANKS
    assert _lib.PilExchange.mailbox.offset == 24
    text = open(os.path.join(ROOT, "include", "pil.h")).read()
    assert f"#define PIL_MAX_RANKS {_lib.PIL_MAX_RANKS}" in text
    assert f"#define PIL_IPC_HANDLE_BYTES {_lib.PIL_IPC_HANDLE_BYTES}" in text
    assert f"#define PIL_NMOMENTS {_lib.PIL_NMOMENTS}" in text
    assert L.pil_exchange_bytes() >= 2 * 2 * _lib.PIL_MAX_RANKS * 128 + 4
    ok = P.LossParams().c()
    wsb = L.pil_workspace_bytes(1, 8, 8)
    buf = (ctypes.c_char * wsb)()
    a = ctypes.addressof(buf)
    ex = _lib.PilExchange()
    ex.rank, ex.world, ex.epoch = 0, 2, 0          # mailbox pointers NULL
    assert L.pil_forward_pointwise_xchg(a, a, 1, 8, 8, 0, 0, 1, ctypes.byref(ok), a, a, wsb, ctypes.byref(ex), None) == -11
    ex.rank, ex.world = 3, 2                       # rank out of range
    ex.mailbox[0] = a
    ex.mailbox[1] = a
    assert L.pil_forward_pointwise_xchg(a, a, 1, 8, 8, 0, 0, 1, ctypes.byref(ok), a, a, wsb, ctypes.byref(ex), None) == -11
    ex.rank, ex.world = 0, _lib.PIL_MAX_RANKS + 1  # more ranks than a mailbox has slots for
    assert L.pil_exchange_finalize(ctypes.byref(ex), -1, ctypes.byref(ok), a, a, None) == -11
    assert L.pil_forward_pointwise_xchg(a, a, 1, 8, 8, 0, 0, 1, ctypes.byref(ok), a, a, wsb, None, None) == -1


def test_sweep_grids_match_reference_definitions():
    """s2_grid / s3_grid restate run_ablation.py:159-224; checked against the live definitions when the
    reference checkout is present (build container), against the documented values otherwise."""
    s2, s3 = P.s2_grid(), P.s3_grid()
    assert [p.diffusion_coeff for p in s2] == [0.5, 1.0, 2.0, 5.0, 10.0, 100.0]
    assert all(p.pde_weight == 1e-3 and p.phase_field_weight == 0.0 for p in s2)
    assert [p.epsilon for p in s3] == [0.001, 0.01, 0.05, 0.1, 0.2]
    assert all(p.pde_weight == 1e-4 and p.phase_field_weight == 1e-4 and p.diffusion_coeff == 5.0
               and p.reaction_threshold == 0.5 for p in s3)
    ref = "/root/reference/run_ablation.py"
    if os.path.exists(ref):
        src = open(ref).read()
        assert "for i, d in enumerate([0.5, 1.0, 2.0, 5.0, 10.0, 100.0])" in src
        assert "for i, eps in enumerate([0.001, 0.01, 0.05, 0.1, 0.2])" in src
    # sweep_finalize validates every setting before touching the device
    with pytest.raises(ValueError, match="reaction_threshold must be in"):
        Fn.sweep_finalize(torch.zeros(16, dtype=torch.float64), 1, [P.LossParams(reaction_threshold=0.0)])


def test_sweep_closed_form_matches_direct_evaluation():
    """The algebra behind pil_sweep_finalize, on the CPU with the oracle: sum r^2 for any (D, a) from the six
    second moments of {lap, g, h} equals the oracle's direct sum (fp64)."""
    from oracle import pil_oracle as po

    rng = np.random.default_rng(0)
    u = rng.random((2, 1, 17, 23))
    t = (rng.random((2, 1, 17, 23)) > 0.5).astype(np.float64)
    up = np.pad(u, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="reflect")
    lap = up[..., 2:, 1:-1] + up[..., :-2, 1:-1] + up[..., 1:-1, 2:] + up[..., 1:-1, :-2] - 4.0 * u
    g = u * (1.0 - u)
    h = g * u
    for D, a in [(0.5, 0.5), (5.0, 0.3), (100.0, 0.7)]:
        closed = (D * D * (lap * lap).sum() + 2 * D * (lap * h).sum() - 2 * a * D * (lap * g).sum() + (h * h).sum()
                  - 2 * a * (h * g).sum() + a * a * (g * g).sum())
        s = po.sums(u, t, po.Params(diffusion_coeff=D, reaction_threshold=a), po.X_PROB)
        assert abs(closed - s[4]) / s[4] < 1e-12


def test_header_is_plain_c(tmp_path):
    """include/pil.h compiles as C99 (no C++-isms) together with a translation unit that takes the entry points by
    their documented signatures -- what a cgo / JNI / ctypes binding of the boundary relies on."""
    import shutil
    import subprocess

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    out = tmp_path / "abi_check.o"
    res = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c",
                          os.path.join(ROOT, "tests", "abi_check.c"), "-o", str(out)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_host_session_rejects_buffers_the_library_would_misread():
    """The host-buffer entry takes raw pointers: non-contiguous arrays (a temporary copy would dangle, or swallow
    the gradient), mismatched shapes / dtypes and read-only gradient buffers must be refused BEFORE the C call."""
    from physics_informed_image_segmentation_b200 import session as S

    shape = (4, 16, 24)
    x = np.zeros((4, 1, 16, 24), np.float32)
    t = np.zeros((4, 1, 16, 24), np.float32)
    assert S._check_buffers(x, t, np.zeros_like(x), shape, _lib.F32, _lib.F32) == 4
    assert S._check_buffers(x[:2], t[:2], None, shape, _lib.F32, _lib.F32) == 2
    with pytest.raises(ValueError):
        S._check_buffers(x, t[:3], None, shape, _lib.F32, _lib.F32)                   # smaller target buffer
    with pytest.raises(ValueError):
        S._check_buffers(x, t, np.zeros((4, 1, 16, 23), np.float32), shape, _lib.F32, _lib.F32)  # gradient shape
    with pytest.raises(TypeError):
        S._check_buffers(x, t, np.zeros(x.shape, np.float64), shape, _lib.F32, _lib.F32)         # gradient dtype
    with pytest.raises(TypeError):
        S._check_buffers(x.astype(np.float64), t, None, shape, _lib.F32, _lib.F32)
    with pytest.raises(ValueError):
        S._check_buffers(np.zeros((5, 1, 16, 24), np.float32), np.zeros((5, 1, 16, 24), np.float32), None, shape, _lib.F32, _lib.F32)
    ro = np.zeros_like(x)
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        S._check_buffers(x, t, ro, shape, _lib.F32, _lib.F32)
    big = np.zeros((4, 1, 16, 48), np.float32)
    with pytest.raises(ValueError):
        S._host_ptr(big[..., ::2], "maps")                                              # strided view: never copied silently
    with pytest.raises(ValueError):
        S._host_ptr(torch.zeros(4, 1, 16, 48)[..., ::2], "maps")
    with pytest.raises(TypeError):
        S._host_ptr([[1.0, 2.0]], "maps")
    assert S._host_ptr(x, "maps") == x.ctypes.data
