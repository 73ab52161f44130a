"""Pin the CPU oracle (oracle/pil_oracle.c and oracle/torch_port.py) against the golden vectors that
tests/golden/make_golden.py produced by running the real reference, and -- when the reference
checkout is present (build container) -- against the reference itself on fresh inputs."""
import numpy as np
import pytest
import torch

from oracle import pil_oracle as po
from oracle import ref_loader, torch_port
from tests.helpers import blob_inputs, iid_inputs, rel_l2, rel_max, rel_scalar

KIND = {"sigmoid": po.X_LOGITS_SIGMOID, "tanh": po.X_LOGITS_TANH}


def _params(d):
    return po.Params(**d)


def test_known_answer_stencils(golden):
    data, _ = golden
    u = data["ka_u"]
    # SURVEY.md section 4 known answers, re-derived from the reference by make_golden.py
    assert np.array_equal(data["ka_lap"].reshape(3, 3), [[6, 7, 6], [6, 2, -16], [-6, -9, 10]])
    assert np.array_equal(data["ka_gms"].reshape(3, 3), [[0, 2.25, 0], [9, 18, 1], [0, 0.25, 0]])
    assert np.array_equal(data["ka_pad1d"].reshape(-1), [2, 1, 2, 3, 4, 3])
    for dt in (np.float32, np.float64):
        assert np.array_equal(po.laplacian(u.astype(dt)), data["ka_lap"].astype(dt))
        assert np.array_equal(po.grad_mag_sq(u.astype(dt)), data["ka_gms"].astype(dt))
    tu = torch.from_numpy(u)
    assert torch.equal(torch_port.laplacian(tu), torch.from_numpy(data["ka_lap"]))
    assert torch.equal(torch_port.grad_mag_sq(tu), torch.from_numpy(data["ka_gms"]))


def _case_names(meta):
    return [c["name"] for c in meta["cases"]]


def test_golden_has_all_cases(golden):
    _, meta = golden
    assert len(meta["cases"]) >= 12


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_c_oracle_matches_reference_golden(golden, prec):
    data, meta = golden
    dt = np.float32 if prec == "f32" else np.float64
    # fp64 oracle vs fp64 reference: identical math up to summation order;
    # fp32 oracle vs fp32 reference: both carry fp32 rounding noise (SURVEY 8c noise floor ~2.6e-7)
    tol_loss, tol_grad = (2e-6, 3e-6) if prec == "f32" else (1e-12, 1e-11)
    for c in meta["cases"]:
        n, p, kind = c["name"], _params(c["params"]), KIND[c["activation"]]
        z, t = data[f"{n}.z"].astype(dt), data[f"{n}.t"].astype(dt)
        # logits entry
        comps, dz = po.loss_and_grad(z, t, p, kind)
        if n != "saturated" or prec == "f64":
            # the reference's own fp32-vs-fp64 disagreement is the floor (tanh head: 2.7e-5 on dz
            # because autograd forms 1 - tanh^2 in fp32) -- never demand more than the reference delivers
            floor_l = rel_scalar(data[f"{n}.f32.loss"], data[f"{n}.f64.loss"])
            floor_g = rel_max(data[f"{n}.f32.dz"], data[f"{n}.f64.dz"])
            assert rel_scalar(comps[0], data[f"{n}.{prec}.loss"]) < max(tol_loss, 4 * floor_l), n
            assert rel_max(dz, data[f"{n}.{prec}.dz"]) < max(tol_grad, 4 * floor_g), n
        # probability entry: bit-identical fp32 probabilities fed to both sides
        if c["activation"] == "sigmoid":
            u = torch.sigmoid(torch.from_numpy(data[f"{n}.z"])).numpy().astype(dt)
        else:
            u = ((torch.tanh(torch.from_numpy(data[f"{n}.z"])) + 1.0) / 2.0).numpy().astype(dt)
        comps_p, du = po.loss_and_grad(u, t, p, po.X_PROB)
        assert rel_scalar(comps_p[0], data[f"{n}.{prec}.loss_p"]) < tol_loss, n
        assert rel_max(du, data[f"{n}.{prec}.du"]) < tol_grad, n
        assert rel_l2(du, data[f"{n}.{prec}.du"]) < tol_grad, n
        ref_c = data[f"{n}.{prec}.comps_p"]
        for k in range(4):
            assert rel_scalar(comps_p[1 + k], ref_c[k]) < tol_loss, (n, k)
        assert rel_max(po.laplacian(u), data[f"{n}.{prec}.lap_p"]) < (1e-6 if prec == "f32" else 1e-14)
        assert rel_max(po.grad_mag_sq(u), data[f"{n}.{prec}.gms_p"]) < (1e-6 if prec == "f32" else 1e-14)


def test_saturated_case_exact_semantics(golden):
    """u==0 / u==1 in fp32: BCE log clamp at -100 and zero gradient through the sigmoid."""
    data, meta = golden
    c = next(c for c in meta["cases"] if c["name"] == "saturated")
    p = _params(c["params"])
    z, t = data["saturated.z"], data["saturated.t"]
    u = torch.sigmoid(torch.from_numpy(z)).numpy()
    comps_p, du = po.loss_and_grad(u, t, p, po.X_PROB)
    assert rel_scalar(comps_p[0], data["saturated.f32.loss_p"]) < 2e-6
    assert rel_max(du, data["saturated.f32.du"]) < 3e-6
    # through the sigmoid, with libm expf instead of torch's vectorised exp, u may differ by an ulp
    # away from saturation; the saturated pixels themselves must agree exactly (dz == 0 there)
    _, dz = po.loss_and_grad(z, t, p, po.X_LOGITS_SIGMOID)
    ref_dz = data["saturated.f32.dz"]
    sat = np.abs(z) >= 17.0
    assert sat.sum() >= 20
    assert np.array_equal(dz[sat & (np.abs(z) > 90)], ref_dz[sat & (np.abs(z) > 90)])
    assert rel_max(dz, ref_dz) < 3e-6


def test_torch_port_bit_identical_to_reference_golden(golden):
    data, meta = golden
    if meta["torch"] != torch.__version__:
        pytest.skip("golden generated with a different torch build; bit-identity is per build")
    torch.set_num_threads(1)
    for c in meta["cases"]:
        n, p, kind = c["name"], _params(c["params"]), KIND[c["activation"]]
        z, t = torch.from_numpy(data[f"{n}.z"]), torch.from_numpy(data[f"{n}.t"])
        L, dz = torch_port.fwd_bwd(z, t, p, kind)
        assert np.array_equal(L.numpy(), data[f"{n}.f32.loss"]), n
        assert np.array_equal(dz.numpy(), data[f"{n}.f32.dz"]), n


def test_dice_bce_only_module_golden(golden):
    data, meta = golden
    for c in meta["cases"]:
        n = c["name"]
        p = _params(dict(c["params"], pde_weight=0.0, phase_field_weight=0.0))
        act = torch.sigmoid if c["activation"] == "sigmoid" else (lambda v: (torch.tanh(v) + 1.0) / 2.0)
        u = act(torch.from_numpy(data[f"{n}.z"])).numpy()
        comps, du = po.loss_and_grad(u, data[f"{n}.t"], p, po.X_PROB)
        assert rel_scalar(comps[0], data[f"{n}.f32.dicebce_loss"]) < 2e-6, n
        assert rel_max(du, data[f"{n}.f32.dicebce_du"]) < 3e-6, n


def test_sharded_sums_equal_unsharded():
    """Data-parallel identity: per-shard sums add up, backward with global sums per shard
    equals the unsharded gradient."""
    z, t = iid_inputs(6, 12, 20, seed=7)
    z, t = z.numpy().astype(np.float64), t.numpy().astype(np.float64)
    p = po.STAGE2
    full = po.sums(z, t, p, po.X_LOGITS_SIGMOID)
    parts = [po.sums(z[a:b], t[a:b], p, po.X_LOGITS_SIGMOID) for a, b in ((0, 1), (1, 4), (4, 6))]
    tot = np.sum(parts, axis=0)
    assert np.allclose(tot, full, rtol=1e-13, atol=0)
    n = int(full[7])
    g_full = po.backward(z, t, p, full, n, po.X_LOGITS_SIGMOID)
    g_parts = np.concatenate([po.backward(z[a:b], t[a:b], p, tot, n, po.X_LOGITS_SIGMOID)
                              for a, b in ((0, 1), (1, 4), (4, 6))])
    assert rel_max(g_parts, g_full) < 1e-13


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout only exists in the build container")
@pytest.mark.parametrize("gen,shape", [(iid_inputs, (8, 64, 64)), (blob_inputs, (4, 96, 80))])
def test_oracle_vs_live_reference(gen, shape):
    """Fresh inputs through the real reference (fp32 and fp64) vs the C oracle."""
    ref = ref_loader.loss()
    z, t = gen(*shape, seed=99)
    p = po.STAGE2
    kw = dict(dice_weight=p.dice_weight, bce_weight=p.bce_weight, pde_weight=p.pde_weight,
              phase_field_weight=p.phase_field_weight, smooth=p.smooth, diffusion_coeff=p.diffusion_coeff,
              reaction_threshold=p.reaction_threshold, epsilon=p.epsilon)
    for dt, tol in ((torch.float64, 1e-11), (torch.float32, 3e-6)):
        crit = ref.DiceBCEPDELoss(**kw)
        crit = crit.double() if dt == torch.float64 else crit
        zz = z.to(dt).clone().requires_grad_(True)
        L = crit(torch.sigmoid(zz), t.to(dt))
        L.backward()
        comps, dz = po.loss_and_grad(z.to(dt).numpy(), t.to(dt).numpy(), p, po.X_LOGITS_SIGMOID)
        assert rel_scalar(comps[0], L.item()) < tol
        assert rel_max(dz, zz.grad.numpy()) < tol
        assert rel_l2(dz, zz.grad.numpy()) < tol
    # and the op-for-op torch port is bit-identical to the reference in this build
    torch.set_num_threads(1)
    crit = ref.DiceBCEPDELoss(**kw)
    zz = z.clone().requires_grad_(True)
    L = crit(torch.sigmoid(zz), t)
    L.backward()
    Lp, dzp = torch_port.fwd_bwd(z, t, p, 1)
    assert torch.equal(Lp, L.detach()) and torch.equal(dzp, zz.grad)


def test_finalize_gates():
    s = np.array([10.0, 30.0, 25.0, 40.0, 3.0, 7.0, 0.0, 100.0])
    p = po.Params(pde_weight=0.0, phase_field_weight=0.0)
    out = po.finalize(s, 100, p)
    dice = 1 - (20 + 1e-6) / (55 + 1e-6)
    assert out[0] == pytest.approx(0.5 * dice + 0.5 * 0.4, rel=1e-14)
    p2 = po.Params(pde_weight=2.0, phase_field_weight=3.0)
    assert po.finalize(s, 100, p2)[0] == pytest.approx(0.5 * dice + 0.5 * 0.4 + 2 * 0.03 + 3 * 0.07, rel=1e-14)


# ------------------------------------------------------------------------------------------------
# golden vectors of the widened path (tests/golden/make_golden_ext.py): S2/S3 grids and Dice/IoU metrics,
# produced by the real reference
# ------------------------------------------------------------------------------------------------
def _ext():
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ext.npz")
    data = np.load(path)
    return data, json.loads(bytes(data["meta"]).decode())


def test_oracle_matches_reference_on_the_s2_s3_grids():
    """the C oracle, evaluated once per grid setting, against the reference's DiceBCEPDELoss per setting"""
    from oracle import pil_oracle as po

    data, meta = _ext()
    u64, t64 = data["u"].astype(np.float64), data["t"].astype(np.float64)
    for k, kw in enumerate(meta["grid"]):
        p = po.Params(dice_weight=0.5, bce_weight=0.5, smooth=1e-6, **kw)
        s = po.sums(u64, t64, p, po.X_PROB)
        comps = po.finalize(s, int(s[7]), p)
        ref = data["sweep_f64"][k]
        for c in range(5):
            if c == 4 and not kw["phase_field_weight"] > 0:
                assert abs(comps[c] - ref[c]) <= 1e-11 * abs(ref[c])  # logged by the reference even when its weight is 0
                continue
            assert abs(comps[c] - ref[c]) <= 1e-11 * max(abs(ref[c]), 1e-30), (k, c, comps[c], ref[c])
        # and the reference's own fp32 evaluation agrees with its fp64 one to fp32 noise
        assert abs(data["sweep_f32"][k][0] - ref[0]) <= 2e-6 * abs(ref[0])


def test_metric_formulas_match_reference_golden():
    """the three-counts formulation (what pil_image_metrics evaluates) against the reference's functions"""
    data, _ = _ext()
    u, t = data["u"], data["t"]
    for thr, tag in ((0.5, "thr5"), (0.3, "thr3")):
        pb = (u > thr).astype(np.float64)
        B = u.shape[0]
        I = (pb * t).reshape(B, -1).sum(1)
        P = pb.reshape(B, -1).sum(1)
        T = t.astype(np.float64).reshape(B, -1).sum(1)
        s = 1e-6
        assert np.allclose((2 * I + s) / (P + T + s), data[f"dice_batch_{tag}"], rtol=2e-6, atol=0)
        assert np.allclose((I + s) / (P + T - I + s), data[f"iou_batch_{tag}"], rtol=2e-6, atol=0)
        assert abs((2 * I.sum() + s) / (P.sum() + T.sum() + s) - data[f"dice_{tag}"]) < 2e-6
        assert abs((I.sum() + s) / (P.sum() + T.sum() - I.sum() + s) - data[f"iou_{tag}"]) < 2e-6
