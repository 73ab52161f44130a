"""Generate tests/golden/ref_step.npz by RUNNING THE REAL REFERENCE's train_epoch / validate (src/train.py:84-286, with
its own DiceBCEPDELoss / DiceBCELoss, its Python-loop metrics and its OpenCV boundary-F1) on a tiny model and fixed
batches (build container only).

    python tests/golden/make_golden_step.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from tests.step_model import TinySegNet, make_batches  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_step.npz")
STAGE2 = dict(dice_weight=0.5, bce_weight=0.5, pde_weight=1e-4, phase_field_weight=1e-4, diffusion_coeff=5.0,
              reaction_threshold=0.5, epsilon=0.05)


def main():
    assert ref_loader.available(), "needs the reference checkout (build container)"
    ref_train, ref_loss = ref_loader.train(), ref_loader.loss()
    torch.set_num_threads(1)
    out, meta = {}, {"torch": torch.__version__, "cases": []}
    batches = make_batches(3, 4, 32, 48, seed=11)
    out["images"] = np.stack([b[0].numpy() for b in batches])
    out["masks"] = np.stack([b[1].numpy() for b in batches])
    for case, (crit_name, kw, act) in {
        "stage2_sigmoid": ("DiceBCEPDELoss", STAGE2, "sigmoid"),
        "stage1_sigmoid": ("DiceBCELoss", dict(dice_weight=0.5, bce_weight=0.5), "sigmoid"),
        "stage2_tanh": ("DiceBCEPDELoss", dict(STAGE2, pde_weight=1e-3, phase_field_weight=0.0), "tanh"),
    }.items():
        torch.manual_seed(123)
        model = TinySegNet(4, act)
        for k, v in model.state_dict().items():
            out[f"{case}/init/{k}"] = v.numpy().copy()
        crit = getattr(ref_loss, crit_name)(**kw)
        opt = torch.optim.SGD(model.parameters(), lr=0.5)
        res_t = ref_train.train_epoch(model, batches, crit, opt, torch.device("cpu"), return_components=True, compute_metrics=True)
        res_v = ref_train.validate(model, batches, crit, torch.device("cpu"), return_components=True, compute_metrics=True)
        for k, v in model.state_dict().items():
            out[f"{case}/final/{k}"] = v.numpy().copy()
        meta["cases"].append({"name": case, "criterion": crit_name, "kwargs": kw, "activation": act,
                              "train": {k: float(v) for k, v in res_t.items()}, "validate": {k: float(v) for k, v in res_v.items()}})
        print(case, "train", {k: round(float(v), 6) for k, v in res_t.items()})
        print(case, "valid", {k: round(float(v), 6) for k, v in res_v.items()})
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
