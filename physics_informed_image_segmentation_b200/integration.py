"""Glue for running the UNMODIFIED reference training code on top of this package.

`install_into_reference()` rebinds the four names the reference exports for this path
(`DiceBCELoss`, `DiceBCEPDELoss`, `PDERegularization`, `create_pde_regularization`; reference
src/__init__.py:3-4) in every already-imported module of the reference package, so that
src/train.py:656,:708 and src/ablation.py:102,:146 construct the fused modules and the
`isinstance(criterion, DiceBCEPDELoss)` checks at src/train.py:142,:174 keep working.

`use_logits_head(model)` makes the reference UNet return logits without touching its code or its
state dict: UNet.forward applies the activation under `if self.activation_name == ...` tests
(src/unet.py:208-214), so any other value skips it.
"""
from __future__ import annotations

import contextlib
import sys

from .loss import DiceBCELoss, DiceBCEPDELoss
from .pde import PDERegularization, create_pde_regularization

_NAMES = {
    "DiceBCELoss": DiceBCELoss,
    "DiceBCEPDELoss": DiceBCEPDELoss,
    "PDERegularization": PDERegularization,
    "create_pde_regularization": create_pde_regularization,
}


def install_into_reference(package: str = "src") -> list:
    """Rebind the loss/PDE names inside the imported reference package.  Returns the list of
    (module, name) pairs that were replaced."""
    replaced = []
    for modname, mod in list(sys.modules.items()):
        if mod is None or not (modname == package or modname.startswith(package + ".")):
            continue
        for name, new in _NAMES.items():
            if hasattr(mod, name) and getattr(mod, name) is not new:
                setattr(mod, name, new)
                replaced.append((modname, name))
    return replaced


@contextlib.contextmanager
def use_logits_head(model):
    """Within the context `model(images)` returns logits; pair with criterion.forward_logits(...,
    activation=<the restored name>).  Yields the original activation name."""
    original = model.activation_name
    model.activation_name = "none"
    try:
        yield original
    finally:
        model.activation_name = original
