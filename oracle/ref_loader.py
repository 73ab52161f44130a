"""Import the REAL reference modules by file path (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference is read-only and does not exist on the GPU box, so this
is used solely (a) by tests/golden/make_golden.py to generate the committed golden vectors and
(b) by CPU tests that skip when the checkout is absent.  `import src` cannot be used:
src/__init__.py:7 pulls in matplotlib, which is not installed; loading
src/pde.py and src/loss.py under a synthetic package resolves the relative import at src/loss.py:4.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("PIL_REFERENCE_ROOT", "/root/reference")
_PKG = "_pil_refsrc"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "loss.py"))


def _load(name: str):
    full = f"{_PKG}.{name}"
    if full in sys.modules:
        return sys.modules[full]
    if _PKG not in sys.modules:
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = [os.path.join(REF_ROOT, "src")]
        sys.modules[_PKG] = pkg
    spec = importlib.util.spec_from_file_location(full, os.path.join(REF_ROOT, "src", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def pde():
    return _load("pde")


def loss():
    _load("pde")
    return _load("loss")


def unet():
    return _load("unet")


def metrics():
    return _load("metrics")


def evaluate():
    """src/evaluate.py (compute_iou, compute_iou_batch): needs scipy + cv2 (present in the build container)
    and, through its relative imports, src/metrics.py and src/dataset.py."""
    _load("metrics")
    try:
        _load("dataset")
    except Exception:  # dataset.py may want packages that are absent; evaluate.py only needs the name
        stub = types.ModuleType(f"{_PKG}.dataset")
        stub.CellSegmentationDataset = object
        sys.modules[f"{_PKG}.dataset"] = stub
    return _load("evaluate")


def train():
    """src/train.py (train_epoch, validate, train_stage).  It imports src/plot.py, which needs matplotlib -- not
    installed here -- so a stub `matplotlib.pyplot` is registered first; nothing on the step path touches it."""
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    _load("pde")
    _load("loss")
    _load("unet")
    evaluate()
    _load("plot")
    return _load("train")
