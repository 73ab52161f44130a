"""Offline guards on the compiled kernels (no GPU needed: cuobjdump on the in-tree libpil.so).

The backward kernel is sensitive to what ptxas does with its 6x-unrolled steady loop: builds in which the
per-strip constants were rematerialised inside the loop (750-760 SASS instructions per 6 rows instead of
~690) measured 10-17% slower on the B200.  This test fails such a build before it reaches a GPU."""
import re
import shutil
import subprocess

import pytest

from physics_informed_image_segmentation_b200 import _lib

BWD_F32 = "_ZN3pil14pil_bwd_kernelILi1EffLb1EEEvNS_7BwdArgsE"   # pil_bwd_kernel<LOGITS_SIGMOID, float, float, ALIGNED>
POINT_F32 = "_ZN3pil16pil_point_kernelILi1EffLb1EEEvNS_9PointArgsE"


def _sass(fun):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-sass", "-fun", fun, _lib.build()], capture_output=True, text=True, timeout=120).stdout
    except (FileNotFoundError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    if not ins:
        pytest.skip("kernel not found in the library (symbol naming changed?)")
    return ins


def _loops(ins):
    """(instructions, rows) of every backward-branch loop that contains activations (4 MUFU.EX2 per row)"""
    out = []
    for a, t in ins:
        m = re.search(r"BRA\s+(?:U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            body = [tt for x, tt in ins if int(m.group(1), 16) <= x <= a]
            ex = sum(1 for tt in body if "MUFU.EX2" in tt)
            if ex >= 4:
                out.append((len(body), ex // 4))
    return out


def test_backward_steady_loop_is_lean_and_native():
    ins = _sass(BWD_F32)
    text = "\n".join(t for _, t in ins)
    assert "LDGSTS" in text, "cp.async stage ring missing"
    assert "FFMA2" in text, "packed fp32x2 arithmetic missing"
    steady = [n for n, rows in _loops(ins) if rows == 6]
    assert steady, "6x unrolled steady loop not found"
    assert min(steady) <= 720, f"steady loop grew to {min(steady)} instructions per 6 rows (expected ~690)"
    body_local = [t for _, t in ins if re.search(r"\b(LDL|STL)\b", t)]
    assert len(body_local) < 160, "unexpected amount of local-memory traffic (spills?)"


def test_pointwise_forward_uses_three_mufu_per_pixel():
    ins = _sass(POINT_F32)
    assert _loops(ins), "pointwise loop not found"
    text = [t for _, t in ins]
    ex2 = sum(1 for t in text if "MUFU.EX2" in t)
    rcp = sum(1 for t in text if re.search(r"MUFU\.RCP\b", t))   # not the fp64 RCP64H of the finalisation
    lg2 = sum(1 for t in text if "MUFU.LG2" in t)
    assert ex2 == rcp == lg2 and ex2 > 0, (ex2, rcp, lg2)  # den = 1 + 2^xs, u = 1/den, L = lg2(den): 3 MUFU per pixel
