"""Offline guards on the compiled kernels (no GPU needed: cuobjdump on the in-tree libpil.so).

The backward kernel is sensitive to what ptxas does with its 6x-unrolled steady loop: builds in which the
per-strip constants were rematerialised inside the loop (750-760 SASS instructions per 6 rows instead of
~690) measured 10-17% slower on the B200.  This test fails such a build before it reaches a GPU."""
import re
import shutil
import subprocess

import pytest

from physics_informed_image_segmentation_b200 import _lib

BWD_F32 = "_ZN3pil14pil_bwd_kernelILi1EffLb1EEEvNS_7BwdArgsE"   # pil_bwd_kernel<LOGITS_SIGMOID, float, float, ALIGNED> (cp.async ring)
BWD_TMA_F32 = "_ZN3pil18pil_bwd_kernel_tmaILi1EffEEvNS_7BwdArgsE14CUtensorMap_stS2_"   # the default: rows staged by TMA boxes
BWD_TMA_BF16 = "_ZN3pil18pil_bwd_kernel_tmaILi1E13__nv_bfloat16S1_EEvNS_7BwdArgsE14CUtensorMap_stS3_"
FWD_TMA_F32 = "_ZN3pil18pil_fwd_kernel_tmaILi1EffLb0EEEvNS_7FwdArgsE14CUtensorMap_stS2_"
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


def _no_local_memory(ins):
    """No LDL / STL at all: a device function that is NOT inlined and takes the kernel's argument block by reference
    makes every thread copy the block (368 bytes) to local memory on entry -- 50 MB of extra DRAM traffic per launch
    and L2 round trips in the serial last-block epilogue (backward 143 -> 135 us when that copy went away)."""
    local = [t for _, t in ins if re.search(r"\b(LDL|STL)\b", t)]
    assert not local, f"{len(local)} local-memory instructions (argument block copied to the stack, or spills): {local[:3]}"


def test_backward_steady_loop_is_lean_and_native():
    ins = _sass(BWD_F32)
    text = "\n".join(t for _, t in ins)
    assert "LDGSTS" in text, "cp.async stage ring missing"
    assert "FFMA2" in text, "packed fp32x2 arithmetic missing"
    steady = [n for n, rows in _loops(ins) if rows == 6]
    assert steady, "6x unrolled steady loop not found"
    assert min(steady) <= 665, f"steady loop grew to {min(steady)} instructions per 6 rows (round 2: 646)"
    _no_local_memory(ins)


def test_default_backward_stages_rows_with_tma():
    """The kernel bench.py times: TMA boxes (UTMALDG) completed on an mbarrier (SYNCS ... TRYWAIT), no per-lane cp.async
    in the steady loop, and a loop no larger than the cp.async variant's."""
    ins = _sass(BWD_TMA_F32)
    text = "\n".join(t for _, t in ins)
    assert "UTMALDG.2D" in text, "cp.async.bulk.tensor did not make it into the kernel"
    assert "SYNCS.PHASECHK.TRANS64.TRYWAIT" in text and "SYNCS.ARRIVE.TRANS64" in text, "mbarrier pipeline missing"
    assert "LDGSTS" not in text, "the TMA variant must not fall back to per-lane cp.async"
    assert "FFMA2" in text
    steady = [n for n, rows in _loops(ins) if rows == 6]
    assert steady and min(steady) <= 660, f"TMA steady loop: {steady} instructions per 6 rows (round 2: 642)"
    _no_local_memory(ins)
    # bf16 maps run the same arithmetic: the instantiation may add the unpack / pack instructions only
    ins_b = _sass(BWD_TMA_BF16)
    steady_b = [n for n, rows in _loops(ins_b) if rows == 6]
    assert steady_b and min(steady_b) <= 725, f"bf16 TMA steady loop: {steady_b} (round 2: 704)"
    _no_local_memory(ins_b)


def test_full_forward_stages_rows_with_tma():
    ins = _sass(FWD_TMA_F32)
    text = "\n".join(t for _, t in ins)
    assert "UTMALDG.2D" in text and "LDGSTS" not in text


def test_pointwise_forward_uses_three_mufu_per_pixel():
    ins = _sass(POINT_F32)
    assert _loops(ins), "pointwise loop not found"
    text = [t for _, t in ins]
    ex2 = sum(1 for t in text if "MUFU.EX2" in t)
    rcp = sum(1 for t in text if re.search(r"MUFU\.RCP\b", t))   # not the fp64 RCP64H of the finalisation
    lg2 = sum(1 for t in text if "MUFU.LG2" in t)
    assert ex2 == rcp == lg2 and ex2 > 0, (ex2, rcp, lg2)  # den = 1 + 2^xs, u = 1/den, L = lg2(den): 3 MUFU per pixel


def test_unity_development_build_compiles(tmp_path):
    """csrc/pil_unity.cu is the single-translation-unit form the instrumented development builds use (-DPIL_TIMELINE for
    tools/timeline.py, -DPIL_BOUNDS for tools/bounds_check.py): it must keep compiling as the pieces evolve."""
    import os

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = tmp_path / "unity.o"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-Xcompiler", "-fPIC", "-DPIL_BOUNDS", "-DPIL_TIMELINE",
           "-DPIL_DEV_F32_ONLY", "-I", _lib.INCLUDE, "-I", _lib.CSRC, "-c", "-o", str(out), os.path.join(_lib.CSRC, "pil_unity.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
