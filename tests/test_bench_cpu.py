"""bench.py on a box without a GPU: the reference arm (the CPU implementation of the path on the host cores) must print
the contract's JSON line, and the product arm must refuse to run -- there is no CPU fallback to fall back to."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")   # also on a GPU box this test looks at the GPU-less behaviour
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "pde_loss_fwd_bwd_gpixels_per_s" and line["unit"] == "Gpixel/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == pytest.approx(line["value"])
    e2e = line["e2e"]
    assert e2e["value"] == pytest.approx(line["value"]) and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("stage2_loss_fwd_bwd_64x1x1024x1024")


def test_product_arm_refuses_to_run_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "0", timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
