"""Launches tests/multigpu_check.py under torchrun when at least two GPUs are visible (real peer memory,
real NCCL).  On a single-GPU box it skips; the host logic is covered by tests/test_sharding_gloo.py and
the kernels' exchange code by tests/test_exchange_gpu.py (ranks emulated on one GPU)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_real_gpus_module_path():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "multigpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "multigpu_check ok" in res.stdout
