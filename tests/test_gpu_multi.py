"""Multi-GPU path on real GPUs (skipped on a single-GPU box): tools/check_multigpu.py under torchrun, 2 ranks."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_ranks_all_modes_equal_single_gpu(built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_multigpu.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "MULTIGPU OK world=2" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
