"""Two-rank NCCL data-parallel step on real GPUs equals the single-GPU step (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600)
def test_two_gpu_data_parallel_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=580)
    print(r.stdout[-3000:])
    assert "DP_CHECK_OK" in r.stdout
