"""Multi-GPU (NCCL) check of the sharded operator - needs at least two visible GPUs, skipped otherwise.
The script it launches (tests/gpu_sharding_nccl.py) can also be run by hand under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_operator_over_nccl():
    here = os.path.dirname(os.path.abspath(__file__))
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(here, "gpu_sharding_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "gpu_sharding_nccl ok" in r.stdout
