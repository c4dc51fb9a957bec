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


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpus_one_process():
    """One process driving two GPUs (DataParallel / pipeline style): the dynamic shared-memory attribute, the SM
    count and the backward workspaces are per device (they were process-wide)."""
    import numpy as np
    from flash_cosine_sim_attention_b200 import flash_cosine_sim_attention
    from oracle import cosine_sim_attention_oracle as oracle
    g = torch.Generator().manual_seed(61)
    q, k, v, do = (torch.randn(2, 4, 300, 64, generator=g).to(torch.bfloat16) for _ in range(4))
    ref = oracle.attention(q.float().numpy(), k.float().numpy(), v.float().numpy(), causal=True,
                           d_out=do.float().numpy(), round_qk="bf16")
    for rep in range(2):
        for dev in ("cuda:1", "cuda:0"):               # the second device first: nothing was initialised on it
            qd, kd, vd = (t.to(dev).requires_grad_() for t in (q, k, v))
            o = flash_cosine_sim_attention(qd, kd, vd, causal=True)
            o.backward(do.to(dev))
            torch.cuda.synchronize(dev)
            for got, want in zip((o, qd.grad, kd.grad, vd.grad), ref):
                got = got.detach().float().cpu().numpy()
                assert np.abs(got - want).max() / np.abs(want).max() < 2e-2, dev
