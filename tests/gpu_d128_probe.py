"""Forward-only and forward+backward time of the two D = 128 shapes (C5 per-GPU share, (4,8,4096,128)) through the
public API - the probe that separated the D = 128 forward regression (MMA-issuer spills) from the backward's.
FCSA_OLD_PKG=1 imports a package copy from _scratch_r1/old_pkg (a build of an earlier commit staged by hand) for a
same-box comparison.  Run under gpurun:  python tests/gpu_d128_probe.py"""
import os, sys, torch
sys.path.insert(0, "_scratch_r1/old_pkg" if os.environ.get("FCSA_OLD_PKG") else ".")
from flash_cosine_sim_attention_b200 import flash_cosine_sim_attention
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for qs in [(1, 16, 16384, 128), (4, 8, 4096, 128)]:
    g = torch.Generator().manual_seed(0)
    q, k, v, do = (torch.randn(qs, generator=g).to(torch.bfloat16).to(dev) for _ in range(4))
    for bwd in (False, True):
        qq, kk, vv = (t.clone().requires_grad_(bwd) for t in (q, k, v))
        def step():
            o = flash_cosine_sim_attention(qq, kk, vv, causal=True)
            if bwd: torch.autograd.grad(o, (qq, kk, vv), do)
        for _ in range(3): step()
        ms = []
        for _ in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
        print("old-pkg" if os.environ.get("FCSA_OLD_PKG") else "current", qs, "fwd+bwd" if bwd else "fwd", "%.3f ms" % sorted(ms)[len(ms)//2], flush=True)
