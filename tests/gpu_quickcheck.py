"""Diagnostic sweep (run under gpurun, not collected by pytest): CUDA path vs the float64 oracle on
seeded inputs, printing error metrics per case instead of stopping at the first failure.

    python tests/gpu_quickcheck.py [fwd|all]
"""
import os
import sys
import time
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flash_cosine_sim_attention_b200 import flash_cosine_sim_attention  # noqa: E402
from oracle import cosine_sim_attention_oracle as oracle  # noqa: E402

CASES = [
    # name, q shape, kv shape, dtype, kwargs, mask
    ("square128", (1, 2, 128, 64), (1, 2, 128, 64), torch.bfloat16, dict(), False),
    ("square256_causal", (1, 2, 256, 64), (1, 2, 256, 64), torch.bfloat16, dict(causal=True), False),
    ("ragged_63", (2, 3, 63, 64), (2, 3, 63, 64), torch.float16, dict(), False),
    ("ragged_127_causal", (2, 3, 127, 64), (2, 3, 127, 64), torch.float16, dict(causal=True), False),
    ("n513_causal", (1, 2, 513, 64), (1, 2, 513, 64), torch.bfloat16, dict(causal=True), False),
    ("cross_300_700_mask", (2, 2, 300, 64), (2, 2, 700, 64), torch.bfloat16, dict(), True),
    ("cross_causal_200_456", (1, 2, 200, 64), (1, 2, 456, 64), torch.bfloat16, dict(causal=True), False),
    ("cross_causal_456_200", (1, 2, 456, 64), (1, 2, 200, 64), torch.bfloat16, dict(causal=True), False),
    ("mqa_groups2", (1, 8, 384, 64), (1, 640, 64), torch.bfloat16, dict(groups=2), True),
    ("merged_bh", (6, 260, 64), (6, 260, 64), torch.float16, dict(groups=4, scale=1), False),
    ("n1024_causal", (1, 4, 1024, 64), (1, 4, 1024, 64), torch.bfloat16, dict(causal=True), False),
    ("d128_causal", (1, 2, 384, 128), (1, 2, 384, 128), torch.bfloat16, dict(causal=True), False),
    ("d128_noncausal_f16", (1, 2, 300, 128), (1, 2, 300, 128), torch.float16, dict(), False),
    ("no_l2norm", (1, 2, 200, 64), (1, 2, 200, 64), torch.bfloat16, dict(l2norm_qk=False, scale=1), False),
]


def run(mode):
    dev = torch.device("cuda:0")
    worst = 0.0
    for name, qs, kvs, dtype, kw, use_mask in CASES:
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
        amp = 0.2 if kw.get("l2norm_qk") is False else 1.0
        q = (torch.randn(qs, generator=g) * amp).to(dtype)
        k = (torch.randn(kvs, generator=g) * amp).to(dtype)
        v = torch.randn(kvs, generator=g).to(dtype)
        mask = None
        if use_mask:
            mask = torch.rand((qs[0], kvs[-2]), generator=g) > 0.3
            mask[:, 0] = True
        do = torch.randn(qs, generator=g).to(dtype)
        want_bwd = mode == "all"
        qd, kd, vd = (t.to(dev).requires_grad_(want_bwd) for t in (q, k, v))
        t0 = time.time()
        o = flash_cosine_sim_attention(qd, kd, vd, mask=None if mask is None else mask.to(dev), **kw)
        grads = None
        if want_bwd:
            o.backward(do.to(dev))
            grads = [t.grad.float().cpu().numpy() for t in (qd, kd, vd)]
        torch.cuda.synchronize()
        dt = time.time() - t0
        ref = oracle.attention(q.float().numpy(), k.float().numpy(), v.float().numpy(),
                               mask=None if mask is None else mask.numpy(),
                               d_out=do.float().numpy() if want_bwd else None, empty_rows="zero", **kw)
        refs = ref if want_bwd else (ref,)
        gots = [o.detach().float().cpu().numpy()] + (grads or [])
        line = f"{name:24s} {str(dtype)[6:]:9s}"
        for nm, got, rf in zip(("o", "dq", "dk", "dv"), gots, refs):
            err = np.abs(got - rf).max()
            mag = np.abs(rf).max()
            rel = err / max(mag, 1e-30)
            worst = max(worst, rel)
            bad = "" if (np.isfinite(got).all() and rel < 2e-2) else "  <<<<<< BAD"
            line += f" | {nm} err {err:.3e} / max {mag:.2e} (rel {rel:.1e}){bad}"
        print(line + f"  [{dt*1e3:.0f} ms]", flush=True)
    print(f"worst relative-to-max error: {worst:.3e}")


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "all")
