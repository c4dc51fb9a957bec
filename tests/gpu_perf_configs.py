"""Timing of the other BASELINE.json configs (they are parity cases, not the bench metric): prints
ms and TFLOP/s per config through the public API.  Run under gpurun:  python tests/gpu_perf_configs.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flash_cosine_sim_attention_b200 import flash_cosine_sim_attention  # noqa: E402

CONFIGS = [
    # name, q shape, kv shape, dtype, kwargs, backward
    ("C2 self-attn (1,8,1024,64) bf16 fwd", (1, 8, 1024, 64), (1, 8, 1024, 64), torch.bfloat16, dict(), False),
    ("C3 causal (4,8,4096,64) f16 fwd+bwd", (4, 8, 4096, 64), (4, 8, 4096, 64), torch.float16, dict(causal=True), True),
    ("C3 causal (4,8,4096,64) bf16 fwd+bwd", (4, 8, 4096, 64), (4, 8, 4096, 64), torch.bfloat16, dict(causal=True), True),
    ("C3 non-causal (4,8,4096,64) bf16 fwd+bwd", (4, 8, 4096, 64), (4, 8, 4096, 64), torch.bfloat16, dict(), True),
    ("C4 cross+mask MQA groups2 (1,8,1024->2048,64) fwd+bwd", (1, 8, 1024, 64), (1, 2048, 64), torch.bfloat16, dict(groups=2), True),
    ("C5 per-GPU share (1,16,16384,128) bf16 causal fwd+bwd", (1, 16, 16384, 128), (1, 16, 16384, 128), torch.bfloat16, dict(causal=True), True),
    ("(4,8,8192,64) bf16 causal fwd+bwd", (4, 8, 8192, 64), (4, 8, 8192, 64), torch.bfloat16, dict(causal=True), True),
    ("(4,8,4096,128) bf16 causal fwd+bwd", (4, 8, 4096, 128), (4, 8, 4096, 128), torch.bfloat16, dict(causal=True), True),
    ("C3 + attn_bias (8,4096,4096) no bias grad, bf16 causal fwd+bwd", (4, 8, 4096, 64), (4, 8, 4096, 64), torch.bfloat16, dict(causal=True, bias="nograd"), True),
    ("C3 + attn_bias (8,4096,4096) with d_bias, bf16 causal fwd+bwd", (4, 8, 4096, 64), (4, 8, 4096, 64), torch.bfloat16, dict(causal=True, bias="grad"), True),
    ("C3 float32 inputs (fp16 kernels), causal fwd+bwd", (4, 8, 4096, 64), (4, 8, 4096, 64), torch.float32, dict(causal=True), True),
]


def main():
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, qs, kvs, dt, kw, bwd in CONFIGS:
        g = torch.Generator().manual_seed(0)
        q = torch.randn(qs, generator=g).to(dt).to(dev).requires_grad_(bwd)
        k = torch.randn(kvs, generator=g).to(dt).to(dev).requires_grad_(bwd)
        v = torch.randn(kvs, generator=g).to(dt).to(dev).requires_grad_(bwd)
        do = torch.randn(qs, generator=g).to(dt).to(dev)
        kw = dict(kw)
        bias_mode = kw.pop("bias", None)
        bias = None
        if bias_mode:
            bias = (torch.randn(qs[1], qs[2], kvs[-2], generator=g) * 0.5).to(dt).to(dev).requires_grad_(bias_mode == "grad")
            kw["attn_bias"] = bias
        mask = None
        if "mask" in name:
            mask = torch.rand(qs[0], kvs[-2], generator=g).to(dev) > 0.25

        def step():
            o = flash_cosine_sim_attention(q, k, v, mask=mask, **kw)
            if bwd:
                torch.autograd.grad(o, (q, k, v) + ((bias,) if bias is not None and bias.requires_grad else ()), do)

        for _ in range(3):
            step()
        ms = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        t = sum(ms) / len(ms)
        B, H, Nq, D = qs if len(qs) == 4 else (qs[0], 1, qs[1], qs[2])
        Nk = kvs[-2]
        flops = 4.0 * B * H * Nq * Nk * D * (0.5 if kw.get("causal") else 1.0) * (3.5 if bwd else 1.0)
        print(f"{name:60s} {t:8.3f} ms  {flops / (t * 1e-3) / 1e12:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
