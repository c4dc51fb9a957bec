"""Per-kernel time of one fwd+bwd step at a given shape, through the public API with the library's event hooks.
    python tests/gpu_breakdown.py [--root DIR] B H N D [causal] [dtype]
--root selects another checkout of the package (A/B against an older build).  Measurement helper only."""
import os
import sys

import torch

args = sys.argv[1:]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if args and args[0] == "--root":
    root = os.path.abspath(args[1])
    args = args[2:]
sys.path.insert(0, root)
from flash_cosine_sim_attention_b200 import _abi, flash_cosine_sim_attention  # noqa: E402

B, H, N, D = (int(x) for x in args[:4])
causal = (args[4] != "0") if len(args) > 4 else True
dt = {"bf16": torch.bfloat16, "f16": torch.float16}[args[5] if len(args) > 5 else "bf16"]
lib = _abi.load()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
q, k, v, do = (torch.randn(B, H, N, D, generator=g, device=dev, dtype=dt) for _ in range(4))
q.requires_grad_(), k.requires_grad_(), v.requires_grad_()


def step():
    o = flash_cosine_sim_attention(q, k, v, causal=causal)
    torch.autograd.grad(o, (q, k, v), do)


for _ in range(3):
    step()
torch.cuda.synchronize()
names = {2: "l2norm_qk", 0: "forward", 3: "preprocess", 1: "backward", 4: "dq_finish"}
K = 8
ev = {w: [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)] for w in names}
tot = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
for w in names:
    for a, b in ev[w]:
        a.record(); b.record()
for i in range(K):
    for w in names:
        lib.fcsa_set_kernel_events(w, ev[w][i][0].cuda_event, ev[w][i][1].cuda_event)
    tot[i][0].record()
    step()
    tot[i][1].record()
torch.cuda.synchronize()
for w in names:
    lib.fcsa_set_kernel_events(w, None, None)
print(f"{root}: (B,H,N,D)=({B},{H},{N},{D}) causal={causal} {dt}")
for w, nm in names.items():
    ms = sorted(a.elapsed_time(b) for a, b in ev[w])[K // 2]
    print(f"   {nm:12s} {ms * 1e3:9.1f} us")
print(f"   {'step':12s} {sorted(a.elapsed_time(b) for a, b in tot)[K // 2] * 1e3:9.1f} us (events between launches)")
