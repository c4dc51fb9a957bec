"""Headline benchmark: fwd+bwd TFLOP/s of causal cosine-sim attention at (B,H,N,D) = (4,8,4096,64),
bf16 - BASELINE.json's metric - on N GPUs of one node (one process per GPU; each rank runs the
same workload: batch x heads shards with no data-path collective, so scaling is "weak").

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this library
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path

One JSON line on stdout (rank 0).  Keys follow the driver contract; in short:
  value        whole-job TFLOP/s, inputs resident in HBM, timed with CUDA events per step
               (an L2 flush runs between steps, outside the events)
  e2e          the same metric through the public API starting from pinned HOST buffers:
               H2D of q,k,v,d_out and D2H of o,dq,dk,dv of every step inside the timed region
               (upload / compute / download on three streams, double-buffered across steps)
  roofline     the dominant kernel (the tcgen05 backward kernel), timed live with events recorded
               around exactly that launch, against the measured bf16 peak (MEASURED_PEAKS.json)
  cpu_baseline the oracle's torch-f32 port of the reference's naive path on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, H, N, D = 4, 8, 4096, 64
SCALE, GROUPS = 8.0, 1
FWD_FLOPS = 4 * B * H * N * N * D / 2            # causal: half the score matrix (SURVEY.md par. 8d)
BWD_FLOPS = 2.5 * FWD_FLOPS                      # 5 GEMMs vs 2
STEP_FLOPS = FWD_FLOPS + BWD_FLOPS               # 2.405e11
METRIC = "fwd+bwd TFLOP/s at (4,8,4096,64) bf16 causal"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16=p.get("bf16_tflops", 1590.0), bf16_sustained=p.get("bf16_tflops_sustained", 1400.0),
                    hbm=p.get("hbm_gbs", 6650.0), source="measured (MEASURED_PEAKS.json)")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for nm, val in zip(names, r[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_sample(torch, heads):
    """One bounded sample of the workload on the host: (1, heads, 4096, 64) f32 causal fwd+bwd through
    the oracle's torch port of the reference's naive path.  Returns (seconds, flops)."""
    from oracle import cosine_sim_attention_oracle as oracle
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(1, heads, N, D, generator=g).requires_grad_() for _ in range(3))
    t0 = time.perf_counter()
    oracle.torch_cpu_forward_backward(q, k, v, scale=SCALE, groups=GROUPS, causal=True)
    dt = time.perf_counter() - t0
    return dt, STEP_FLOPS * heads / (B * H)


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (naive PyTorch ops on the
    host cores; the reference is Python, so the oracle's torch port is what runs)."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    heads = 2                                    # bounded sample: 1/16 of the workload per step
    for _ in range(max(args.warmup, 1)):
        cpu_reference_sample(torch, heads)
    tot_t, tot_f = 0.0, 0.0
    for _ in range(args.steps):
        dt, fl = cpu_reference_sample(torch, heads)
        tot_t += dt
        tot_f += fl
    val = tot_f / tot_t / 1e12
    cores = torch.get_num_threads()
    sample = f"(1,{heads},4096,64) f32 causal fwd+bwd per step = {heads}/{B*H} of the workload; {os.cpu_count()} logical cpus"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3 * (B * H / heads),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "causal self-attn (4,8,4096,64) fwd+bwd, cosine-sim, scale 8", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from flash_cosine_sim_attention_b200 import _abi, debug, flash_cosine_sim_attention

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    lib = _abi.load()
    W = max(args.warmup, 3)
    K = args.steps
    dt = torch.bfloat16

    g = torch.Generator().manual_seed(rank)
    host = [torch.randn(B, H, N, D, generator=g).to(dt).pin_memory() for _ in range(4)]   # q, k, v, d_out
    q, k, v, do = (t.to(dev) for t in host)
    q.requires_grad_(), k.requires_grad_(), v.requires_grad_()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)                 # > 126 MB L2

    def step(qq, kk, vv, dd):
        o = flash_cosine_sim_attention(qq, kk, vv, causal=True, scale=SCALE, groups=GROUPS)
        return (o,) + torch.autograd.grad(o, (qq, kk, vv), dd)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step(q, k, v, do)
    barrier()

    # ---- device-resident timing -------------------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(10)] for _ in range(K)]   # s/e of fwd, bwd, l2norm, prep, finish
    for row in kev:                                                                       # materialise handles
        for e in row:
            e.record()
    sampler = ClockSampler(local)
    launches0 = debug()
    barrier()
    sampler.start()
    # pass 1 - the headline: exactly K steps of the public API, nothing else on the stream between the
    # five kernels of a step (an event recorded between two launches would turn their programmatic
    # dependent launch back into a full serialisation)
    # The host needs ~0.2 ms to enqueue a step and the device ~0.35 ms to run it, so the launch queue
    # runs ahead of the device - except for the very first step after a barrier, whose kernels would be
    # issued into an idle GPU one Python call at a time.  A ~1 ms spin kernel in front of the loop
    # (outside every event pair) lets the queue fill first.
    torch.cuda._sleep(2_000_000)
    for _ in range(2):          # untimed: the device idled during the barrier / sampler start-up (clock ramp)
        flush.zero_()
        step(q, k, v, do)
    for i in range(K):
        flush.zero_()
        ev[i][0].record()
        step(q, k, v, do)
        ev[i][1].record()
    barrier()
    clocks = sampler.stop()
    launches = debug() - launches0 - 2 * 5      # the two untimed steps above
    step_ms = [a.elapsed_time(b) for a, b in ev]
    # pass 2 - the roofline numerators: the same K steps again with events recorded around exactly
    # the forward and the backward tcgen05 kernel (library hook); not part of `value`
    torch.cuda._sleep(2_000_000)
    for i in range(K):
        flush.zero_()
        for w in range(5):
            lib.fcsa_set_kernel_events(w, kev[i][2 * w].cuda_event, kev[i][2 * w + 1].cuda_event)
        step(q, k, v, do)
    barrier()
    for w in range(5):
        lib.fcsa_set_kernel_events(w, None, None)
    fwd_ms = [r[0].elapsed_time(r[1]) for r in kev]
    bwd_ms = [r[2].elapsed_time(r[3]) for r in kev]
    aux_ms = [sum(r[2 * w].elapsed_time(r[2 * w + 1]) for r in kev) / K for w in (2, 3, 4)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)

    # ---- end to end from pinned host buffers ---------------------------------------------------
    # Every step copies its four inputs host->device and its four results device->host; the copies
    # are inside the timed region.  Three streams (upload, compute, download) and two sets of device
    # buffers: the upload of step i+1 and the download of step i-1 run under the compute of step i,
    # the way an input pipeline feeds a training loop.  The compute is the public API call.
    houts = [[torch.empty(B, H, N, D, dtype=dt).pin_memory() for _ in range(4)] for _ in range(2)]
    dins = [[torch.empty(B, H, N, D, dtype=dt, device=dev) for _ in range(4)] for _ in range(2)]
    for bufs in dins:
        for t in bufs[:3]:
            t.requires_grad_()
    s_up, s_cmp, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_up = [torch.cuda.Event() for _ in range(2)]      # inputs of buffer set b are on the device
    ev_cmp = [torch.cuda.Event() for _ in range(2)]     # results of buffer set b are computed (inputs consumed)
    ev_dn = [torch.cuda.Event() for _ in range(2)]      # results of buffer set b are on the host

    def e2e_run(n):
        keep = [None, None]
        for i in range(n):
            bsel = i & 1
            with torch.cuda.stream(s_up):
                s_up.wait_event(ev_cmp[bsel])               # step i-2 has consumed this buffer set
                with torch.no_grad():
                    for dst, src in zip(dins[bsel], host):
                        dst.copy_(src, non_blocking=True)
                ev_up[bsel].record(s_up)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_up[bsel])
                outs = step(*dins[bsel])
                ev_cmp[bsel].record(s_cmp)
            with torch.cuda.stream(s_dn):
                s_dn.wait_event(ev_cmp[bsel])
                s_dn.wait_event(ev_dn[bsel])                # host buffers of step i-2 are written
                for dst, src in zip(houts[bsel], outs):
                    src.record_stream(s_dn)
                    dst.copy_(src.detach(), non_blocking=True)
                ev_dn[bsel].record(s_dn)
            keep[bsel] = outs
        for st in (s_up, s_cmp, s_dn):
            torch.cuda.current_stream().wait_stream(st)

    e2e_run(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in (s_up, s_cmp, s_dn):
        st.wait_stream(torch.cuda.current_stream())
    e2e_run(K)
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)

    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(total_ms.item()), float(e2e_ms.item())

    if rank == 0:
        peaks = load_peaks()
        value = STEP_FLOPS * K * world / (total_ms * 1e-3) / 1e12
        e2e_val = STEP_FLOPS * K * world / (e2e_ms * 1e-3) / 1e12
        bwd_avg = sum(bwd_ms) / K
        fwd_avg = sum(fwd_ms) / K
        ach = BWD_FLOPS / (bwd_avg * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("fcsa_bwd_kernel_dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "causal self-attn (B,H,N,D)=(4,8,4096,64) per GPU, cosine-sim (l2norm fused), "
                                   "scale 8, fwd+bwd through flash_cosine_sim_attention()",
                       "global_batch": B * world, "seq_len": N, "heads": H, "head_dim": D,
                       "parallelism": f"batch-shard x{world} (no data-path collective)",
                       "l2": "256 MiB L2 flush between timed steps (outside the events)",
                       "timing": "W warm-up steps, barrier+sync, 2 more untimed steps (clock ramp, launch queue), K steps "
                                 "each between its own CUDA events, barrier+sync; max over ranks",
                       "flops_per_step_per_gpu": STEP_FLOPS},
            "frac_of_peak": value / world / peaks["bf16"], "peak_source": peaks["source"],
            "roofline": {"bound": "tensor", "kernel": "fcsa_bwd_kernel<bf16,64>", "achieved": ach,
                         "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": ach / peaks["bf16"], "traffic": traffic,
                         "ms": bwd_avg, "flops_per_launch": BWD_FLOPS},
            "roofline_fwd": {"bound": "tensor", "kernel": "fcsa_fwd_kernel<bf16,64>",
                             "achieved": FWD_FLOPS / (fwd_avg * 1e-3) / 1e12, "peak": peaks["bf16"],
                             "unit": "TFLOP/s", "frac": FWD_FLOPS / (fwd_avg * 1e-3) / 1e12 / peaks["bf16"],
                             "ms": fwd_avg, "flops_per_launch": FWD_FLOPS},
            "e2e": {"value": e2e_val, "unit": "TFLOP/s", "h2d_bytes_per_step": 4 * B * H * N * D * 2 * world,
                    "d2h_bytes_per_step": 4 * B * H * N * D * 2 * world, "ms_per_step": e2e_ms / K},
            "step_breakdown_ms": {"l2norm_qk": aux_ms[0], "forward": fwd_avg, "preprocess": aux_ms[1],
                                  "backward": bwd_avg, "dq_finish": aux_ms[2],
                                  "note": "instrumented second pass (events between the launches); the "
                                          "headline pass has none"},
            "gpu_launches": int(launches), "clocks": clocks,
            "ms_per_step_min_median_max": [min(step_ms), sorted(step_ms)[len(step_ms) // 2], max(step_ms)],
            "slowest_step_index": int(max(range(len(step_ms)), key=lambda j: step_ms[j])),
        }
        if world == 1 and not args.no_cpu_baseline:
            cpu_reference_sample(torch, 1)                        # warm-up
            t, f = 0.0, 0.0
            for _ in range(3):
                a, b = cpu_reference_sample(torch, 2)
                t, f = t + a, f + b
            line["cpu_baseline"] = {"value": f / t / 1e12, "unit": "TFLOP/s", "cores": torch.get_num_threads(),
                                    "kind": "port",
                                    "sample": "3 x (1,2,4096,64) f32 causal fwd+bwd (2/32 of the workload each), "
                                              f"oracle torch port of the naive reference path, {os.cpu_count()} logical cpus"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
