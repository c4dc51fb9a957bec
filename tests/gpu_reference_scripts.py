"""Runs the REFERENCE'S OWN scripts, unmodified, against this repository's operator on the GPU box
(verdict r1 item 7: "prove the drop-in with the reference's own scripts"):

  1. tests/test.py            the reference's pytest grid (its CPU-path test function is excluded: this
                              build has no CPU path by design)
  2. benchmark.py [--causal]  the reference's timing sweep (fused op vs its naive baseline, f32 and f16)
  3. train.py --use-cuda-kernel   20 optimizer steps of the enwik8 recipe on a synthetic stand-in corpus
  4. the reference's own CUDA kernel (its .cu compiled for sm_100a by oracle/stage_reference.py --cuda)
     timed at the metric shape next to this repository's kernels: the wmma "kernel to beat"

The scripts come from baseline/_ref/ (staged by oracle/stage_reference.py, git-ignored, byte-identical copies)
or /root/reference when present; `import flash_cosine_sim_attention` resolves to this repository's drop-in
package because the repository root is first on PYTHONPATH.  Logs go to --out (default gpurun_out/).
TEST / MEASUREMENT INFRASTRUCTURE ONLY.
"""
import argparse
import gzip
import importlib.util
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.exists(os.path.join(cand, "tests", "test.py")):
            return cand
    return None


def env():
    e = dict(os.environ)
    e["PYTHONPATH"] = ROOT + os.pathsep + e.get("PYTHONPATH", "")
    return e


def run_tests(ref, out):
    log = os.path.join(out, "ref_tests.log")
    cmd = [sys.executable, "-m", "pytest", os.path.join(ref, "tests", "test.py"), "-q", "-p", "no:cacheprovider",
           "-k", "not cpu", "--tb=line", "-rf"]
    t0 = time.time()
    p = subprocess.run(cmd, cwd=out, env=env(), capture_output=True, text=True, timeout=1500)
    text = p.stdout + p.stderr
    # per-dtype tally from the parametrised ids (float16 flag is the 5th-from-last parameter: True/False)
    failed = re.findall(r"FAILED .*?::(test_\w+)\[(.*?)\]", text)
    with open(log, "w") as f:
        f.write("$ " + " ".join(cmd) + "\n" + text[-60000:])
        f.write(f"\n[wall {time.time() - t0:.1f} s]  failed cases: {len(failed)}\n")
    tail = [ln for ln in text.splitlines() if re.search(r"\d+ (passed|failed)", ln)]
    print("reference tests/test.py:", tail[-1] if tail else "no summary", flush=True)
    by = {}
    for name, ids in failed:
        parts = ids.split("-")
        key = (name, "f16" if parts[2] == "True" else "f32")          # ids: single_head_kv-bias_batch_dim-float16-dim_head-...
        by[key] = by.get(key, 0) + 1
    for key in sorted(by):
        print("   failed", key, by[key], flush=True)


def run_benchmark(ref, out):
    for flags in (["--causal", "--num-times", "10"], ["--num-times", "10"]):
        log = os.path.join(out, "ref_benchmark" + ("_causal" if "--causal" in flags else "") + ".log")
        cmd = [sys.executable, os.path.join(ref, "benchmark.py")] + flags
        p = subprocess.run(cmd, cwd=out, env=env(), capture_output=True, text=True, timeout=900)
        with open(log, "w") as f:
            f.write("$ " + " ".join(cmd) + "\n" + p.stdout + p.stderr[-5000:])
        print(f"reference benchmark.py {' '.join(flags)}: rc={p.returncode}", flush=True)
        for ln in p.stdout.splitlines():
            if "4096" in ln or "8192" in ln:
                print("   ", ln, flush=True)


def run_train(ref, out, steps=20):
    work = tempfile.mkdtemp(prefix="fcsa_train_")
    os.makedirs(os.path.join(work, "data"))
    # synthetic enwik8 stand-in: 95 MB of byte "text" (the real corpus is not shipped; there is no network)
    import numpy as np
    rng = np.random.default_rng(0)
    words = [bytes(rng.integers(97, 123, size=int(n)).tolist()) for n in rng.integers(2, 9, size=4096)]
    idx = rng.integers(0, len(words), size=17_000_000)
    blob = b" ".join(words[i] for i in idx)[: int(95e6)]
    assert len(blob) == int(95e6)
    with gzip.open(os.path.join(work, "data", "enwik8.gz"), "wb", compresslevel=1) as f:
        f.write(blob)
    log = os.path.join(out, "ref_train.log")
    cmd = [sys.executable, "-u", os.path.join(ref, "train.py"), "--use-cuda-kernel", "--seq-len", "1024"]
    t0 = time.time()
    p = subprocess.Popen(cmd, cwd=work, env=env(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    lines, losses = [], []
    try:
        for ln in p.stdout:
            lines.append(ln)
            m = re.match(r"training loss: ([\d.eE+-]+|nan|inf)", ln)
            if m:
                losses.append(float(m.group(1)))
                if len(losses) >= steps:
                    break
            if time.time() - t0 > 900:
                break
    finally:
        p.kill()                      # our own child, by handle
        p.wait()
    with open(log, "w") as f:
        f.write("$ " + " ".join(cmd) + f"   (cwd: synthetic corpus, stopped after {steps} optimizer steps)\n")
        f.write("".join(lines)[-40000:])
        f.write(f"\n[wall {time.time() - t0:.1f} s]  training losses: {losses}\n")
    ok = len(losses) >= steps and all(x == x and x < 1e4 for x in losses)
    print(f"reference train.py --use-cuda-kernel: {len(losses)} steps, loss {losses[0] if losses else None} -> "
          f"{losses[-1] if losses else None}, finite={ok}", flush=True)


def time_reference_cuda_kernel(out):
    """The reference's own .cu (wmma / scalar-FMA kernels, compiled for sm_100a) at (4,8,4096,64) f16 causal,
    next to this repository's kernels - same tensors, CUDA events, best of 10."""
    import torch
    cand = [f for f in os.listdir(os.path.join(ROOT, "oracle", "_ref"))
            if f.startswith("flash_cosine_sim_attention_cuda_ref") and f.endswith(".so")] if os.path.isdir(
        os.path.join(ROOT, "oracle", "_ref")) else []
    log = os.path.join(out, "ref_cuda_kernel.log")
    if not cand:
        open(log, "w").write("oracle/_ref/flash_cosine_sim_attention_cuda_ref*.so not present (run oracle/stage_reference.py --cuda)\n")
        print("reference CUDA kernel: not built", flush=True)
        return
    spec = importlib.util.spec_from_file_location("flash_cosine_sim_attention_cuda_ref",
                                                  os.path.join(ROOT, "oracle", "_ref", cand[0]))
    refk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refk)
    sys.path.insert(0, ROOT)
    import flash_cosine_sim_attention_b200 as ours
    from flash_cosine_sim_attention_b200.flash_cosine_sim_attention import backward as our_bwd, forward as our_fwd
    B, H, N, D = 4, 8, 4096, 64
    dt = torch.float16
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v, do = (torch.randn(B, H, N, D, generator=g, device="cuda", dtype=dt) for _ in range(4))
    qn, kn = ours.l2norm_tensors(q, k)
    flops_f, flops_b = 4 * B * H * N * N * D / 2, 10 * B * H * N * N * D / 2

    def best(fn, n=10):
        ts = []
        for _ in range(n + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return min(ts[2:]), r
    rows = []
    t, (o_r, l_r, _) = best(lambda: refk.forward(qn, kn, v, None, None, False, 8.0, True))
    rows.append(("reference forward_kernel (wmma)", t, flops_f))
    t2, grads_r = best(lambda: refk.backward(do, o_r, l_r, qn, kn, v, None, None, False, 8.0, True))
    rows.append(("reference backward (preprocess + backward_kernel + casts)", t2, flops_b))
    t3, (o_o, l_o, _) = best(lambda: our_fwd(qn, kn, v, None, None, False, 8.0, True))
    rows.append(("this repo fcsa_fwd_kernel (tcgen05)", t3, flops_f))
    t4, grads_o = best(lambda: our_bwd(do, o_o, l_o, qn, kn, v, None, None, False, 8.0, True))
    rows.append(("this repo backward (prep + fcsa_bwd_kernel + dq conversion)", t4, flops_b))
    err_o = float((o_r.float() - o_o.float()).abs().max())
    err_g = [float((a.float() - b.float()).abs().max() / b.float().abs().max()) for a, b in zip(grads_r[:3], grads_o[:3])]
    with open(log, "w") as f:
        f.write("(B,H,N,D) = (4,8,4096,64) float16 causal, already-normalised q, k; CUDA events, best of 10, one B200\n")
        for name, ms, fl in rows:
            line = f"{name:62s} {ms * 1e3:9.1f} us  {fl / (ms * 1e-3) / 1e12:8.1f} TFLOP/s"
            f.write(line + "\n")
            print("   " + line, flush=True)
        f.write(f"speed-up forward {rows[0][1] / rows[2][1]:.1f}x, backward {rows[1][1] / rows[3][1]:.1f}x\n")
        f.write(f"max |o_ref - o_ours| = {err_o:.3e}; relative-to-max differences of dq, dk, dv = {err_g}\n")
    print(f"reference CUDA kernel vs ours: fwd {rows[0][1] / rows[2][1]:.1f}x, bwd {rows[1][1] / rows[3][1]:.1f}x; "
          f"max|do| {err_o:.2e}, grads rel {['%.2e' % x for x in err_g]}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    ap.add_argument("--only", default="tests,benchmark,train,refkernel")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    ref = find_reference()
    if ref is None:
        print("no reference copy found (run oracle/stage_reference.py in the build container)")
        return 1
    print("reference scripts from", ref, flush=True)
    only = args.only.split(",")
    if "tests" in only:
        run_tests(ref, args.out)
    if "benchmark" in only:
        run_benchmark(ref, args.out)
    if "train" in only:
        run_train(ref, args.out)
    if "refkernel" in only:
        time_reference_cuda_kernel(args.out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
