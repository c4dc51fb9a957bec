"""Runs the REFERENCE'S OWN scripts, unmodified, against this repository's operator on the GPU box
(verdict r1 item 7: "prove the drop-in with the reference's own scripts"):

  1. tests/test.py            the reference's pytest grid (its CPU-path test function is excluded: this
                              build has no CPU path by design)
  2. benchmark.py [--causal]  the reference's timing sweep (fused op vs its naive baseline, f32 and f16)
  3. train.py --use-cuda-kernel   20 optimizer steps of the enwik8 recipe on a synthetic stand-in corpus
  4. the reference's own CUDA kernel (its .cu compiled for sm_100a by oracle/stage_reference.py --cuda)
     timed at the metric shape next to this repository's kernels: the wmma "kernel to beat"

The scripts come from baseline/_ref/scripts/ (staged by oracle/stage_reference.py, git-ignored, byte-identical
copies); `import flash_cosine_sim_attention` resolves to this repository's drop-in package because the repository
root is on PYTHONPATH and the reference's own package does not sit next to the scripts.  Logs go to --out (default gpurun_out/).
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
    """Directory holding the reference's scripts WITHOUT its package next to them (a script's own directory comes
    first on sys.path; the drop-in `flash_cosine_sim_attention` package of this repository must win)."""
    cand = os.path.join(ROOT, "baseline", "_ref", "scripts")
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
    # The reference compares float32 at atol = 1e-4 (tests/test.py:33,81).  float32 inputs run here at tf32-class
    # operand precision (11-bit significands, fp32 accumulation), which north_star budgets at 1e-3: the same
    # grid once more with ONLY the tolerance of the f32 rows changed (a conftest that patches `allclose`).
    conf = os.path.join(out, "fcsa_f32_tol_plugin.py")
    with open(conf, "w") as f:
        f.write("import pytest\n\n@pytest.fixture(autouse=True)\ndef _f32_tolerance_1e3(request, monkeypatch):\n"
                "    mod = request.module\n    orig = mod.allclose\n"
                "    monkeypatch.setattr(mod, 'allclose', lambda a, b, atol=1e-4: orig(a, b, atol=max(atol, 1e-3)))\n")
    cmd2 = [sys.executable, "-m", "pytest", os.path.join(ref, "tests", "test.py"), "-q", "-p", "no:cacheprovider",
            "-k", "not cpu", "--tb=line", "-p", "fcsa_f32_tol_plugin"]
    e2 = env()
    e2["PYTHONPATH"] = out + os.pathsep + e2["PYTHONPATH"]
    p2 = subprocess.run(cmd2, cwd=out, env=e2, capture_output=True, text=True, timeout=1500)
    os.remove(conf)
    text2 = p2.stdout + p2.stderr
    tail2 = [ln for ln in text2.splitlines() if re.search(r"\d+ (passed|failed)", ln)]
    with open(log, "a") as f:
        f.write("\n$ same grid, float32 rows compared at atol 1e-3 instead of 1e-4 (only change: tolerance)\n" + text2[-8000:])
    print("   with f32 atol 1e-3:", tail2[-1] if tail2 else "no summary", flush=True)


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


REFK_CHILD = r"""
import importlib.util, json, os, sys, torch
root, so, what = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, root)
spec = importlib.util.spec_from_file_location("flash_cosine_sim_attention_cuda_ref", so)
refk = importlib.util.module_from_spec(spec); spec.loader.exec_module(refk)
import flash_cosine_sim_attention_b200 as ours
from flash_cosine_sim_attention_b200.flash_cosine_sim_attention import backward as our_bwd, forward as our_fwd
B, H, N, D = 4, 8, 4096, 64
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v, do = (torch.randn(B, H, N, D, generator=g, device="cuda", dtype=torch.float16) for _ in range(4))
qn, kn = ours.l2norm_tensors(q, k)
def best(fn, n=10):
    ts = []
    for _ in range(n + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[2:]), r
res = {}
t, (o_o, l_o, _) = best(lambda: our_fwd(qn, kn, v, None, None, False, 8.0, True)); res["ours_fwd_ms"] = t
t, g_o = best(lambda: our_bwd(do, o_o, l_o, qn, kn, v, None, None, False, 8.0, True)); res["ours_bwd_ms"] = t
t, (o_r, l_r, _) = best(lambda: refk.forward(qn, kn, v, None, None, False, 8.0, True)); res["ref_fwd_ms"] = t
res["max_abs_o_diff"] = float((o_r.float() - o_o.float()).abs().max())
print("RESULT " + json.dumps(res), flush=True)
if what == "bwd":
    t, g_r = best(lambda: refk.backward(do, o_r, l_r, qn, kn, v, None, None, False, 8.0, True)); res["ref_bwd_ms"] = t
    res["grad_rel_diff"] = [float((a.float() - b.float()).abs().max() / b.float().abs().max()) for a, b in zip(g_r[:3], g_o[:3])]
    print("RESULT " + json.dumps(res), flush=True)
"""


def time_reference_cuda_kernel(out):
    """The reference's own .cu (wmma / scalar-FMA kernels, compiled for sm_100a) at (4,8,4096,64) f16 causal,
    next to this repository's kernels - same tensors, CUDA events, best of 10.  Run in a child process: a fault
    inside the reference's kernel must not take the runner down."""
    import json
    d = os.path.join(ROOT, "oracle", "_ref")
    cand = [f for f in os.listdir(d) if f.startswith("flash_cosine_sim_attention_cuda_ref") and f.endswith(".so")] \
        if os.path.isdir(d) else []
    log = os.path.join(out, "ref_cuda_kernel.log")
    if not cand:
        open(log, "w").write("oracle/_ref/flash_cosine_sim_attention_cuda_ref*.so not present (run oracle/stage_reference.py --cuda)\n")
        print("reference CUDA kernel: not built", flush=True)
        return
    res, notes = {}, []
    for what in ("bwd", "fwd"):
        p = subprocess.run([sys.executable, "-c", REFK_CHILD, ROOT, os.path.join(d, cand[0]), what], capture_output=True,
                           text=True, timeout=600)
        lines = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
        if lines:
            res.update(json.loads(lines[-1][7:]))
        if p.returncode != 0:
            err = [ln for ln in (p.stdout + p.stderr).splitlines() if "rror" in ln]
            notes.append(f"child '{what}' exited with {p.returncode}: " + (err[0][:200] if err else "no message"))
        if "ref_bwd_ms" in res or what == "fwd":
            break
    ff, fb = 4 * 4 * 8 * 4096 * 4096 * 64 / 2, 10 * 4 * 8 * 4096 * 4096 * 64 / 2
    with open(log, "w") as f:
        f.write("(B,H,N,D) = (4,8,4096,64) float16 causal, already-normalised q, k; CUDA events, best of 10, one B200\n")
        for name, key, fl in (("reference forward_kernel (wmma, compiled for sm_100a)", "ref_fwd_ms", ff),
                              ("reference backward (preprocess + backward_kernel + casts)", "ref_bwd_ms", fb),
                              ("this repo forward (fcsa_fwd_kernel, tcgen05)", "ours_fwd_ms", ff),
                              ("this repo backward (prep + fcsa_bwd_kernel + dq conversion)", "ours_bwd_ms", fb)):
            if key in res:
                line = f"{name:62s} {res[key] * 1e3:9.1f} us  {fl / (res[key] * 1e-3) / 1e12:8.1f} TFLOP/s"
            else:
                line = f"{name:62s}   did not complete"
            f.write(line + "\n")
            print("   " + line, flush=True)
        if "ref_fwd_ms" in res:
            f.write(f"forward speed-up {res['ref_fwd_ms'] / res['ours_fwd_ms']:.1f}x; max |o_ref - o_ours| = {res.get('max_abs_o_diff')}\n")
        if "ref_bwd_ms" in res:
            f.write(f"backward speed-up {res['ref_bwd_ms'] / res['ours_bwd_ms']:.1f}x; relative-to-max differences of dq, dk, dv = {res.get('grad_rel_diff')}\n")
        for n in notes:
            f.write(n + "\n")
            print("   note:", n, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    ap.add_argument("--only", default="tests,benchmark,train,refkernel")
    args = ap.parse_args()
    args.out = os.path.abspath(args.out)
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
