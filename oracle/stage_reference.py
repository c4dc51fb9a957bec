"""Stages the UNMODIFIED reference next to the repo for the GPU box, which has no /root/reference.

    python oracle/stage_reference.py            # python files only (seconds)
    python oracle/stage_reference.py --cuda     # + the reference's own CUDA extension built for sm_100a

Nothing staged here is tracked by git (`baseline/_ref/` and `oracle/_ref/` are git-ignored; both travel
with gpurun).  Reference sources are never copied into the tracked tree.

  baseline/_ref/                the reference's python package and its own scripts, byte for byte:
      flash_cosine_sim_attention/{__init__,flash_cosine_sim_attention,transformer,benchmark,version}.py
      scripts/{tests/test.py, benchmark.py, train.py}
    used by (a) `bench.py --impl reference`, which times the reference's own plain_cosine_sim_attention on
    the host cores, and (b) tests/gpu_reference_scripts.py, which runs the reference's test-suite,
    benchmark.py and train.py against THIS repo's operator (verdict r1 item 7).

  oracle/_ref/flash_cosine_sim_attention_cuda_ref*.so   (--cuda) the reference's .cu compiled from the
    sources where they lie (/root/reference/flash_cosine_sim_attention/flash_cosine_sim_attention_cuda.cu)
    for sm_100a: the wmma kernel-to-beat, timed on the same B200 by tests/gpu_reference_scripts.py.
    It needs -DC10_UNUSED_DISPATCH_CUDA_WORKAROUND= (the macro left ATen/Dispatch.h in torch 2.11,
    SURVEY.md par. 7.2); no reference file is modified.
TEST / MEASUREMENT INFRASTRUCTURE ONLY - the product never imports anything from these directories.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
CUDA_DST = os.path.join(HERE, "_ref")
CUDA_NAME = "flash_cosine_sim_attention_cuda_ref"

# source (relative to the reference) -> destination (relative to baseline/_ref).  The scripts go into their own
# directory: python puts a script's directory first on sys.path, and next to the reference's package directory
# `import flash_cosine_sim_attention` would pick the REFERENCE package (which cannot even be imported without
# its compiled extension) instead of the drop-in one under test.
FILES = {
    "flash_cosine_sim_attention/__init__.py": "flash_cosine_sim_attention/__init__.py",
    "flash_cosine_sim_attention/flash_cosine_sim_attention.py": "flash_cosine_sim_attention/flash_cosine_sim_attention.py",
    "flash_cosine_sim_attention/transformer.py": "flash_cosine_sim_attention/transformer.py",
    "flash_cosine_sim_attention/benchmark.py": "flash_cosine_sim_attention/benchmark.py",
    "flash_cosine_sim_attention/version.py": "flash_cosine_sim_attention/version.py",
    "tests/test.py": "scripts/tests/test.py",
    "benchmark.py": "scripts/benchmark.py",
    "train.py": "scripts/train.py",
}


def stage_python():
    if not os.path.isdir(REF):
        return False
    for rel, out in FILES.items():
        src, dst = os.path.join(REF, rel), os.path.join(DST, out)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copyfile(src, dst)
    return True


def cuda_ext_path():
    return os.path.join(CUDA_DST, CUDA_NAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build_reference_cuda(force=False):
    """nvcc on the reference's single .cu, in place; output only into oracle/_ref/."""
    src = os.path.join(REF, "flash_cosine_sim_attention", "flash_cosine_sim_attention_cuda.cu")
    out = cuda_ext_path()
    if not os.path.exists(src):
        return None
    if os.path.exists(out) and not force:
        return out
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(CUDA_DST, exist_ok=True)
    inc = list(ce.include_paths()) + [sysconfig.get_paths()["include"]]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--shared",
           "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "--expt-extended-lambda",
           "-DC10_UNUSED_DISPATCH_CUDA_WORKAROUND=", f"-DTORCH_EXTENSION_NAME={CUDA_NAME}",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_BFLOAT16_CONVERSIONS__",
           "-D__CUDA_NO_HALF2_OPERATORS__"]
    cmd += [f"-I{d}" for d in inc] + [src, "-o", out]
    cmd += [f"-L{d}" for d in ce.library_paths()] + ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
                                                     "-ltorch_python"]
    cmd += [f"-Xlinker=-rpath={d}" for d in ce.library_paths()]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(CUDA_DST, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("reference .cu did not compile for sm_100a:\n" + (proc.stdout + proc.stderr)[-4000:])
    return out


if __name__ == "__main__":
    print("python staged:", stage_python(), "->", DST)
    if "--cuda" in sys.argv:
        print("reference CUDA extension:", build_reference_cuda(force="-f" in sys.argv))
