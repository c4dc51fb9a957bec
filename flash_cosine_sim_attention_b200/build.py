"""Builds libfcsa_b200.so (the C-ABI CUDA library) in-tree with nvcc, for sm_100a only.

Replaces the reference's setup.py CUDAExtension (setup.py:30-39), which passed no arch
flags at all.  There is deliberately a single `-gencode`: this library has no other target.

    python -m flash_cosine_sim_attention_b200.build        # build (skips if up to date)
    python -m flash_cosine_sim_attention_b200.build -f     # force rebuild
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libfcsa_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libfcsa_b200.so cannot be built")


def _sources():
    return [os.path.join(CSRC, "fcsa_abi.cu")]


def _deps():
    out = [os.path.join(ROOT, "include", "fcsa_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        out.append(os.path.join(CSRC, f))
    return out


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in _deps())


def build_library(force=False, verbose=False):
    """Compile the library if needed; returns its path."""
    if not force and up_to_date():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB] + _sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-8000:])
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    path = build_library(force="-f" in sys.argv, verbose="-v" in sys.argv)
    print(path)
