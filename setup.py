"""Build / install script - the counterpart of the reference's setup.py (setup.py:30-39: one CUDAExtension
named by version.py:3, no arch flags).  Here:

  * libfcsa_b200.so                          nvcc, ONE target: -gencode arch=compute_100a,code=sm_100a
    (the C-ABI library, include/fcsa_b200.h; installed inside the package)
  * flash_cosine_sim_attention_cuda_0_1_40   the PyTorch extension module over that ABI, under the very name
    the reference's flash_cosine_sim_attention.py imports (py:15-20) - installed TOP-LEVEL like the reference's

    python setup.py build_ext --inplace        # what __graft_entry__.build() / build.py do in-tree
    pip install --no-build-isolation .         # site-packages install (sm_100a only; needs nvcc and g++)

Both steps are implemented once, in flash_cosine_sim_attention_b200/build.py.
"""
import importlib.util
import os
import shutil
import sys

from setuptools import Extension, find_packages, setup
from setuptools.command.build_ext import build_ext

ROOT = os.path.dirname(os.path.abspath(__file__))


def _build_module():
    spec = importlib.util.spec_from_file_location(
        "fcsa_b200_build", os.path.join(ROOT, "flash_cosine_sim_attention_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class BuildNative(build_ext):
    """Runs build.py (nvcc for the library, the host compiler for the torch module) and drops the artefacts
    where setuptools expects them: the library inside the package, the module top-level."""

    def run(self):
        b = _build_module()
        lib = b.build_library(force=False)
        ext = b.build_extension(force=False)
        if self.inplace:
            return                                           # already in-tree, next to the sources
        pkg_dir = os.path.join(self.build_lib, "flash_cosine_sim_attention_b200")
        os.makedirs(pkg_dir, exist_ok=True)
        shutil.copy2(lib, os.path.join(pkg_dir, os.path.basename(lib)))
        shutil.copy2(ext, os.path.join(self.build_lib, os.path.basename(ext)))     # top-level, like the reference


exec(open(os.path.join(ROOT, "flash_cosine_sim_attention_b200", "version.py")).read())

setup(
    name="flash-cosine-sim-attention-b200",
    version=__version__,                                      # noqa: F821  (from version.py)
    description="B200 (sm_100a) native fused cosine-similarity attention - drop-in for flash-cosine-sim-attention",
    packages=find_packages(include=["flash_cosine_sim_attention_b200", "flash_cosine_sim_attention"]),
    package_data={"flash_cosine_sim_attention_b200": ["csrc/*", "*.so"]},
    data_files=[("include", ["include/fcsa_b200.h"])],
    ext_modules=[Extension(__cuda_pkg_name__, sources=[])],  # noqa: F821  built by BuildNative, not by distutils
    cmdclass={"build_ext": BuildNative},
    python_requires=">=3.9",
    install_requires=["torch>=2.4"],
    zip_safe=False,
)
