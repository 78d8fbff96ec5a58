"""Builds feature_intertwiner_b200/lib/libfi_b200.so from csrc/*.cu with nvcc for sm_100a.

In-tree on purpose: the .so is git-ignored but travels with the repo snapshot to the GPU box.
`python -m feature_intertwiner_b200.build` or `build_library()`; nvcc cross-compiles without a GPU.
"""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
# FI_LIB_NAME / FI_NVCC_EXTRA: experiment builds (e.g. FI_NVCC_EXTRA="-DFI_TILE_PAIR=0" FI_LIB_NAME=libfi_b200_nopair.so)
LIB_PATH = os.path.join(LIB_DIR, os.environ.get("FI_LIB_NAME", "libfi_b200.so"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--shared", "-cudart", "static",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(PKG, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("FI_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + _sources()
    env = dict(os.environ)
    # the image exports CC/CXX wrappers; nvcc wants the distro host compiler
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libfi_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
