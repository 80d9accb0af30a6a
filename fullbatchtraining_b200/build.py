"""Build the sm_100a CUDA library in-tree (``libfullbatch_b200.so`` next to this file) with plain nvcc.

The library is a C-ABI shared object (include/fullbatch_b200.h); it is loaded with ctypes (lib.py), not as a torch
extension, so no torch headers are involved and the build takes seconds.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libfullbatch_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    srcs = glob.glob(os.path.join(CSRC, "*")) + [os.path.join(os.path.dirname(HERE), "include", "fullbatch_b200.h")]
    return any(os.path.getmtime(s) > t for s in srcs)


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    cmd = [nvcc] + flags + ["-shared", "-o", LIB_PATH] + sources + ["-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libfullbatch_b200.so")
    return LIB_PATH


if __name__ == "__main__":
    build_library(force=True, verbose=True)
    print(LIB_PATH)
