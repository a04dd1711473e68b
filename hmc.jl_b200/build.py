"""Builds libhmcgpu.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.  No torch, no JIT cache."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "hmcgpu.cu")
DEPS = [SRC] + [os.path.join(HERE, "csrc", f) for f in ("gibbs_kernel.cuh", "hmm_device.cuh", "rng.cuh")] + [
    os.path.join(HERE, "..", "include", "hmcgpu.h")]
LIB = os.path.join(HERE, "lib", "libhmcgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
         "-Xcompiler", "-fPIC,-O2", "-Xptxas", "-v", "--fmad=true"]


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
        return LIB
    cmd = [NVCC, *FLAGS, "-o", LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    with open(os.path.join(HERE, "lib", "ptxas.log"), "w") as f:
        f.write(res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
