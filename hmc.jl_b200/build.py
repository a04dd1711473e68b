"""Builds libhmcgpu.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.  No torch, no JIT cache.

The Gibbs sweep kernel family is instantiated once per (precision, K) in its own translation unit
(csrc/gibbs_inst.cu with -DHMC_R/-DHMC_K); the units are compiled in parallel and linked with csrc/hmcgpu.cu."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HEADERS = [os.path.join(CSRC, f) for f in ("gibbs_kernel.cuh", "gibbs_pair_kernel.cuh", "gibbs_wide_kernel.cuh", "gibbs_scan_kernel.cuh", "hmm_device.cuh", "rng.cuh")] + [
    os.path.join(HERE, "..", "include", "hmcgpu.h")]
LIB = os.path.join(HERE, "lib", "libhmcgpu.so")
TAG = os.environ.get("HMC_TAG")             # experiment knob: build/load a side library lib/libhmcgpu_<tag>.so ...
DEFS = os.environ.get("HMC_DEFS", "").split()   # ... compiled with these extra -D flags (e.g. -DHMC_DEV_F3 -DHMC_MINBLOCKS=4)
OBJ = os.path.join(HERE, "lib", "obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-O2",
         "-Xptxas", "-v", "--fmad=true"]
UNITS = [("hmcgpu", "hmcgpu.cu", [])] + [
    (f"gibbs_{r}_{k}", "gibbs_inst.cu", [f"-DHMC_R={r}", f"-DHMC_K={k}"] + (["-DHMC_WITH_PAIR"] if r == "float" else []))
    for r in ("float", "double") for k in (2, 3, 4)] + [
    (f"gibbs_{r}_{k}", "gibbs_inst.cu", [f"-DHMC_R={r}", f"-DHMC_K={k}"]) for r in ("float", "double") for k in (5, 6, 7, 8)] + [
    (f"gibbs_wide_{r}", "gibbs_wide_inst.cu", [f"-DHMC_R={r}"]) for r in ("float", "double")]


def _compile(unit):
    name, src, defs = unit
    defs = defs + DEFS
    srcp, obj = os.path.join(CSRC, src), os.path.join(OBJ, name + ".o")
    deps = [srcp] + HEADERS
    if os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(d) for d in deps):
        return name, obj, None
    res = subprocess.run([NVCC, *FLAGS, *defs, "-c", "-o", obj, srcp], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {name}:\n" + res.stdout + res.stderr)
    return name, obj, res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    global LIB, OBJ
    units = UNITS
    if TAG:
        LIB = os.path.join(HERE, "lib", f"libhmcgpu_{TAG}.so")
        OBJ = os.path.join(HERE, "lib", f"obj_{TAG}")
        if os.path.exists(LIB) and not force:
            return LIB                          # side libraries are never rebuilt implicitly
        if "-DHMC_DEV_F3" in DEFS:              # fp32, K=3 only: the bench path, for quick A/B builds
            units = [u for u in UNITS if u[0] in ("hmcgpu", "gibbs_float_3", "gibbs_wide_float")]
    sources = [os.path.join(CSRC, f) for f in ("hmcgpu.cu", "gibbs_inst.cu", "gibbs_wide_inst.cu")] + HEADERS
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in sources):
        return LIB                              # e.g. on the GPU box: the prebuilt library travels, the objects do not
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(_compile, units))
    rebuilt = [r for r in results if r[2] is not None]
    objs = [r[1] for r in results]
    if rebuilt or not os.path.exists(LIB):
        res = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB, *objs],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
        for name, _, log in rebuilt:
            with open(os.path.join(HERE, "lib", f"ptxas_{name}{'_' + TAG if TAG else ''}.log"), "w") as f:
                f.write(log)
            if verbose:
                print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
