"""Builds libhmcgpu.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.  No torch, no JIT cache.

The Gibbs sweep kernel family is instantiated once per (precision, K) in its own translation unit
(csrc/gibbs_inst.cu with -DHMC_R/-DHMC_K); the units are compiled in parallel and linked with csrc/hmcgpu.cu."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HEADERS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join(HERE, "..", "include", "hmcgpu.h")]
LIB = os.path.join(HERE, "lib", "libhmcgpu.so")
TAG = os.environ.get("HMC_TAG")             # experiment knob: build/load a side library lib/libhmcgpu_<tag>.so ...
DEFS = os.environ.get("HMC_DEFS", "").split()   # ... compiled with these extra -D flags (e.g. -DHMC_DEV_F3 -DHMC_MINBLOCKS=4)
OBJ = os.path.join(HERE, "lib", "obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-O2",
         "-Xptxas", "-v", "--fmad=true"]
UNITS = [("hmcgpu", "hmcgpu.cu", []), ("build_info", "build_info.cu", [])] + [
    (f"gibbs_{r}_{k}", "gibbs_inst.cu", [f"-DHMC_R={r}", f"-DHMC_K={k}"])
    for r in ("float", "double") for k in (2, 3, 4)] + [
    (f"gibbs_{r}_{k}", "gibbs_inst.cu", [f"-DHMC_R={r}", f"-DHMC_K={k}"]) for r in ("float", "double") for k in (5, 6, 7, 8)] + [
    (f"gibbs_wide_{r}", "gibbs_wide_inst.cu", [f"-DHMC_R={r}"]) for r in ("float", "double")]


def source_hash():
    """sha256 (16 hex digits) over every file the library is built from and the compiler flags.  The library carries it
    (hmcgpu_build_info), build() reuses a library only when it matches, and binding.load() refuses one that does not.
    None when the sources are not there (a deployment that ships only the .so)."""
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) if os.path.isdir(CSRC) else []
    hdr = os.path.join(HERE, "..", "include", "hmcgpu.h")
    if not files or not os.path.exists(hdr):
        return None
    h = hashlib.sha256()
    for f in files + [hdr]:
        h.update(os.path.basename(f).encode() + b"\0")
        h.update(open(f, "rb").read())
    h.update(" ".join(FLAGS + DEFS).encode())
    return h.hexdigest()[:16]


HOT_SOURCES = ("gibbs_kernel.cuh", "hmm_device.cuh", "rng.cuh")


def hot_source_hash():
    """sha256 (16 hex digits) over the sources of the headline sweep kernel and the compiler flags: what an ncu capture of that
    kernel is a capture OF (profiles/r2_headline_profile.json carries it; bench.py ignores a capture of other sources)."""
    h = hashlib.sha256()
    for f in HOT_SOURCES:
        h.update(f.encode() + b"\0")
        h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(" ".join(FLAGS + DEFS).encode())
    return h.hexdigest()[:16]


def built_hash(lib):
    """The hash a built library carries, read without loading it into this process (the sidecar build() writes)."""
    try:
        return open(lib + ".hash").read().strip()
    except OSError:
        return None


def _compile(unit):
    name, src, defs = unit
    defs = defs + DEFS
    srcp, obj = os.path.join(CSRC, src), os.path.join(OBJ, name + ".o")
    deps = [srcp] + HEADERS
    if name == "build_info":                    # carries the source hash: rebuilt whenever any source changed
        defs = defs + [f'-DHMC_SRC_HASH="{source_hash()}"']
        deps = None
    if deps is not None and os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(d) for d in deps):
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
            units = [u for u in UNITS if u[0] in ("hmcgpu", "build_info", "gibbs_float_3", "gibbs_wide_float")]
        if "-DHMC_DEV_D3" in DEFS:              # fp64, K=3 only
            units = [u for u in UNITS if u[0] in ("hmcgpu", "build_info", "gibbs_double_3", "gibbs_wide_double")]
    want = source_hash()
    if not force and os.path.exists(LIB) and built_hash(LIB) == want:
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
        with open(LIB + ".hash", "w") as f:
            f.write(want + "\n")
        for name, _, log in rebuilt:
            with open(os.path.join(HERE, "lib", f"ptxas_{name}{'_' + TAG if TAG else ''}.log"), "w") as f:
                f.write(log)
            if verbose:
                print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
