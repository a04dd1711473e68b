#!/bin/bash
# A/B of a side library (hmc.jl_b200/build.py with HMC_TAG / HMC_DEFS) on the GPU box: the GPU test suite with it, then the
# C2 width curve.   usage: scripts/ab_lib.sh <tag> "<defs>" [chains...]
tag=$1; defs=$2; shift 2
export HMC_TAG=$tag HMC_DEFS="$defs"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests($tag) rc=$?"; tail -3 gpurun_out/${tag}_tests.log
CHAINS="${*:-32 64 128 256}" scripts/midwidth.sh $tag
