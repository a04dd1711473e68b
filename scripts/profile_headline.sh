#!/bin/bash
# The ncu evidence behind bench.py's roofline record, taken on the GPU box AFTER the plain bench has exited 0:
#   (1) launch list of the default bench command (every kernel launch with its duration; serialised, cold cache: shares, not absolutes)
#   (2) ncu --set full (+ source view) of two sweep-kernel launches with ALL warp tasks in one launch (one of burn-in sweeps, one of
#       saved draws) -> gpurun_out/r2_headline_profile.json (copied to profiles/ and committed; it carries the hot-source hash)
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-records > gpurun_out/r2_pre_profile_bench.json 2> gpurun_out/r2_pre_profile_bench.err || { echo "bench failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-side-records > gpurun_out/r2_launches.log 2>&1
echo "launch list rc=$?"
HMCGPU_GROUPS=1 HMCGPU_SWEEPS_PER_LAUNCH=1000 timeout 900 ncu --set full --import-source on --clock-control none -k regex:gibbs_sweeps -c 2 -f \
  -o gpurun_out/r2_headline python bench.py --steps 1 --warmup 0 --burnin 16 --nrun 16 --no-cpu-baseline --no-side-records > gpurun_out/r2_headline.log 2>&1
echo "full capture rc=$?"
# 16 sweeps x sum_w T_w (175 250) x 256 chains
python profiles/headline_profile.py gpurun_out/r2_headline.ncu-rep 717824000 "all 4000 warp tasks in one launch of 16 sweeps (HMCGPU_GROUPS=1), burn-in launch and saved-draw launch" > gpurun_out/r2_headline_profile.json
python profiles/ncu_summary.py gpurun_out/r2_headline.ncu-rep > gpurun_out/r2_headline_ncu_full.txt
cat gpurun_out/r2_headline_profile.json | head -20
