#!/bin/bash
# what is the segment kernel's fixed per-sweep cost made of?  16 384 chains x T = 304, 4 lanes: phase barriers, warm-up length, block size
mkdir -p gpurun_out
one() { # label env...
  label=$1; shift
  env HMCGPU_SEG_LANES=4 "$@" python bench.py --workload c5 --states 3 --length ${T:-304} --chains 16384 --steps 2 --warmup 1 --burnin 200 --nrun 200 --no-cpu-baseline --no-side-records > gpurun_out/r2_segfix_$label.json 2> gpurun_out/r2_segfix_$label.err
  python -c "
import json; l=json.load(open('gpurun_out/r2_segfix_$label.json')); print('$label  us/sweep %.2f' % (1e3*l['ms_per_step']/400))"
}
one base A=1
one barriers0 HMCGPU_SEG_BARRIERS=0
one barriers1 HMCGPU_SEG_BARRIERS=1
one warm16 HMCGPU_SEG_WARMUP=16
one warm64 HMCGPU_SEG_WARMUP=64
one warm0 HMCGPU_SEG_WARMUP=0
one thr128 HMCGPU_SEG_THREADS=128
one groups1 HMCGPU_GROUPS=1
one groups8 HMCGPU_GROUPS=8
one spl64 HMCGPU_SWEEPS_PER_LAUNCH=64
one spl4 HMCGPU_SWEEPS_PER_LAUNCH=4
# the same on the shared-series C2 job (500 windows x 32 chains)
for v in "A=1" "HMCGPU_SWEEPS_PER_LAUNCH=64" "HMCGPU_SWEEPS_PER_LAUNCH=4" "HMCGPU_GROUPS=1" "HMCGPU_GROUPS=8" "HMCGPU_SEG_BARRIERS=0"; do
  env $v python bench.py --chains 32 --steps 3 --warmup 2 --no-cpu-baseline --no-side-records > gpurun_out/r2_segfix_c2.json 2> gpurun_out/r2_segfix_c2.err
  python -c "
import json; l=json.load(open('gpurun_out/r2_segfix_c2.json')); print('C2x32 $v  us/sweep %.2f value %.4e' % (1e3*l['ms_per_step']/2000, l['value']))"
done
