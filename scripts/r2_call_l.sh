#!/bin/bash
# per-sweep time of the segment kernel against the window length (16 384 chains, every warp resident): fixed per-sweep cost vs per-step cost
mkdir -p gpurun_out
for l in 4 8 0; do for T in 152 304 608 1216; do
  HMCGPU_SEG_LANES=$l python bench.py --workload c5 --states 3 --length $T --chains 16384 --steps 2 --warmup 1 --burnin 200 --nrun 200 --no-cpu-baseline --no-side-records > gpurun_out/r2_segT_l${l}_T$T.json 2> gpurun_out/r2_segT_l${l}_T$T.err
  python -c "
import json; l=json.load(open('gpurun_out/r2_segT_l${l}_T$T.json')); print('lanes $l T $T  us/sweep %.2f  value %.3e' % (1e3*l['ms_per_step']/400, l['value']), l['roofline']['kernel'][:16])"
done; done
