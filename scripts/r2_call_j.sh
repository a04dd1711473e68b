#!/bin/bash
# how many of the longest windows should get 8 lanes per chain? (HMCGPU_SEG_LONG sweep on C2 x 16 / 24 / 32 chains)
set -u
mkdir -p gpurun_out
for c in 16 24 32; do for n in 0 50 100 150 200 300; do
  HMCGPU_SEG_LONG=$n python bench.py --chains $c --steps 3 --warmup 3 --no-cpu-baseline --no-side-records > gpurun_out/r2_long_c${c}_n$n.json 2> gpurun_out/r2_long_c${c}_n$n.err
  python -c "
import json; l=json.load(open('gpurun_out/r2_long_c${c}_n$n.json')); print('chains $c long $n value %.4e ms %.2f' % (l['value'], l['ms_per_step']))"
done; done
