#!/bin/bash
# mixed segment counts (8 lanes for the longest windows, 4 for the rest) vs uniform 4 lanes on the mid-width C2 curve
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests11.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests11.log
for c in 16 32 48 64; do for m in 1 0; do
  HMCGPU_VERBOSE=1 HMCGPU_SEG_MIXED=$m python bench.py --chains $c --steps 3 --warmup 3 --no-cpu-baseline --no-side-records > gpurun_out/r2_mix_c${c}_m$m.json 2> gpurun_out/r2_mix_c${c}_m$m.err
  python -c "
import json; l=json.load(open('gpurun_out/r2_mix_c${c}_m$m.json')); print('chains $c mixed $m value %.4e ms %.2f' % (l['value'], l['ms_per_step']), l['roofline']['kernel'][:16], l['check'])"
  grep -m1 "longest of" gpurun_out/r2_mix_c${c}_m$m.err
done; done
