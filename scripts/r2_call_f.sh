#!/bin/bash
# 8 lanes per chain in the segment kernel: parity (the whole [seg] leg of the GPU suite with 8 lanes), speed on C2 x 32 / x 64; K = 32 lane kernel
set -u
mkdir -p gpurun_out
HMC_TEST_SEG_LANES=8 python -m pytest tests -m gpu -x -q -k "seg" > gpurun_out/r2_tests_seg8.log 2>&1; echo "seg8 tests rc=$?"; tail -3 gpurun_out/r2_tests_seg8.log
for c in 32 64; do for l in 4 8; do
  HMCGPU_SEG_LANES=$l python bench.py --chains $c --steps 3 --warmup 3 --no-cpu-baseline --no-side-records > gpurun_out/r2_seg_c${c}_l$l.json 2> gpurun_out/r2_seg_c${c}_l$l.err
  python -c "
import json; l=json.load(open('gpurun_out/r2_seg_c${c}_l$l.json')); print('chains $c lanes $l value %.4e ms %.2f' % (l['value'], l['ms_per_step']), l['roofline']['kernel'][:20])"
done; done
python bench.py --workload c5 --states 32 --length 600 --chains 16384 --steps 2 --warmup 2 --burnin 30 --nrun 30 --no-cpu-baseline --no-side-records > gpurun_out/r2_c5_K32_T600.json 2> gpurun_out/r2_c5_K32_T600.err
python -c "
import json; l=json.load(open('gpurun_out/r2_c5_K32_T600.json')); print('K32 T600 value %.4e e2e %.4e' % (l['value'], l['e2e']['value']))"
python bench.py --workload c5 --states 16 --length 600 --chains 32768 --steps 2 --warmup 2 --burnin 30 --nrun 30 --no-cpu-baseline --no-side-records > gpurun_out/r2_c5_K16_T600.json 2> gpurun_out/r2_c5_K16_T600.err
python -c "
import json; l=json.load(open('gpurun_out/r2_c5_K16_T600.json')); print('K16 T600 value %.4e e2e %.4e' % (l['value'], l['e2e']['value']))"
