#!/bin/bash
# lane-per-state kernel: packed FFMA2 matrix-vector step vs scalar FFMA (side library), K = 32 and K = 16, same box back to back
set -u
mkdir -p gpurun_out
run() { # tag defs K chains
  HMC_TAG=$1 HMC_DEFS="$2" python bench.py --workload c5 --states $3 --length 600 --chains $4 --steps 3 --warmup 2 --burnin 30 --nrun 30 --no-cpu-baseline --no-side-records > gpurun_out/r2_wide_${1:-packed}_K$3.json 2> gpurun_out/r2_wide_${1:-packed}_K$3.err
  python -c "
import json; l=json.load(open('gpurun_out/r2_wide_${1:-packed}_K$3.json')); print('${1:-packed} K$3 value %.4e ms %.2f clk %s' % (l['value'], l['ms_per_step'], l['clocks']['sm_mhz']))"
}
for rep in 1 2; do
python bench.py --workload c5 --states 32 --length 600 --chains 16384 --steps 3 --warmup 2 --burnin 30 --nrun 30 --no-cpu-baseline --no-side-records > gpurun_out/r2_wide_packed_K32.json 2> gpurun_out/r2_wide_packed_K32.err
python -c "
import json; l=json.load(open('gpurun_out/r2_wide_packed_K32.json')); print('packed K32 value %.4e ms %.2f clk %s' % (l['value'], l['ms_per_step'], l['clocks']['sm_mhz']))"
run widescalar "-DHMC_WIDE_PACKED=0 -DHMC_DEV_F3" 32 16384
done
python bench.py --workload c5 --states 16 --length 600 --chains 32768 --steps 3 --warmup 2 --burnin 30 --nrun 30 --no-cpu-baseline --no-side-records > gpurun_out/r2_wide_packed_K16.json 2> gpurun_out/r2_wide_packed_K16.err
python -c "
import json; l=json.load(open('gpurun_out/r2_wide_packed_K16.json')); print('packed K16 value %.4e ms %.2f' % (l['value'], l['ms_per_step']))"
run widescalar "-DHMC_WIDE_PACKED=0 -DHMC_DEV_F3" 16 32768
