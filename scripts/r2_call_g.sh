#!/bin/bash
# lane-per-state kernel: scalar FFMA matrix-vector step (default) vs packed FFMA2 (side library -DHMC_WIDE_PACKED=1), K = 32 and K = 16,
# same box back to back; first the lane-kernel GPU tests on both libraries
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests10.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests10.log
HMC_TAG=widepacked HMC_DEFS="-DHMC_WIDE_PACKED=1 -DHMC_DEV_F3" python -m pytest tests -m gpu -q -k "lane_kernel_fp32" > gpurun_out/r2_tests_widepacked.log 2>&1; echo "packed lane tests rc=$?"; tail -2 gpurun_out/r2_tests_widepacked.log
run() { # tag defs K chains
  HMC_TAG=$1 HMC_DEFS="$2" python bench.py --workload c5 --states $3 --length 600 --chains $4 --steps 3 --warmup 2 --burnin 30 --nrun 30 --no-cpu-baseline --no-side-records > gpurun_out/r2_wide_${1:-scalar}_K$3.json 2> gpurun_out/r2_wide_${1:-scalar}_K$3.err
  python -c "
import json; l=json.load(open('gpurun_out/r2_wide_${1:-scalar}_K$3.json')); print('${1:-scalar} K$3 value %.4e ms %.2f clk %s events %s' % (l['value'], l['ms_per_step'], l['clocks']['sm_mhz'], l['check']['events']))"
}
def() { # K chains
  python bench.py --workload c5 --states $1 --length 600 --chains $2 --steps 3 --warmup 2 --burnin 30 --nrun 30 --no-cpu-baseline --no-side-records > gpurun_out/r2_wide_scalar_K$1.json 2> gpurun_out/r2_wide_scalar_K$1.err
  python -c "
import json; l=json.load(open('gpurun_out/r2_wide_scalar_K$1.json')); print('scalar K$1 value %.4e ms %.2f clk %s events %s' % (l['value'], l['ms_per_step'], l['clocks']['sm_mhz'], l['check']['events']))"
}
for rep in 1 2; do
def 32 16384
run widepacked "-DHMC_WIDE_PACKED=1 -DHMC_DEV_F3" 32 16384
done
def 16 32768
run widepacked "-DHMC_WIDE_PACKED=1 -DHMC_DEV_F3" 16 32768
