#!/bin/bash
# fp64 sweep kernel: blocks per SM / registers A/B on the C2 job (default: 3 blocks at 168 registers, ring depth 4)
set -u
mkdir -p gpurun_out
one() { # tag defs
  HMC_TAG=$1 HMC_DEFS="$2" python bench.py --precision 64 --steps 2 --warmup 2 --no-cpu-baseline --no-side-records > gpurun_out/r2_f64_${1:-default}.json 2> gpurun_out/r2_f64_${1:-default}.err
  python -c "
import json; l=json.load(open('gpurun_out/r2_f64_${1:-default}.json')); print('${1:-default} value %.4e ms %.2f clk %s check %s' % (l['value'], l['ms_per_step'], l['clocks']['sm_mhz'], l['check']))"
}
python bench.py --precision 64 --steps 2 --warmup 2 --no-cpu-baseline --no-side-records > gpurun_out/r2_f64_default.json 2> gpurun_out/r2_f64_default.err
python -c "
import json; l=json.load(open('gpurun_out/r2_f64_default.json')); print('default value %.4e ms %.2f clk %s check %s' % (l['value'], l['ms_per_step'], l['clocks']['sm_mhz'], l['check']))"
one d2 "-DHMC_MINBLOCKS_F64=2 -DHMC_DEV_D3"
one d4r3 "-DHMC_MINBLOCKS_F64=4 -DHMC_RING_STAGES=3 -DHMC_DEV_D3"
one d4 "-DHMC_MINBLOCKS_F64=4 -DHMC_DEV_D3"
