#!/bin/bash
# fp32: the one shared series from the shared-memory copy (side library -DHMC_YSM_F32=1) vs through L1 (default), now that the Chain
# struct keeps its size; GPU suite on the default library first
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests16.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests16.log
HMC_TAG=ysm32 HMC_DEFS="-DHMC_YSM_F32=1 -DHMC_DEV_F3" python -m pytest tests -m gpu -q -k "golden_summaries or full_size or repeatable or complete_official" > gpurun_out/r2_tests_ysm32.log 2>&1; echo "ysm32 tests rc=$?"; tail -1 gpurun_out/r2_tests_ysm32.log
for rep in 1 2; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-records > gpurun_out/r2_ysm32_def.json 2> gpurun_out/r2_ysm32_def.err
python -c "
import json; l=json.load(open('gpurun_out/r2_ysm32_def.json')); print('default value %.4e ms %.2f clk %s' % (l['value'], l['ms_per_step'], l['clocks']['sm_mhz']), l['check'])"
HMC_TAG=ysm32 HMC_DEFS="-DHMC_YSM_F32=1 -DHMC_DEV_F3" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-records > gpurun_out/r2_ysm32_on.json 2> gpurun_out/r2_ysm32_on.err
python -c "
import json; l=json.load(open('gpurun_out/r2_ysm32_on.json')); print('ysm32 value %.4e ms %.2f clk %s' % (l['value'], l['ms_per_step'], l['clocks']['sm_mhz']), l['check'])"
done
python bench.py --precision 64 --steps 2 --warmup 2 --no-cpu-baseline --no-side-records > gpurun_out/r2_f64_new3.json 2> gpurun_out/r2_f64_new3.err
python -c "
import json; l=json.load(open('gpurun_out/r2_f64_new3.json')); print('fp64 value %.4e ms %.2f' % (l['value'], l['ms_per_step']), l['check'])"
