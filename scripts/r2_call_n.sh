#!/bin/bash
# ring rows read through a 32-bit shared address (default) vs the generic form (side library), fp64 with 4 blocks / ring depth 3;
# then the ncu evidence of the headline kernel again (the hot sources changed) and the closing measurement
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests13.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests13.log
ab() { # tag defs
  HMC_TAG=$1 HMC_DEFS="$2" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-records > gpurun_out/r2_lds_${1:-new}.json 2> gpurun_out/r2_lds_${1:-new}.err
  python -c "
import json; l=json.load(open('gpurun_out/r2_lds_${1:-new}.json')); print('${1:-new} value %.4e ms %.2f clk %s' % (l['value'], l['ms_per_step'], l['clocks']['sm_mhz']), l['check'])"
}
for rep in 1 2; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-records > gpurun_out/r2_lds_new.json 2> gpurun_out/r2_lds_new.err
python -c "
import json; l=json.load(open('gpurun_out/r2_lds_new.json')); print('new value %.4e ms %.2f clk %s' % (l['value'], l['ms_per_step'], l['clocks']['sm_mhz']), l['check'])"
ab lds0 "-DHMC_RING_LDS32=0 -DHMC_DEV_F3"
done
python bench.py --precision 64 --steps 2 --warmup 2 --no-cpu-baseline --no-side-records > gpurun_out/r2_f64_new.json 2> gpurun_out/r2_f64_new.err
python -c "
import json; l=json.load(open('gpurun_out/r2_f64_new.json')); print('fp64 value %.4e ms %.2f' % (l['value'], l['ms_per_step']), l['check'])"
bash scripts/profile_headline.sh
