#!/bin/bash
# observations of the one shared series read from a shared-memory copy (new) vs from global memory (previous commit, built as a side
# tree under _old/): GPU suite on the new library, then the headline workload on both, back to back, twice
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests15.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests15.log
for rep in 1 2; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-records > gpurun_out/r2_ysm_new.json 2> gpurun_out/r2_ysm_new.err
python -c "
import json; l=json.load(open('gpurun_out/r2_ysm_new.json')); print('new value %.4e ms %.2f clk %s' % (l['value'], l['ms_per_step'], l['clocks']['sm_mhz']), l['check'])"
(cd _old && HMC_TAG=old HMC_DEFS="-DHMC_DEV_F3" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-records > ../gpurun_out/r2_ysm_old.json 2> ../gpurun_out/r2_ysm_old.err)
python -c "
import json; l=json.load(open('gpurun_out/r2_ysm_old.json')); print('old value %.4e ms %.2f clk %s' % (l['value'], l['ms_per_step'], l['clocks']['sm_mhz']), l['check'])"
done
python bench.py --precision 64 --steps 2 --warmup 2 --no-cpu-baseline --no-side-records > gpurun_out/r2_f64_new2.json 2> gpurun_out/r2_f64_new2.err
python -c "
import json; l=json.load(open('gpurun_out/r2_f64_new2.json')); print('fp64 value %.4e ms %.2f' % (l['value'], l['ms_per_step']), l['check'])"
