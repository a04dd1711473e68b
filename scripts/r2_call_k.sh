#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "overlapped or staged or estimatesignals_mirror or multiple_series" > gpurun_out/r2_tests12a.log 2>&1; echo "new tests rc=$?"; tail -3 gpurun_out/r2_tests12a.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests12.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests12.log
for ov in 1 0; do
HMCGPU_OVERLAP=$ov HMCGPU_VERBOSE=1 timeout 600 python bench.py --workload c4 --steps 2 --warmup 2 --burnin 10 --nrun 100 --no-cpu-baseline --no-side-records > gpurun_out/r2_c4_ov$ov.json 2> gpurun_out/r2_c4_ov$ov.err; echo "c4 overlap=$ov rc=$?"
grep -E "hmcgpu\] plan|e2e per step" gpurun_out/r2_c4_ov$ov.err | tail -5
python -c "
import json; l = json.load(open('gpurun_out/r2_c4_ov$ov.json')); print('c4 overlap=$ov value %.4e e2e %.4e' % (l['value'], l['e2e']['value']), l['check'])"
done
