#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests9.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests9.log
HMCGPU_VERBOSE=1 python bench.py --workload c4 --steps 2 --warmup 2 --burnin 10 --nrun 100 --no-cpu-baseline --no-side-records > gpurun_out/r2_c4_trace3.json 2> gpurun_out/r2_c4_trace3.err; echo "c4 rc=$?"
grep -E "hmcgpu\]|e2e per step" gpurun_out/r2_c4_trace3.err | tail -14
python -c "
import json; l = json.load(open('gpurun_out/r2_c4_trace3.json')); print('c4 value %.4e e2e %.4e' % (l['value'], l['e2e']['value']), l['roofline']['kernel'])"
