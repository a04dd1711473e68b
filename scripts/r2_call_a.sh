#!/bin/bash
# Round-2 GPU call A: full GPU suite on the default library, A/B of the ring variants (LDGSTS ring vs 1-D TMA bulk copies) on the
# headline workload, phase trace of the wide batch C4 end to end, ncu full capture of the fp64 sweep kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests7.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests7.log
ab() {  # tag defs
  HMC_TAG=$1 HMC_DEFS="$2" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-records > gpurun_out/r2_ab_${1:-default}.json 2> gpurun_out/r2_ab_${1:-default}.err
  python - <<PY
import json
try:
    l = json.load(open("gpurun_out/r2_ab_${1:-default}.json"))
    print("${1:-default}", "value %.4e" % l["value"], "ms", round(l["ms_per_step"], 2), "clk", l["clocks"]["sm_mhz"], "check", l["check"])
except Exception as e:
    print("${1:-default}", "failed", e)
PY
}
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-records > gpurun_out/r2_ab_default.json 2> gpurun_out/r2_ab_default.err
ab tma "-DHMC_TMA=1 -DHMC_DEV_F3"
ab tma2 "-DHMC_TMA=2 -DHMC_RING_STAGES=6 -DHMC_DEV_F3"
ab ring6 "-DHMC_RING_STAGES=6 -DHMC_DEV_F3"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-records > gpurun_out/r2_ab_default2.json 2> gpurun_out/r2_ab_default2.err
python - <<PY
import json
for t in ("default", "default2"):
    l = json.load(open(f"gpurun_out/r2_ab_{t}.json"))
    print(t, "value %.4e" % l["value"], "ms", round(l["ms_per_step"], 2), "clk", l["clocks"]["sm_mhz"], "check", l["check"])
PY
HMC_TAG=tma HMC_DEFS="-DHMC_TMA=1 -DHMC_DEV_F3" python -m pytest tests -m gpu -q -k "golden_summaries or full_size or sharding_invariance or repeatable or complete_official or reference_integration" > gpurun_out/r2_tma_tests.log 2>&1; echo "tma tests rc=$?"; tail -3 gpurun_out/r2_tma_tests.log
HMCGPU_VERBOSE=1 python bench.py --workload c4 --steps 2 --warmup 2 --burnin 10 --nrun 100 --no-cpu-baseline > gpurun_out/r2_c4_trace.json 2> gpurun_out/r2_c4_trace.err; echo "c4 rc=$?"
grep -E "hmcgpu\]|e2e per step" gpurun_out/r2_c4_trace.err | tail -24
HMCGPU_STAGE_MB=0 HMCGPU_VERBOSE=1 python bench.py --workload c4 --steps 2 --warmup 2 --burnin 10 --nrun 100 --no-cpu-baseline > gpurun_out/r2_c4_trace_nostage.json 2> gpurun_out/r2_c4_trace_nostage.err
grep -E "hmcgpu\]|e2e per step" gpurun_out/r2_c4_trace_nostage.err | tail -12
HMCGPU_GROUPS=1 HMCGPU_SWEEPS_PER_LAUNCH=1000 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gibbs_sweeps -c 1 -f \
  -o gpurun_out/r2_fp64 python bench.py --precision 64 --steps 1 --warmup 0 --burnin 16 --nrun 16 --no-cpu-baseline --no-side-records > gpurun_out/r2_fp64_ncu.log 2>&1
echo "fp64 capture rc=$?"
python profiles/ncu_summary.py gpurun_out/r2_fp64.ncu-rep > gpurun_out/r2_fp64_ncu_full.txt 2>&1
ncu -i gpurun_out/r2_fp64.ncu-rep --page source --csv > gpurun_out/r2_fp64.source.csv 2>/dev/null
head -40 gpurun_out/r2_fp64_ncu_full.txt
