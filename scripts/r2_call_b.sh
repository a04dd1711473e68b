#!/bin/bash
# Round-2 GPU call B: GPU suite, C4 end-to-end phase trace, the ncu evidence of the headline kernel (scripts/profile_headline.sh),
# then the default bench line (which reads the fresh profile) and the reference arm.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests8.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests8.log
HMCGPU_VERBOSE=1 python bench.py --workload c4 --steps 2 --warmup 2 --burnin 10 --nrun 100 --no-cpu-baseline > gpurun_out/r2_c4_trace2.json 2> gpurun_out/r2_c4_trace2.err; echo "c4 rc=$?"
grep -E "hmcgpu\]|e2e per step" gpurun_out/r2_c4_trace2.err | tail -12
bash scripts/profile_headline.sh
cp gpurun_out/r2_headline_profile.json profiles/r2_headline_profile.json
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
l = json.load(open("gpurun_out/r2_bench_n1.json"))
print("value %.4e e2e %.4e" % (l["value"], l["e2e"]["value"]), "roofline", {k: l["roofline"].get(k) for k in ("bound", "frac", "achieved", "peak", "unit")})
print("c4", {k: (v["value"], v["e2e"]) for k, v in l["c4"].items() if isinstance(v, dict)})
print("fp64", l["fp64"]["value"])
PY
