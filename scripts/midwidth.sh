#!/bin/bash
# Mid-width curve of the C2 job on one GPU: 500 windows x {16,32,64,128,256} chains (bench lines under gpurun_out/<tag>_mid_<c>.json).
# 500 x 32 = 16 000 chains is one GPU's share of the fixed 500 x 256 job at 8 GPUs (strong scaling).
tag=${1:-r2}
mkdir -p gpurun_out
for c in ${CHAINS:-32 64 128}; do
  python bench.py --chains $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_mid_$c.json 2> gpurun_out/${tag}_mid_$c.err
  python - <<PY
import json
try:
    l = json.load(open("gpurun_out/${tag}_mid_$c.json"))
    print("chains/window", $c, "value %.3e" % l["value"], "e2e %.3e" % l["e2e"]["value"], "ms", round(l["ms_per_step"], 2), "clk", l["clocks"]["sm_mhz"])
except Exception as e:
    print("chains", $c, "failed", e)
PY
done
