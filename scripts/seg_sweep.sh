#!/bin/bash
# segment-kernel parameter sweep on the C2 job at a given width: lanes per chain x warm-up length (fallback counts on stderr)
#   scripts/seg_sweep.sh <chains per window> "<lanes list>" "<warm-up list>"
c=${1:-32}; lanes=${2:-"4"}; warms=${3:-"32"}
mkdir -p gpurun_out
for L in $lanes; do for W in $warms; do
  HMCGPU_VERBOSE=1 HMCGPU_SEG_LANES=$L HMCGPU_SEG_WARMUP=$W python bench.py --chains $c --steps 2 --warmup 1 --no-cpu-baseline --no-side-records \
     > gpurun_out/segsweep.json 2> gpurun_out/segsweep.err
  python - <<PY
import json
try:
    l = json.load(open("gpurun_out/segsweep.json"))
    fb = [x for x in open("gpurun_out/segsweep.err") if "chain-sweeps" in x]
    print("chains", $c, "L", $L, "W", $W, "value %.3e" % l["value"], "ms %.1f" % l["ms_per_step"], "mu", [round(v, 4) for v in l["check"]["mu_mean_longest_window"]], "ev", l["check"]["events"], "|", fb[-1].split("):")[-1].strip() if fb else "")
except Exception as e:
    print("L", $L, "W", $W, "failed", e); print(open("gpurun_out/segsweep.err").read()[-800:])
PY
done; done
