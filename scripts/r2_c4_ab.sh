#!/bin/bash
# C4 (65 536 series x T=2000, 10+100 sweeps) device time by sweep kernel: thread per chain vs 2 / 4 lanes per chain
mkdir -p gpurun_out
for cfg in "0:" "2:" "4:" "2:128" "2:256"; do
  lanes=${cfg%%:*}; thr=${cfg#*:}
  HMCGPU_SEG_LANES=$lanes ${thr:+HMCGPU_SEG_THREADS=$thr} python bench.py --workload c4 --steps 2 --warmup 2 --burnin 10 --nrun 100 --no-cpu-baseline --no-side-records > gpurun_out/r2_c4_l${lanes}_t${thr:-auto}.json 2> gpurun_out/r2_c4_l${lanes}_t${thr:-auto}.err
  python - <<PY
import json
l = json.load(open("gpurun_out/r2_c4_l${lanes}_t${thr:-auto}.json"))
print("lanes $lanes threads ${thr:-auto}", "value %.4e" % l["value"], "ms", round(l["ms_per_step"], 2), "e2e %.4e" % l["e2e"]["value"], l["roofline"]["kernel"][:24])
PY
done
for cfg in "0:" "2:"; do
  lanes=${cfg%%:*}
  HMCGPU_SEG_LANES=$lanes python bench.py --workload c4 --precision 64 --steps 2 --warmup 2 --burnin 10 --nrun 100 --no-cpu-baseline --no-side-records > gpurun_out/r2_c4f64_l${lanes}.json 2> gpurun_out/r2_c4f64_l${lanes}.err
  python - <<PY
import json
l = json.load(open("gpurun_out/r2_c4f64_l${lanes}.json"))
print("fp64 lanes $lanes", "value %.4e" % l["value"], "ms", round(l["ms_per_step"], 2), "e2e %.4e" % l["e2e"]["value"], l["roofline"]["kernel"][:24])
PY
done
