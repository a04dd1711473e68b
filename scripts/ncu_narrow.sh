#!/bin/bash
# ncu --set full capture (with source view) of ONE sweep-kernel launch of the C2 job at a given width:
#   scripts/ncu_narrow.sh <name> <chains per window>        (honours HMC_TAG / HMC_DEFS for side libraries)
name=$1; chains=${2:-32}
mkdir -p gpurun_out
HMCGPU_GROUPS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gibbs_ -s 1 -c 1 -f -o gpurun_out/$name \
  python bench.py --chains $chains --steps 1 --warmup 0 --burnin 16 --nrun 16 --no-cpu-baseline --no-side-records > gpurun_out/$name.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
ls -la gpurun_out/$name.*
