# experiment driver: bench under different (min blocks per SM, task groups, sweeps per launch)
run() { # mb groups L
  if [ "$1" = 6 ]; then unset HMC_MINBLOCKS; else export HMC_MINBLOCKS=$1; fi
  export HMCGPU_GROUPS=$2 HMCGPU_SWEEPS_PER_LAUNCH=$3
  timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/b_$1_$2_$3.json 2> gpurun_out/b_$1_$2_$3.err
  python -c "
import json; d=json.load(open('gpurun_out/b_$1_$2_$3.json')); print('mb$1 G$2 L$3', '%.4g'%d['value'], '%.1f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['gpu_launches'])" || tail -3 gpurun_out/b_$1_$2_$3.err
}
for cfg in "$@"; do run $cfg; done
