# experiment driver: bench side libraries (HMC_TAG) under launch configs "tag groups sweeps_per_launch"
run() {
  export HMC_TAG=$1 HMCGPU_GROUPS=$2 HMCGPU_SWEEPS_PER_LAUNCH=$3
  timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/v_$1_$2_$3.json 2> gpurun_out/v_$1_$2_$3.err
  python -c "
import json; d=json.load(open('gpurun_out/v_$1_$2_$3.json')); print('tag=$1 G$2 L$3', '%.4g'%d['value'], '%.1f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['gpu_launches'], d['check']['mu_mean_longest_window'])" || tail -3 gpurun_out/v_$1_$2_$3.err
}
for cfg in "$@"; do run $cfg; done
