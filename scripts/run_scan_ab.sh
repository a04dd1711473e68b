# narrow-batch A/B: time-parallel warp-per-chain kernel (default below HMCGPU_SCAN_MAX_CHAINS chains) vs thread-per-chain
for mode in $MODES; do
  if [ $mode = thread ]; then export HMCGPU_SCAN_MAX_CHAINS=0; else export HMCGPU_SCAN_MAX_CHAINS=1000000; fi
  python bench.py --workload c1 --steps 2 --warmup 1 > gpurun_out/sc_c1_$mode.json 2> gpurun_out/sc_c1_$mode.err
  for ch in $CHAINS; do
    python bench.py --chains $ch --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/sc_c2x${ch}_$mode.json 2> gpurun_out/sc_c2x${ch}_$mode.err
  done
done
