# C5 (BASELINE configs[4]): state-count / length sweep on one GPU; every run writes gpurun_out/c5_K<k>_T<t>.json
run() { # K T chains burnin nrun
  timeout 900 python bench.py --workload c5 --states $1 --length $2 --chains $3 --steps 2 --warmup 2 --burnin $4 --nrun $5 --no-cpu-baseline > gpurun_out/c5_K$1_T$2.json 2> gpurun_out/c5_K$1_T$2.err || tail -3 gpurun_out/c5_K$1_T$2.err
}
run 3 600 131072 100 100
run 3 2000 65536 50 50
run 3 20000 8192 20 20
run 8 600 65536 100 100
run 8 2000 32768 50 50
run 8 20000 4096 10 10
run 32 600 16384 30 30
run 32 2000 8192 10 10
run 32 20000 1024 4 4
