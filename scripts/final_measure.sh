# Round-end measurement pass on one B200 (run through gpurun): tests, default bench, launch list, side workloads.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python bench.py --workload c4 --steps 2 --warmup 3 --burnin 10 --nrun 100 > gpurun_out/r1_c4_fp32.json 2> gpurun_out/r1_c4_fp32.err
python bench.py --workload c4 --precision 64 --steps 2 --warmup 2 --burnin 10 --nrun 100 > gpurun_out/r1_c4_fp64.json 2> gpurun_out/r1_c4_fp64.err
python bench.py --precision 64 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r1_c2_fp64.json 2> gpurun_out/r1_c2_fp64.err
python bench.py --workload sig --chains 16 --steps 2 --warmup 3 --burnin 100 --nrun 100 > gpurun_out/r1_sig.json 2> gpurun_out/r1_sig.err
python bench.py --workload c1 --steps 2 --warmup 1 > gpurun_out/r1_c1.json 2> gpurun_out/r1_c1.err
python bench.py --chains 1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_c2_1chain.json 2> gpurun_out/r1_c2_1chain.err
python bench.py --chains 32 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_c2_32chains.json 2> gpurun_out/r1_c2_32chains.err
python scripts/official_run_report.py > gpurun_out/official_report_fp32.json 2> gpurun_out/official.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
