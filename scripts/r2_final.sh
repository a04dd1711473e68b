#!/bin/bash
# Round-2 closing measurement on one B200: GPU suite, default bench line (reads profiles/r2_headline_profile.json), reference arm,
# smoke(), the complete official run against the reference's published summaries.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_final.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests_final.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
python scripts/official_run_report.py > gpurun_out/r2_official_run_report_fp32.json 2> gpurun_out/r2_official.err; echo "official rc=$?"
python - <<PY
import json
l = json.load(open("gpurun_out/r2_bench_n1.json"))
print("value %.4e e2e %.4e" % (l["value"], l["e2e"]["value"]), "roofline", {k: l["roofline"].get(k) for k in ("bound", "frac", "algorithmic_frac")}, "clk", l["clocks"])
print("hbm_alg", l["roofline"]["hbm_algorithmic"]["frac"], "hbm_dram", (l["roofline"]["hbm_dram"] or {}).get("frac"), "profile", l["roofline"]["profile"])
print("c4", {k: (v["value"], v["e2e"], v.get("kernel", "")[:14]) for k, v in l["c4"].items() if isinstance(v, dict)})
print("fp64", l["fp64"]["value"], "mid", [(m["chains_per_window"], "%.3e" % m["value"]) for m in l["mid_width"]])
print("cpu", l["cpu_baseline"]["value"], "one-chain", l["one_chain_per_end_date"]["value"])
r = json.load(open("gpurun_out/r2_bench_ref.json")); print("ref", r["value"], r.get("cpu_baseline", {}).get("sample", "")[:80])
PY
