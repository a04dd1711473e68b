"""Runs the reference's complete official workload (all 460 end dates of data/output/official, real inflation series)
through the GPU path and reports the deviation of the posterior means from the reference's own published summaries.
Usage (GPU box): python scripts/official_run_report.py [--chains 64 --burnin 2000 --nrun 2000 --precision 32]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hmc_jl_b200 as H
from conftest import load_inflation, GOLDEN

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=64)
ap.add_argument("--burnin", type=int, default=2000)
ap.add_argument("--nrun", type=int, default=2000)
ap.add_argument("--precision", type=int, default=32)
args = ap.parse_args()
y, dates = load_inflation()
g = json.load(open(os.path.join(GOLDEN, "official_summary_all.json")))
ends = np.array(g["end_index"], dtype=np.int32)
t0 = time.perf_counter()
o = H.estimate_windows(y, np.ones_like(ends), ends, K=3, n_chains=args.chains, burnin=args.burnin, nrun=args.nrun,
                       horizons=(12,), precision=args.precision)
dt = time.perf_counter() - t0
m = o.summary_mean
dev = {"filtered_means": m[:, 0:3] - np.array(g["filtered_means"]),
       "filtered_variances_rel": m[:, 3:6] / np.array(g["filtered_variances"]) - 1,
       "filtered_trans_probs": m[:, 6:15] - np.array(g["filtered_trans_probs"]),
       "filtered_state_probs": m[:, 15:18] - np.array(g["filtered_state_probs"])}
fc = np.array(g["forecasts"])
ok = ends + 12 <= len(y)
dev["forecast_12"] = (m[:, 18] - fc[:, 0])[ok]
rep = {"end_dates": len(ends), "chains": args.chains, "burnin": args.burnin, "nrun": args.nrun, "precision": args.precision,
       "wall_s": dt, "gpu_ms": o.gpu_ms, "state_steps": int(o.state_steps), "events": int(o.events)}
for k, v in dev.items():
    a = np.abs(v)
    rep[k] = {"max_abs": float(a.max()), "p99_abs": float(np.quantile(a, 0.99)), "median_abs": float(np.median(a)),
              "worst_end_index": int(ends[ok][np.argmax(a)] if k == "forecast_12" else ends[np.unravel_index(np.argmax(a), a.shape)[0]])}
print(json.dumps(rep, indent=1))
