"""GPU check of api.sample_and_forecast_all (warm-started window chaining, the intent of src/Hmc.jl:584-638) — NOT YET RUN:
it was written after round 1's GPU budget was spent.  Run it first thing on a B200:

    python scripts/chaining_check.py [--first 300 --n 8 --burnin 200 --nrun 2000]

It chains n end dates of the real inflation series (short burn-in from the carried state path) and compares every end date's
posterior means with (a) the reference's published summaries and (b) a cold-start estimation of the same end dates in one
batched call.  Expected: agreement at the Monte-Carlo level of 2000 draws of one chain (|dmu| of a few 1e-2)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hmc_jl_b200 as H
from conftest import load_inflation, GOLDEN

ap = argparse.ArgumentParser()
ap.add_argument("--first", type=int, default=300)
ap.add_argument("--n", type=int, default=8)
ap.add_argument("--burnin", type=int, default=200)
ap.add_argument("--nrun", type=int, default=2000)
args = ap.parse_args()
y, dates = load_inflation()
g = json.load(open(os.path.join(GOLDEN, "official_summary_all.json")))
ends = list(range(args.first, args.first + args.n))
ctx = H.Context(0)
t0 = time.perf_counter()
out = H.sample_and_forecast_all(y, dates, range(1, len(y) + 1), [12], ends, D=3, burnin=args.burnin, Nrun=args.nrun,
                                initialburn=3000, initialNrun=100, ctx=ctx)
t_chain = time.perf_counter() - t0
cold = H.estimate_windows(y, np.ones(len(ends), dtype=np.int32), np.array(ends, dtype=np.int32), K=3, n_chains=16, burnin=2000,
                          nrun=2000, horizons=(12,), precision=64, ctx=ctx)
ctx.close()
gi = [g["end_index"].index(e) for e in ends]
gold_mu = np.array([g["filtered_means"][i] for i in gi])
gold_fc = np.array([g["forecasts"][i][0] for i in gi])
rep = {"end_dates": ends, "seconds_chained": t_chain, "events": out["events"],
       "max_abs_dmu_vs_reference": float(np.abs(out["μ"] - gold_mu).max()),
       "max_abs_dmu_vs_cold_start": float(np.abs(out["μ"] - cold.summary_mean[:, 0:3]).max()),
       "max_abs_dforecast_vs_reference": float(np.abs(out["forecasts"][:, 0] - gold_fc).max()),
       "mu_chained": out["μ"].round(4).tolist(), "mu_reference": gold_mu.round(4).tolist()}
print(json.dumps(rep, indent=1))
