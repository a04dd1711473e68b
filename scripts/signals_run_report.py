"""Runs the reference's complete noisy-signal workload — every end date of data/output/signals_official_noise_<κ>_allsignal
(the "make everything a signal" block of code/run_hmm.jl:160-176: 100 perturbed copies of the real inflation series per end
date, every observation a signal with relative imprecision κ) — through the GPU path, ONE hmcgpu_estimate call per noise
level (≈ 45 500 windows, each on its own perturbed series with its own signal mask), and reports the deviation of the
dispersion tables (mean / std over the copies of each copy's posterior means, what code/aggregate.jl writes) from the
reference's own.  σsignal per end date is the realised spread of the golden's saved signal values.
Usage (GPU box): python scripts/signals_run_report.py [--chains 4 --burnin 1500 --nrun 1000 --copies 100 --precision 32]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hmc_jl_b200 as H
from conftest import load_inflation, GOLDEN

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=4)
ap.add_argument("--burnin", type=int, default=1500)
ap.add_argument("--nrun", type=int, default=1000)
ap.add_argument("--copies", type=int, default=100)
ap.add_argument("--precision", type=int, default=32)
ap.add_argument("--noise", type=str, default="0.1,0.3,0.6")
ap.add_argument("--max-dates", type=int, default=0, help="first N end dates only (0 = all)")
args = ap.parse_args()
y, dates = load_inflation()
g = json.load(open(os.path.join(GOLDEN, "signals_allsignal_all.json")))
K, S, T = 3, args.copies, len(y)
names = ["filtered_means", "filtered_variances", "filtered_trans_probs", "filtered_state_probs", "forecasts"]
cols = {"filtered_means": slice(0, 3), "filtered_variances": slice(3, 6), "filtered_trans_probs": slice(6, 15),
        "filtered_state_probs": slice(15, 18), "forecasts": slice(18, 20)}          # hmcgpu_result summary layout (F = 21)
ctx = H.Context(0)
report = {"copies": S, "chains": args.chains, "burnin": args.burnin, "nrun": args.nrun, "precision": args.precision, "noise": {}}
for noise in args.noise.split(","):
    blk = g["noise"][noise]
    ends = np.array(blk["end_index"], dtype=np.int64)
    if args.max_dates:
        ends = ends[:args.max_dates]
    nd = len(ends)
    ssig = np.sqrt(np.mean(np.square(np.array(blk["signal_std"][:nd])), axis=1))
    rng = np.random.default_rng(1234)
    live = (np.arange(T)[None, :] < ends[:, None])                                 # [nd, T]: rows of the window = all signals
    series = np.empty((1 + nd * S, T))
    series[0] = y
    pert = rng.standard_normal((nd, S, T)) * ssig[:, None, None] * live[:, None, :]
    series[1:] = (y[None, None, :] + pert).reshape(nd * S, T)
    del pert
    mask = np.zeros((1 + nd * S, T), dtype=np.uint8)
    mask[1:] = np.repeat(live, S, axis=0)
    nw = nd * S
    two = np.full(K, 2.0)
    spec = H.ProblemSpec(series, np.ones(nw, dtype=np.int32), np.repeat(ends, S), K=K, n_chains=args.chains, burnin=args.burnin,
                         nrun=args.nrun, seed=1234, horizons=(12,), precision=args.precision, flags=H.FLAG_REF_Q1 | H.FLAG_SUMMARY,
                         win_series=np.arange(1, nw + 1), win_init_series=np.zeros(nw), alpha=two, nu=two, kappa=float(noise),
                         is_signal=mask)                                            # xi = NULL: mean of the REAL window (:148-159)
    t0 = time.perf_counter()
    o = H.estimate(ctx, spec)
    dt = time.perf_counter() - t0
    per_copy = o.summary_mean.reshape(nd, S, -1)
    m, sd = per_copy.mean(1), per_copy.std(1, ddof=1)
    rep = {"end_dates": nd, "windows": nw, "wall_s": dt, "gpu_ms": o.gpu_ms, "state_steps": int(o.state_steps),
           "state_steps_per_s": o.state_steps / (o.gpu_ms * 1e-3), "events": int(o.events)}
    for n in names:
        gm, gs = np.array(blk[n]["mean"][:nd]), np.array(blk[n]["std"][:nd])
        om, osd = m[:, cols[n]], sd[:, cols[n]]
        ok = np.isfinite(gm) & np.isfinite(om)
        if n == "forecasts":
            ok &= (ends + 12 <= T)[:, None]
        d = np.abs(om - gm)[ok]
        z = (d / np.sqrt(osd ** 2 / S + gs ** 2 / 100 + 1e-12)[ok])
        rel = (d / np.maximum(np.abs(gm[ok]), 1e-9))
        big = ok & (gs > 0.02)
        rep[n] = {"median_abs": float(np.median(d)), "p99_abs": float(np.quantile(d, 0.99)), "max_abs": float(d.max()),
                  "median_rel": float(np.median(rel)), "median_z": float(np.median(z)), "frac_z_below_3": float((z < 3).mean()),
                  "frac_z_below_5": float((z < 5).mean()),
                  "spread_ratio_median": float(np.median((osd / np.maximum(gs, 1e-12))[big])) if big.any() else None}
    report["noise"][noise] = rep
    del series, mask, spec, o
ctx.close()
print(json.dumps(report, indent=1))
