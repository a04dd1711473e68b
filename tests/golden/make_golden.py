"""Regenerates the golden fixtures in this directory from the reference checkout (run in the build container only).

Inputs  (read-only, /root/reference): data/processed/inflation.csv (the monthly series run_hmm.jl estimates on) and
data/output/official/*_summary.csv (the reference's own posterior means over 250 000 draws per end date, produced by
code/run_hmm.jl official <idx> + code/aggregate.jl).  Outputs: inflation_offic_inf.csv, official_summary_subset.json
(every 27th end date, 19 windows) and official_summary_all.json (all 460 end dates, values rounded to 6 digits) — small enough to commit; nothing at test time reads /root/reference.
Header caveat: in the summary files column trans_a_b holds A[b,a].
"""
import csv
import json
import os

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

rows = list(csv.DictReader(open(f"{REF}/data/processed/inflation.csv")))
with open(f"{HERE}/inflation_offic_inf.csv", "w") as f:
    f.write("date,offic_inf\n")
    for r in rows:
        f.write(f"{r['date']},{r['offic_inf']}\n")
dates = [r["date"] for r in rows]

def load(name):
    r = list(csv.reader(open(f"{REF}/data/output/official/{name}_summary.csv")))
    return r[0][1:], {x[0]: [float(v) for v in x[1:]] for x in r[1:]}

names = ["filtered_means", "filtered_variances", "filtered_state_probs", "filtered_trans_probs", "forecasts"]
tabs = {n: load(n) for n in names}
subset = {"source": "data/output/official/<name>_summary.csv", "burnin": 100000, "nrun": 250000, "K": 3,
          "note": "trans_a_b column holds A[b,a]; per-draw values were rounded to 5 digits before averaging",
          "windows": []}
for idx in list(range(120, 580, 27)) + [579]:
    d = dates[idx - 1]
    if all(d in tabs[n][1] for n in names):
        subset["windows"].append({"end_index": idx, "date": d, **{n: tabs[n][1][d] for n in names}})
subset["columns"] = {n: tabs[n][0] for n in names}
json.dump(subset, open(f"{HERE}/official_summary_subset.json", "w"), indent=1)
print(len(subset["windows"]), "windows")

# the complete official run (all 460 end dates, idx 120..579), compact: one row per end date
full = {"source": subset["source"], "note": subset["note"], "columns": subset["columns"], "end_index": [], "date": []}
for n in names:
    full[n] = []
for idx in range(120, 580):
    d = dates[idx - 1]
    if all(d in tabs[n][1] for n in names):
        full["end_index"].append(idx); full["date"].append(d)
        for n in names:
            full[n].append([round(v, 6) for v in tabs[n][1][d]])
json.dump(full, open(f"{HERE}/official_summary_all.json", "w"), separators=(",", ":"))
print(len(full["end_index"]), "end dates in the full table")

# ---- noisy-signal Monte Carlo goldens (estimatesignals!, src/Hmc.jl:868-914): data/output/signals_official_noise_<κ>_allsignal/
# produced by the "make everything a signal" block of code/run_hmm.jl (:160-176): sampleRange = signalRange = start:end,
# signalSave = end-1:end, noise = κ, 100 perturbed copies x (100 000 + 250 000) sweeps, then aggregate.jl's dispersion
# tables = mean and std over the 100 copies of each copy's posterior means.  signal_<j>_std is the sample std of the
# perturbed values at the two saved dates, i.e. an estimate of the σsignal the run used (mean(σ²-draws)·κ of a base run
# whose σ² posterior is heavy-tailed, so it cannot be recomputed to better than ±20 %: tests inject it instead).
sig_names = ["filtered_means", "filtered_variances", "filtered_state_probs", "filtered_trans_probs", "forecasts"]
sig = {"source": "data/output/signals_official_noise_<noise>_allsignal/<name>_dispersion.csv", "K": 3, "noise_samples": 100,
       "signalburnin": 100000, "signalNrun": 250000,
       "note": "trans_a_b column holds A[b,a]; <col>_mean / <col>_std = mean / std over the 100 perturbed copies of the per-copy posterior mean",
       "cases": []}
for noise in ("0.1", "0.3", "0.6"):
    tabs = {}
    for n in sig_names:
        r = list(csv.reader(open(f"{REF}/data/output/signals_official_noise_{noise}_allsignal/{n}_dispersion.csv")))
        tabs[n] = (r[0], {x[0]: x for x in r[1:]})
    for idx in (121, 200, 300, 450, 570):
        d = dates[idx - 1]
        case = {"noise": float(noise), "end_index": idx, "date": d}
        for n in sig_names:
            head, rows_ = tabs[n]
            row = dict(zip(head, rows_[d]))
            cols = [c[:-5] for c in head if c.endswith("_mean") and not c.startswith("signal")]
            case[n] = {"columns": cols, "mean": [float(row[c + "_mean"]) for c in cols], "std": [float(row[c + "_std"]) for c in cols]}
        row = dict(zip(*[tabs["filtered_means"][0], tabs["filtered_means"][1][d]]))
        case["signal_std"] = [float(row["signal_1_std"]), float(row["signal_2_std"])]
        sig["cases"].append(case)
json.dump(sig, open(f"{HERE}/signals_allsignal_subset.json", "w"), indent=1)
print(len(sig["cases"]), "signal cases")

# the complete allsignal tables (every end date of the three noise levels), compact: used by scripts/signals_run_report.py
sig_all = {"source": sig["source"], "note": sig["note"], "noise": {}}
for noise in ("0.1", "0.3", "0.6"):
    tabs = {}
    for n in sig_names:
        r = list(csv.reader(open(f"{REF}/data/output/signals_official_noise_{noise}_allsignal/{n}_dispersion.csv")))
        tabs[n] = (r[0], r[1:])
    date_idx = {d: i + 1 for i, d in enumerate(dates)}
    block = {"end_index": [date_idx[x[0]] for x in tabs["filtered_means"][1]], "signal_std": []}
    for n in sig_names:
        head, body = tabs[n]
        assert [date_idx[x[0]] for x in body] == block["end_index"]
        cols = [c[:-5] for c in head if c.endswith("_mean") and not c.startswith("signal")]
        block[n] = {"columns": cols,
                    "mean": [[float(f"{float(x[head.index(c + '_mean')]):.6g}") for c in cols] for x in body],
                    "std": [[float(f"{float(x[head.index(c + '_std')]):.4g}") for c in cols] for x in body]}
    head, body = tabs["filtered_means"]
    block["signal_std"] = [[float(f"{float(x[head.index('signal_1_std')]):.5g}"), float(f"{float(x[head.index('signal_2_std')]):.5g}")] for x in body]
    sig_all["noise"][noise] = block
json.dump(sig_all, open(f"{HERE}/signals_allsignal_all.json", "w"), separators=(",", ":"))
print({k: len(v["end_index"]) for k, v in sig_all["noise"].items()}, "end dates in the full signal tables")

# ---- the published in-sample table (code/run_insamplefcasts.jl -> data/output/official_insample/forecats_insample.csv):
# one full-sample estimation (idx 1..576, burnin 20 000, Nrun 10 000), per date the 12-month forecast from that date's state
# probabilities and the probabilities themselves.  Its s1..s3 are FILTERED probabilities.
ins = list(csv.DictReader(open(f"{REF}/data/output/official_insample/forecats_insample.csv")))
assert [r["date"] for r in ins] == dates[:len(ins)]
r6 = lambda v: float(f"{float(v):.6g}")
json.dump({"source": "data/output/official_insample/forecats_insample.csv", "first_index": 1, "last_index": len(ins), "horizon": 12,
           "burnin": 20000, "nrun": 10000, "forecast": [r6(r["forecast"]) for r in ins],
           "probs": [[r6(r["s1"]), r6(r["s2"]), r6(r["s3"])] for r in ins]},
          open(f"{HERE}/official_insample.json", "w"), separators=(",", ":"))
print(len(ins), "in-sample rows")
