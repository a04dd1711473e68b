"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file <csv>):
python profiles/launch_summary.py gpurun_out/r2_launches.csv > profiles/r2_launches_summary.csv
Durations under ncu are serialised and cold-cache: the SHARES are what the bench line's sweep_kernel_share_of_step is checked against."""
import csv, io, re, sys
from collections import defaultdict

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}[unit]
    name = re.sub(r"\(.*\)$", "", r["Kernel Name"]).strip()
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print("kernel,launches,total_ms,avg_us,share")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k},{cnt[k]},{tot[k]:.3f},{1e3 * tot[k] / cnt[k]:.1f},{tot[k] / total:.4f}")
