"""Per-instruction stall samples from an .ncu-rep source page: python profiles/ncu_hotspots.py file.ncu-rep [topN]
Groups samples by the stall reason and lists the hottest SASS instructions (needs -lineinfo builds, --import-source on)."""
import csv, subprocess, sys, io, collections
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
lines = out.splitlines()
# one table per kernel launch: take the first
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
tot = sum(int(r['# Samples'] or 0) for r in rows)
reasons = [k for k in rows[0] if k.startswith('stall_') and '(' not in k]
agg = collections.Counter()
for r in rows:
    for k in reasons:
        agg[k] += int(r[k] or 0)
print("total samples", tot, "instructions", len(rows), "executed warp-instr", sum(int(r['Instructions Executed'] or 0) for r in rows))
print("by reason:", ", ".join(f"{k[6:]}={v / max(1, sum(agg.values())):.3f}" for k, v in agg.most_common(8)))
rows_s = sorted(rows, key=lambda r: -int(r['# Samples'] or 0))
for r in rows_s[:top]:
    n = int(r['# Samples'] or 0)
    why = sorted(((int(r[k] or 0), k[6:]) for k in reasons), reverse=True)[:2]
    print(f"{n / tot:6.3%} {r['Address'][-6:]} {r['Source'][:70]:70s} {why}")
