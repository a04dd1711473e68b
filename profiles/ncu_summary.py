"""Prints the metrics we track from an .ncu-rep (run here, no GPU needed): python profiles/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max']
STALL = 'smsp__average_warps_issue_stalled_'
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:75s} {rows[1][i]:>8s} " + " ".join(r[i] for r in rows[2:]))
st = [(h, [float(r[hdr.index(h)]) for r in rows[2:]]) for h in hdr if h.startswith(STALL) and h.endswith('_per_issue_active.ratio')]
for h, v in sorted(st, key=lambda kv: -kv[1][0])[:9]:
    print(f"stall {h[len(STALL):-len('_per_issue_active.ratio')]:40s} " + " ".join(f"{x:.3f}" for x in v))
