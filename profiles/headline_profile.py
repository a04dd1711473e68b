"""Turns an `ncu --set full` capture of the headline sweep kernel into profiles/r2_headline_profile.json:
    python profiles/headline_profile.py <file.ncu-rep> <state-steps per captured launch> [source note] > profiles/r2_headline_profile.json
bench.py reads warp_inst_per_warp_step and dram_bytes_per_state_step from that file — and only when hot_source_hash equals the
hash of the kernel sources being run (hmc.jl_b200/build.py::hot_source_hash), so a stale capture can never feed the roofline."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
        "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def main():
    rep, steps = sys.argv[1], float(sys.argv[2])
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def col(name, scale=True):
        i = hdr.index(name)
        f = UNIT.get(units[i], 1.0) if scale else 1.0
        return [float(r[i].replace(",", "")) * f for r in data]

    inst = col("smsp__inst_executed.sum")
    dram = [a + b for a, b in zip(col("dram__bytes_read.sum"), col("dram__bytes_write.sum"))]
    dur = col("gpu__time_duration.sum")
    names = [r[hdr.index("Kernel Name")] for r in data]
    from hmc_jl_b200 import build
    n = len(data)
    prof = {
        "source": os.path.basename(rep) + (": " + note if note else ""),
        "hot_source_hash": build.hot_source_hash(),
        "kernel": names[0],
        "launches_captured": n,
        "state_steps_per_launch": steps,
        "warp_inst_per_warp_step": sum(inst) / n / (steps / 32.0),
        "dram_bytes_per_state_step": sum(dram) / n / steps,
        "per_launch": [{"inst_executed": a, "dram_bytes": b, "duration_ms": 1e3 * c} for a, b, c in zip(inst, dram, dur)],
        "issue_active_pct": col("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        "dram_throughput_pct": col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
        "registers_per_thread": col("launch__registers_per_thread", False)[0],
        "grid_size": col("launch__grid_size", False)[0],
    }
    print(json.dumps(prof, indent=1))


if __name__ == "__main__":
    main()
