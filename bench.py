#!/usr/bin/env python
"""bench.py — HMM state-steps/sec of the rolling-window Gibbs/FFBS estimation (BASELINE.json metric).

One "step" = one whole estimation pass over the batch: every window x chain runs burnin + nrun Gibbs sweeps
(draws, forward filter, backward sampling, forecasts h=1..12, per-window posterior summaries).
Workload at N=1 = BASELINE.json configs[1] (SURVEY §8d "C2"): one synthetic K=3 series of length 612 (numpy
default_rng(1234)), 500 expanding windows T_w = 101..600, 1000 burn-in + 1000 saved sweeps, 256 chains per window.
N>1 (torchrun, one rank per GPU): end dates are sharded over the ranks (LPT on T_w) and the chains per window scale
with N so per-GPU work is fixed (weak scaling); no data-path collective, only a final gather of per-window summaries.

  value : state-steps / device time of the sweeps with inputs resident in HBM (hmcgpu_plan_run; CUDA events on the
          library's stream, max over ranks).
  e2e   : the same through the public call with HOST buffers (hmcgpu_estimate: H2D of the series, all sweeps, D2H of
          the summaries, device allocation included), wall clock between barriers.
  --impl reference : the CPU oracle (port of the reference's algorithm; Julia is not installed, so the reference's own
          implementation cannot run) on all host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "HMM state-steps/sec (windows x chains x T per Gibbs sweep)"
UNIT = "state-steps/s"
K = 3
TRUTH = dict(A=np.array([[0.96, 0.02, 0.02], [0.02, 0.96, 0.02], [0.02, 0.02, 0.96]]),
             mu=np.array([1.6, 3.5, 8.3]), sigma2=np.array([0.8, 0.55, 7.9]))
HORIZONS = tuple(range(1, 13))


def synth_series(T=612, seed=1234):
    """generateData semantics (src/Hmc.jl:210-229) with numpy's default_rng(1234)."""
    rng = np.random.default_rng(seed)
    X = np.zeros(T, dtype=np.int64)
    for t in range(1, T):
        X[t] = rng.choice(K, p=TRUTH["A"][X[t - 1]])
    return TRUTH["mu"][X] + np.sqrt(TRUTH["sigma2"][X]) * rng.standard_normal(T)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def issue_roofline(warp_inst_per_warp_step, state_steps, sweep_seconds, sms, sm_mhz):
    """Warp-instruction issue roofline of the sweep kernel: nominal peak = SMs x 4 schedulers x the SM clock sampled under
    load; next to it the MEASURED rate of a pure-arithmetic kernel with the sweep's instruction blend
    (scripts/issue_peak.cu -> profiles/r1_issue_peak.json: FFMA / MUFU.EX2 / IMAD.WIDE / integer folds), scaled to that clock."""
    ach = warp_inst_per_warp_step * (state_steps / 32.0) / sweep_seconds
    pk = sms * 4 * sm_mhz * 1e6
    out = {"bound": "issue", "achieved": ach / 1e9, "peak": pk / 1e9, "unit": "Gwarp-inst/s", "frac": ach / pk,
           "warp_inst_per_warp_step": warp_inst_per_warp_step, "mufu_per_state_step": 4, "peak_source": "SMs x 4 x sampled SM clock"}
    try:
        m = json.load(open(os.path.join(ROOT, "profiles", "r1_issue_peak.json")))
        mixed = float(m["mix_gwarp_inst_s"]) * 1e9 * (sm_mhz * 1e6 * 4 * int(m["sms"])) / (float(m["nominal_issue_gwarp_inst_s_at_max_clock"]) * 1e9)
        out["measured_mixed_peak"] = mixed / 1e9
        out["frac_of_measured_mixed_peak"] = ach / mixed
    except Exception:
        pass
    return out


# Rank 0 prints ONE JSON line on stdout.  Libraries (NCCL's version banner, ...) write to file descriptor 1 directly, so
# the descriptor is pointed at stderr for the whole run and the line goes to a private duplicate of the real stdout.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def dist_setup(n_gpus):
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        # (NCCL prints its version banner on stdout from NCCL_DEBUG=VERSION upwards; stdout is redirected, see emit_line)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def barrier_sync(dist, local):
    import torch
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(local)


def all_max(dist, local, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(dist, local, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def cpu_baseline(y, ws, we, target_seconds=15.0, threads=0, burnin=1000, nrun=1000):
    """The oracle (kind "port") on all host cores over a bounded sample of the workload: every window with the
    workload's own burn-in + draws, and as many chains per window as fit ~target_seconds (calibrated on a short run)."""
    from oracle import oracle as O
    cores = threads or os.cpu_count() or 1
    T = (we - ws + 1).astype(np.int64)

    def jobs(burn, run, chains):
        out = []
        for c in range(chains):
            for w in range(len(T)):
                e = int(we[w])
                yf = [y[e - 1 + h] if e - 1 + h < len(y) else np.nan for h in HORIZONS]
                out.append(dict(y=y[int(ws[w]) - 1:e], K=K, burnin=burn, nrun=run, seed=1234, chain=w * 4096 + c,
                                horizons=HORIZONS, y_future=yf))
        return out

    t0 = time.perf_counter()
    O.gibbs_batch(jobs(16, 16, 1), n_threads=cores)
    cal = max(time.perf_counter() - t0, 1e-3)
    rate = T.sum() * 32 / cal
    sweeps = burnin + nrun
    chains = int(max(1, min(64, round(target_seconds * rate / (T.sum() * sweeps)))))
    if chains == 1 and T.sum() * sweeps / rate > 2 * target_seconds:     # slow host: fewer sweeps, same mix of windows
        sweeps = int(max(8, target_seconds * rate / T.sum()))
        burnin, nrun = sweeps // 2, sweeps - sweeps // 2
    t0 = time.perf_counter()
    _, used = O.gibbs_batch(jobs(burnin, nrun, chains), n_threads=cores)
    dt = time.perf_counter() - t0
    steps = int(T.sum()) * sweeps * chains
    return {"value": steps / dt, "unit": UNIT, "cores": used, "kind": "port",
            "sample": f"all {len(T)} windows (T=101..600) x {chains} chain(s), {burnin}+{nrun} sweeps, h=1..12 forecasts; "
                      f"{steps:.3e} state-steps in {dt:.1f}s; fp64 C port of src/Hmc.jl on {used} threads (Julia absent)"}, dt, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    y = synth_series()
    ws = np.ones(500, dtype=np.int32); we = np.arange(101, 601, dtype=np.int32)
    per = []
    base = None
    target = max(3.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        b, dt, steps = cpu_baseline(y, ws, we, target_seconds=target, burnin=args.burnin, nrun=args.nrun)
        if i >= args.warmup:
            per.append((dt, steps))
            base = b
    dt = sum(p[0] for p in per); steps = sum(p[1] for p in per)
    v = steps / dt
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / len(per), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus) | {"note": "reference arm = CPU port of the reference on host cores; "
                                                                   "Julia is not installed so the reference itself cannot run"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit_line(line)


def workload_config(args, world):
    if getattr(args, "workload", "c2") != "c2":
        return {"workload": {"c1": "C1: one window T=600, K=3, one chain (latency-bound by construction)",
                             "c4": "C4: independent series of T=2000, K=3, one chain each (wide batch)",
                             "c5": f"C5: independent series of T={args.length}, K={args.states}, one chain each",
                             "sig": "S1 noisy-signal Monte Carlo (estimatesignals!, src/Hmc.jl:868-914): 64 end dates x 100 perturbed "
                                    "copies, 12 signals after each end date (kappa = 1), smoothed state probabilities at the end date"}[args.workload],
                "K": getattr(args, "K_run", K), "chains": getattr(args, "chains_run", args.chains), "burnin": args.burnin, "nrun": args.nrun,
                "precision": f"fp{args.precision}", "l2": "pif spill working set exceeds the 126 MB L2; no flush needed"}
    return {"workload": "C2 rolling estimation: 500 expanding windows T=101..600 of one synthetic K=3 series (len 612, "
                        "default_rng(1234)), burnin+nrun Gibbs sweeps, forecasts h=1..12, per-window posterior summaries",
            "K": K, "windows": 500, "chains_per_window": args.chains * world, "burnin": args.burnin, "nrun": args.nrun,
            "precision": f"fp{args.precision}", "sharding": f"end dates over {world} rank(s) (LPT on T_w), no collective",
            "l2": "pif spill working set (sum_T*chains*12 B per rank) exceeds the 126 MB L2; no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=256, help="chains per window per GPU")
    ap.add_argument("--states", type=int, default=3, help="K for --workload c5 (K=3: SURVEY truth; otherwise mu_k=2k, sigma2=0.5)")
    ap.add_argument("--length", type=int, default=2000, help="T for --workload c4/c5")
    ap.add_argument("--workload", default="c2", choices=["c2", "c1", "c4", "c5", "sig"],
                    help="c2 (default, BASELINE configs[1]): 500 expanding windows; c1: one window T=600, one chain; "
                         "c4: --chains independent series (default 65536) of T=2000, K=3")
    ap.add_argument("--burnin", type=int, default=1000)
    ap.add_argument("--nrun", type=int, default=1000)
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--traffic", type=float, default=2.51e9,
                    help="dram__bytes_read+write per sweep-kernel launch from the ncu --set full capture "
                         "(profiles/r1_gibbs_sweeps_ncu_full.txt: one task group of 1000 warp tasks x 16 sweeps = 1.79e8 state-steps)")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import hmc_jl_b200 as H
    rank, world, local, dist = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    torch.zeros(1, device=f"cuda:{local}")          # create the primary context torch.cuda.synchronize() needs

    win_series = None
    sig_kw = {}
    if args.workload == "c2":
        y = synth_series()
        ws_all, we_all = H.expanding_windows(101, 600)
        n_chains = args.chains * world              # weak scaling: per-GPU work fixed
    elif args.workload == "c1":
        y = synth_series()
        ws_all, we_all = np.array([1], dtype=np.int32), np.array([600], dtype=np.int32)
        n_chains = 1
    elif args.workload == "sig":                    # signals_official_noise_* runs, many end dates per call
        y0 = synth_series()
        n_dates, n_copies, sig_len = 64, 100, 12
        ends = np.arange(600 - n_dates + 1 - sig_len, 600 - sig_len + 1)           # end dates; the window runs to end + 12
        rng = np.random.default_rng(1234)
        n_ser = n_dates * n_copies * world
        y = np.tile(y0, (n_ser + 1, 1))                                             # series 0 = the real data
        mask = np.zeros((n_ser + 1, len(y0)), dtype=np.uint8)
        we_all = np.tile(np.repeat(ends + sig_len, n_copies), world).astype(np.int32)
        for i in range(n_ser):
            e = we_all[i] - sig_len
            y[i + 1, e:e + sig_len] += rng.standard_normal(sig_len) * 0.8
            mask[i + 1, e:e + sig_len] = 1
        ws_all = np.ones(n_ser, dtype=np.int32)
        win_series = np.arange(1, n_ser + 1, dtype=np.int32)
        n_chains = args.chains if args.chains != 256 else 4
        sig_kw = dict(is_signal=mask, kappa=1.0, alpha=np.full(K, 2.0), nu=np.full(K, 2.0), pi_row_back=sig_len)
    else:                                           # c4 / c5: wide batch of independent series
        Kw = args.states if args.workload == "c5" else 3
        truth = TRUTH if Kw == 3 else dict(A=np.full((Kw, Kw), 0.1 / (Kw - 1)) + np.eye(Kw) * (0.9 - 0.1 / (Kw - 1)),
                                           mu=2.0 * np.arange(Kw), sigma2=np.full(Kw, 0.5))
        n_ser = (args.chains if args.chains != 256 else 65536) * world
        n_gen = min(n_ser, 2048)                    # distinct series generated; tiled to n_ser (device work is unaffected)
        L = args.length + 12
        rng = np.random.default_rng(1234)
        X = np.zeros((n_gen, L), dtype=np.int64)
        u = rng.random((n_gen, L))
        cum = np.cumsum(truth["A"], axis=1)
        for t in range(1, L):
            X[:, t] = np.minimum((u[:, t, None] > cum[X[:, t - 1]]).sum(1), Kw - 1)
        y = truth["mu"][X] + np.sqrt(truth["sigma2"][X]) * rng.standard_normal((n_gen, L))
        y = np.tile(y, ((n_ser + n_gen - 1) // n_gen, 1))[:n_ser]
        ws_all, we_all = np.ones(n_ser, dtype=np.int32), np.full(n_ser, args.length, dtype=np.int32)
        win_series = np.arange(n_ser, dtype=np.int32)
        n_chains = 1
        args.K_run = Kw
    args.chains_run = int(len(ws_all) * n_chains)      # chains in the whole job
    shard = H.shard_windows(we_all - ws_all + 1, world)[rank]
    ws, we = ws_all[shard], we_all[shard]
    spec = H.ProblemSpec(y, ws, we, K=getattr(args, "K_run", K), n_chains=n_chains, burnin=args.burnin, nrun=args.nrun, seed=1234, horizons=HORIZONS,
                         precision=args.precision, flags=H.FLAG_REF_Q1 | H.FLAG_SUMMARY, win_id=shard,
                         win_series=None if win_series is None else win_series[shard],
                         win_init_series=np.zeros(len(shard), dtype=np.int32) if sig_kw else None, **sig_kw)
    ctx = H.Context(local)
    plan = H.Plan(ctx, spec)
    for _ in range(args.warmup):
        plan.run()
    sampler = ClockSampler(local)
    barrier_sync(dist, local)
    sampler.start()
    t0 = time.perf_counter()
    dev_ms = sweep_ms = 0.0
    launches = sweep_launches = 0
    for _ in range(args.steps):
        plan.run()
        r = H.binding.Result()                      # timing fields only (no output pointers -> no D2H)
        ctx._check(ctx.L.hmcgpu_plan_fetch(plan.h, ctypes.byref(r)))
        dev_ms += r.gpu_ms; sweep_ms += r.sweep_kernel_ms
        launches += r.n_launches; sweep_launches += r.n_sweep_launches
        steps_per_run = r.state_steps
    barrier_sync(dist, local)
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    res = plan.fetch()
    plan.close()
    dev_s = all_max(dist, local, dev_ms / 1e3)
    total_steps = all_sum(dist, local, float(steps_per_run) * args.steps)
    value = total_steps / dev_s

    # ---- end to end through the public call with host buffers
    o = H.estimate(ctx, spec)                        # warm the allocator path and the gather's lazy NCCL connections once
    H.gather_window_summaries(np.concatenate([o.summary_mean, o.summary_var], axis=1), shard, len(we_all), dist,
                              device=f"cuda:{local}" if dist is not None else None)
    barrier_sync(dist, local)
    t0 = time.perf_counter()
    h2d = d2h = 0
    t_est = t_gat = 0.0
    for _ in range(args.steps):
        ta = time.perf_counter()
        o = H.estimate(ctx, spec)
        tb = time.perf_counter()
        h2d += o.h2d_bytes; d2h += o.d2h_bytes
        # final gather of per-window summaries on rank 0 (the only cross-rank step)
        H.gather_window_summaries(np.concatenate([o.summary_mean, o.summary_var], axis=1), shard, len(we_all), dist,
                                  device=f"cuda:{local}" if dist is not None else None)
        t_est += tb - ta; t_gat += time.perf_counter() - tb
    print(f"[bench rank {rank}] e2e per step: estimate {1e3 * t_est / args.steps:.1f} ms (device {o.gpu_ms:.1f} ms), "
          f"gather {1e3 * t_gat / args.steps:.1f} ms", file=sys.stderr, flush=True)
    barrier_sync(dist, local)
    e2e_s = all_max(dist, local, time.perf_counter() - t0)
    e2e_value = total_steps / e2e_s

    # ---- roofline of the dominant kernel (gibbs_sweeps_kernel): algorithmic bytes = (2K+2)*b per state-step
    bpe = 4 if args.precision == 32 else 8
    alg_bytes_per_step = (2 * getattr(args, "K_run", K) + 2) * bpe
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes_per_step * float(steps_per_run) * args.steps / (sweep_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": args.traffic, "kernel": "gibbs_sweeps_kernel", "peak_source": "MEASURED_PEAKS.json (measured copy)" if peaks else "fallback 6650",
                "algorithmic_bytes_per_state_step": alg_bytes_per_step, "launches_timed": sweep_launches,
                "avg_launch_ms": sweep_ms / max(1, sweep_launches),
                "sweep_kernel_share_of_step": sweep_ms / dev_ms,
                "note": "launches of the 4 task groups overlap on separate streams, so `achieved` uses the enclosing device time "
                        "of all sweep launches; real DRAM traffic is 14-21 B per state-step (< 32 algorithmic: y and part of pif "
                        "hit L1/L2) and the kernel is FP32/INT issue-bound (DESIGN.md section 6)"}

    # ---- second roofline (north_star: the slower of HBM traffic and FP32/MUFU issue): warp-instruction issue rate.
    # instructions per warp-step come from the ncu capture of the same kernel (smsp__inst_executed.sum / warp-steps,
    # profiles/r1_gibbs_sweeps_ncu_full_alltasks.txt); the peak is 4 schedulers x 1 warp-instr/clk x SMs at the SM clock
    # sampled under load during this run.
    wi = {3: 99.4}.get(getattr(args, "K_run", K)) if args.precision == 32 and args.workload == "c2" and len(ws) * n_chains > 16000 else None   # thread-per-chain kernel only
    issue = None
    if wi is not None and clocks.get("sm_mhz"):
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        issue = issue_roofline(wi, float(steps_per_run) * args.steps, sweep_ms / 1e3, sms, clocks["sm_mhz"])
    # ---- the reference's own shape of the same job (BASELINE configs[1] read literally): ONE chain per end date.  Such a
    # narrow batch runs on the time-parallel warp-per-chain kernel; reported next to the headline, not instead of it.
    literal = None
    if args.workload == "c2" and world == 1:
        spec1 = H.ProblemSpec(y, ws, we, K=K, n_chains=1, burnin=args.burnin, nrun=args.nrun, seed=1234, horizons=HORIZONS,
                              precision=args.precision, flags=H.FLAG_REF_Q1 | H.FLAG_SUMMARY)
        H.estimate(ctx, spec1)
        t0 = time.perf_counter()
        for _ in range(3):
            o1 = H.estimate(ctx, spec1)
        dt1 = (time.perf_counter() - t0) / 3
        literal = {"workload": "500 end dates x ONE chain (the reference's configuration), same sweeps and outputs",
                   "value": float(o1.state_steps) / dt1, "unit": UNIT, "ms_per_pass_e2e": 1e3 * dt1, "device_ms": o1.gpu_ms,
                   "kernel": "gibbs_scan_kernel (one warp per chain, time-parallel)"}
    if rank == 0:
        cb = None
        if not args.no_cpu_baseline and world == 1 and args.workload == "c2":
            cb, _, _ = cpu_baseline(y, ws_all, we_all, burnin=args.burnin, nrun=args.nrun)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if args.precision == 32 else "f64", "data": "synthetic", "config": workload_config(args, world),
                "wall_ms_per_step": 1e3 * wall / args.steps, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                        "ms_per_step": 1e3 * e2e_s / args.steps},
                "gpu_launches": int(launches), "roofline": roofline, "roofline_issue": issue, "one_chain_per_end_date": literal, "cpu_baseline": cb,
                "check": {"mu_mean_longest_window": res.summary_mean[int(np.argmax(we - ws))][0:getattr(args, "K_run", K)].tolist(),
                          "events": int(res.events)}}
        emit_line(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
