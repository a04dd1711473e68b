"""world_size-2 gloo tests (CPU) of the multi-rank host logic: window sharding (LPT on T_w, as bench.py and
hmcgpu_estimate_multi apply it) and the final gather of per-window summaries.  The GPU compute itself is covered by
test_gpu_parity.py::test_expanding_windows_ragged_batch_and_sharding_invariance (results independent of sharding)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import hmc_jl_b200 as H
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ws, we = H.expanding_windows(101, 160)
        T = we - ws + 1
        shard = H.shard_windows(T, world)[rank]
        # stand-in for the per-window summaries a rank computes: a deterministic function of the window
        local = np.stack([np.array([w, T[w], T[w] * 0.5, -w], dtype=np.float64) for w in shard])
        full = H.gather_window_summaries(local, shard, len(T), dist)
        dist.barrier()
        q.put((rank, None if full is None else full.tolist(), int(T[shard].sum())))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, full, load0), (r1, none, load1) = res
    assert r0 == 0 and r1 == 1 and none is None
    full = np.array(full)
    T = np.arange(101, 161)
    np.testing.assert_array_equal(full[:, 0], np.arange(60))       # every window arrived exactly once, in caller order
    np.testing.assert_array_equal(full[:, 1], T)
    np.testing.assert_allclose(full[:, 2], T * 0.5)
    assert abs(load0 - load1) <= T.max()                          # balanced by sum of window lengths


def _worker_skewed(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import hmc_jl_b200 as H
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T = np.array([1000, 2, 2, 2, 2, 2, 2])          # LPT by sum of T: one shard holds 1 window, the other 6
        shard = H.shard_windows(T, world)[rank]
        local = np.stack([np.array([w, T[w]], dtype=np.float64) for w in shard])
        full = H.gather_window_summaries(local, shard, len(T), dist)
        dist.barrier()
        q.put((rank, None if full is None else full.tolist(), len(shard)))
    finally:
        dist.destroy_process_group()


def test_gather_with_skewed_shards():
    """A shard of short windows holds more rows than ceil(n_windows / world) + 1 (round-1 advisor finding): the padded
    payload must be sized by the real maximum over the ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_skewed, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, full, n0), (_, _, n1) = res
    assert sorted((n0, n1)) == [1, 6]
    full = np.array(full)
    np.testing.assert_array_equal(full[:, 0], np.arange(7))
    np.testing.assert_array_equal(full[:, 1], [1000, 2, 2, 2, 2, 2, 2])


def test_single_process_gather_is_a_permutation():
    import hmc_jl_b200 as H
    shard = np.array([3, 0, 2, 1])
    local = np.arange(8.0).reshape(4, 2)
    out = H.gather_window_summaries(local, shard, 4)
    np.testing.assert_array_equal(out[shard], local)
    # one process holding every window in order: nothing is moved (no 45 MB copy per call on the wide batches)
    ident = H.gather_window_summaries(local, np.arange(4), 4)
    np.testing.assert_array_equal(ident, local)
    assert np.shares_memory(ident, local)
    # ... but a shard that is only a prefix of the windows still yields the full table
    part = H.gather_window_summaries(local[:2], np.arange(2), 4)
    assert part.shape == (4, 2)
    np.testing.assert_array_equal(part[:2], local[:2])


def test_bench_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """The driver launches `bench.py --impl reference --gpus N` under torchrun like the GPU arm: rank 0 alone runs the CPU port
    and prints the contract line, the other ranks exit 0 without work (no process group, no GPU needed)."""
    import json
    import socket
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--burnin", "2", "--nrun", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
