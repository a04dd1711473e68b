"""CPU tests of the host logic and of the C-ABI library as a binary artefact (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(H):
    header = open(os.path.join(ROOT, "include", "hmcgpu.h")).read()
    declared = set(re.findall(r"\b(hmcgpu_[a-z_]+)\s*\(", header))
    assert declared == set(H.SYMBOLS), declared ^ set(H.SYMBOLS)
    L = H.load()
    for s in declared:
        assert hasattr(L, s), s
    assert L.hmcgpu_version() >= 100


def test_struct_layout_matches_header(H):
    # field order/size of the ctypes mirrors = the C structs (8-byte pointers, natural alignment)
    from hmc_jl_b200 import binding as B
    assert ctypes.sizeof(B.Problem) == 176
    assert ctypes.sizeof(B.Result) == 10 * 8 + 2 * 8 + 5 * 8
    assert B.Problem.flags.offset == 172 and B.Problem.K.offset == 56


def test_no_device_fails_loudly(H):
    L = H.load()
    if L.hmcgpu_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(H.HmcGpuError) as e:
        H.Context(0)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The product path must not import, link, load or call anything under oracle/ (comments may mention it)."""
    pkg = os.path.join(ROOT, "hmc.jl_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|libhmc_oracle|hmc_oracle\.h|\borc_[a-z_]+\s*\(")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl", ".h")):
                assert not bad.search(open(os.path.join(dp, f)).read()), f


def test_shard_windows_lpt_balance(H):
    starts, ends = H.expanding_windows(101, 600)
    T = ends - starts + 1
    for n in (1, 2, 4, 8):
        shards = H.shard_windows(T, n)
        assert sorted(np.concatenate(shards).tolist()) == list(range(len(T)))
        loads = np.array([T[s].sum() for s in shards])
        assert loads.max() - loads.min() <= T.max()
        assert loads.max() <= 1.01 * T.sum() / n + T.max()


def test_estopt_mirror_defaults_and_validation(H):
    y = np.arange(200.0)
    o = H.EstOpt(y, None)
    assert (o.D, o.burnin, o.Nrun, o.seed, list(o.horizons), o.endIndex) == (3, 1000, 1000, 1234, [12], 121)
    assert list(o.sampleRange)[:2] == [1, 2] and len(o.sampleRange) == 121 and len(o.signalRange) == 0
    with pytest.raises(NotImplementedError):
        H.EstOpt(y, None, signalRange=range(100, 102))
    with pytest.raises(ValueError):
        H.EstOpt(y, None, sampleRange=range(1, 300))


def test_problem_spec_layouts(H):
    y = np.random.default_rng(0).normal(size=(3, 50))
    spec = H.ProblemSpec(y, [1, 5], [40, 50], K=3, n_chains=2, burnin=1, nrun=4, horizons=(1, 12),
                         flags=H.FLAG_DRAWS | H.FLAG_SUMMARY | H.FLAG_SMOOTHED_MEAN)
    o, res = spec.alloc_result()
    assert o.mu.shape == (2, 3, 8) and o.A.shape == (2, 3, 3, 8) and o.forecasts.shape == (2, 4, 8)
    assert o.summary_mean.shape == (2, 3 * 3 + 9 + 4 + 1)
    assert o.pib_mean.size == (40 + 46) * 3
    st = spec.struct()
    assert st.y_len == 50 and st.n_series == 3 and st.n_windows == 2 and st.n_h == 2
