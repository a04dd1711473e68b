import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_inflation():
    import csv
    rows = list(csv.DictReader(open(os.path.join(GOLDEN, "inflation_offic_inf.csv"))))
    return np.array([float(r["offic_inf"]) for r in rows]), [r["date"] for r in rows]


def synth_hmm(T, A, mu, sigma2, seed=1234):
    """generateData semantics (src/Hmc.jl:210-229): X_1 = 1, X_t ~ Cat(A[X_{t-1},:]), y_t ~ N(mu_x, sqrt(sigma2_x))."""
    rng = np.random.default_rng(seed)
    A, mu, sigma2 = map(np.asarray, (A, mu, sigma2))
    K = len(mu)
    X = np.zeros(T, dtype=np.int64)
    for t in range(1, T):
        X[t] = rng.choice(K, p=A[X[t - 1]])
    y = mu[X] + np.sqrt(sigma2[X]) * rng.standard_normal(T)
    return y, X + 1


K3_TRUTH = dict(A=np.array([[0.96, 0.02, 0.02], [0.02, 0.96, 0.02], [0.02, 0.02, 0.96]]),
                mu=np.array([1.6, 3.5, 8.3]), sigma2=np.array([0.8, 0.55, 7.9]))


def random_params(rng, B, K):
    A = rng.dirichlet(np.ones(K) * 2.0, size=(B, K))
    mu = np.sort(rng.normal(0, 3, size=(B, K)), axis=1)
    sigma2 = rng.uniform(0.3, 3.0, size=(B, K))
    rho = rng.dirichlet(np.ones(K), size=B)
    return A, mu, sigma2, rho


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def H():
    import hmc_jl_b200
    return hmc_jl_b200


@pytest.fixture(scope="session")
def ctx(H):
    c = H.Context(0)
    yield c
    c.close()


# ---- the reference's noisy-signal goldens (tests/golden/signals_allsignal_subset.json, made by tests/golden/make_golden.py)

def signals_golden_case(end_index, noise):
    """(y, dates, case, sigma_signal) of one end date x noise level of data/output/signals_official_noise_<noise>_allsignal.
    sigma_signal = the realised spread of the reference's perturbed values at its two saved dates."""
    import json
    y, dates = load_inflation()
    g = json.load(open(os.path.join(GOLDEN, "signals_allsignal_subset.json")))
    case = next(c for c in g["cases"] if c["end_index"] == end_index and c["noise"] == noise)
    assert dates[end_index - 1] == case["date"]
    return y, dates, case, float(np.sqrt(np.mean(np.square(case["signal_std"]))))


def dispersion_close(per_copy, gold, name, nse=5.0, atol=0.01):
    """per_copy: (copies, n) per-copy posterior means; gold: {"mean": [...], "std": [...]} over the reference's 100 copies."""
    gm, gs = np.array(gold["mean"]), np.array(gold["std"])
    m, sd = per_copy.mean(0), per_copy.std(0, ddof=1)
    se = np.sqrt(sd ** 2 / len(per_copy) + gs ** 2 / 100)
    assert np.all(np.abs(m - gm) <= nse * se + atol), (name, m, gm, se)
    # spread across copies: same scale.  The std of a std over 100 copies is ~7 %; ours also carries the Monte-Carlo error of
    # each copy's posterior mean (a few thousand draws here, 250 000 in the reference), which only widens it
    big = gs > 0.02
    assert np.all((sd[big] > 0.55 * gs[big]) & (sd[big] < 2.5 * gs[big])), (name, sd, gs)
