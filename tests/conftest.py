import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_fixed_inputs.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def sv():
    """The product package with the CUDA library initialised on cuda:0 — fails loudly without it."""
    import severo_jl_b200 as sv
    sv.init()
    return sv


@pytest.fixture(scope="session")
def orc():
    from oracle import severo_oracle
    return severo_oracle


def planted_counts(m, g, k_programs, seed, mean_nnz=60.0):
    """Small host-side Poisson count matrix with planted programs (tests only)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    p = np.exp(1.5 * rng.standard_normal(g))
    p /= p.sum()
    fold = np.where(rng.random((k_programs, g)) < 0.05, 6.0, 1.0)
    prog = rng.integers(0, k_programs, m)
    lib = np.exp(0.35 * rng.standard_normal(m))
    lam = lib[:, None] * p[None, :] * fold[prog] * mean_nnz * 3.0
    X = rng.poisson(lam)
    X[X.sum(axis=1) == 0, 0] = 1  # no empty cells (library size 0 divides by zero upstream too)
    return sp.csc_matrix(X.astype(np.int64))
