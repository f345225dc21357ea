"""GPU: size-independent properties at BASELINE.json's full 1.3M-cell configuration (C3), where the CPU oracle
cannot run in test time: the reference's own acceptance criterion ||S'U - V Sigma|| / ||S|| < tol
(test/test_irlba.jl:30), orthonormality, the adjoint identity, Welford-carry == one-pass moments."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_c3_full_size_properties(sv):
    import bench
    cfg = bench.CONFIGS["C3"]
    B, mu, info = bench.build_workload(sv, cfg, 0, 1)
    m, n, nu = cfg["m"], cfg["n"], cfg["nu"]
    assert B.shape == (m, n) and info["Z_local"] > 2.0e9
    # scale_features output: every HVG column of the centred operator has zero mean up to the upper clip
    S = sv.CenteredMatrix(B, mu)
    rng = np.random.default_rng(0)
    v, w = rng.standard_normal(n), rng.standard_normal(m)
    Sv, Stw = S @ v, S.T @ w
    assert abs(Sv @ w - v @ Stw) <= 1e-11 * np.linalg.norm(Sv) * np.linalg.norm(w)      # <Sv, w> == <v, S'w>
    ones = S.T @ np.ones(m)
    col_norm = np.sqrt(m - 1.0)            # an unclipped standardised column has norm sqrt(m-1)
    assert np.max(np.abs(ones)) / (col_norm * np.sqrt(m)) < 0.05                      # clipping moves the mean only slightly
    init = rng.standard_normal(n)
    tol = 1e-5
    G = sv.irlba(S, nu, init=init, tol=tol)
    # ||S||_F^2 = sum of squared column norms = trace(S'S): estimate from the exact relation with the data:
    # columns are standardised (clipped): ||S||_F <= sqrt(n (m-1)); use the sharper sigma-based lower bound too
    StU = S.mul(np.asfortranarray(G.U), trans=True)                                      # n x nu through the SpMM kernels
    resid = np.linalg.norm(StU - G.V * G.S)
    S_fro_lower = np.linalg.norm(G.S)                                                    # ||S||_F >= ||sigma_1..nu||
    assert resid / S_fro_lower < tol                                                     # stricter than /||S||_F
    np.testing.assert_allclose(G.V.T @ G.V, np.eye(nu), atol=1e-8)
    np.testing.assert_allclose(G.U.T @ G.U, np.eye(nu), atol=1e-8)
    assert np.all(np.diff(G.S) <= 0) and G.S[0] > G.S[-1] > 0
    # sigma_i^2 = ||S v_i||^2 : forward products agree with the singular values
    SV = S.mul(np.asfortranarray(G.V[:, :4]))
    np.testing.assert_allclose(np.linalg.norm(SV, axis=0), G.S[:4], rtol=1e-6)
    # order-exact moments: chaining the Welford state over three cell ranges == one pass (bitwise), at full size
    mu1, var1 = sv.mean_var(B)
    count = (m - np.diff(B.colptr())).astype(np.int64)
    cm, cs = np.zeros(n), np.zeros(n)
    for lo, hi in ((0, 500_000), (500_000, 500_004), (500_004, m)):
        part = B.rows(lo, hi)
        sv._lib.check(sv.lib().svb_welford_carry(part._h, sv._lib.ptr(count), sv._lib.ptr(cm), sv._lib.ptr(cs)))
        part.free()
    np.testing.assert_array_equal(cm, mu1)
    np.testing.assert_array_equal(cs / (m - 1.0), var1)
    S.free()
    B.free()
