"""GPU parity on the BASELINE.json configurations, counts -> PCs on BOTH sides independently.

GPU side: synthetic counts -> normalize_cells -> :vst metric -> top-n -> scale_features (explicit operator) and
scale_features_counts (count-level operator) -> irlba, all through the C ABI. Oracle side: the SAME host copy of the counts
through the oracle's own normalise (glibc log1p) / Welford / clipped variance / scale_data / IRLBA — nothing computed on the
GPU is fed to the oracle (VERDICT r1 "weak" #3). The loess between the two :vst sweeps is third-party and un-pinned upstream
(SURVEY 8c); both sides use the same deterministic parametric trend so that the selection itself can be compared.

Bars (BASELINE.json north_star; the reference's own criteria test/test_irlba.jl:29-30,113-114): singular values rel <= 1e-6;
principal angle < 1e-4 on the gapped leading block after a tol = 1e-9 solve; residual ||S'U - V Sigma|| / ||S|| < tol.

  C2        68,579 x 32,738, ~46 M nnz, 2,000 HVGs, nu = 50                  (full configuration)
  C4-shaped 65,536 cells, 5,000 HVGs, nu = 100, work = 107 (device SVD of B at its shared-memory limit) and work = 120
            (host small_svd fallback)                                          (C4's regime at a size the oracle can run)
  C3-sample the first 163,840 cells of C3 (the bench's cpu_baseline sample), 2,000 HVGs, nu = 50
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def trend(mu):
    """Deterministic stand-in for the loess fit of log10(sd) on log10(mu): Poisson-like sd = sqrt(mu (1 + 0.5 mu))."""
    return np.sqrt(mu * (1.0 + 0.5 * mu))


def oracle_pipeline(orc, X, n_hvg, scale_max=10.0):
    """counts -> (hvf, CenteredMatrix) entirely on the host (normalize.jl:17-55, variablefeatures.jl:19-50, scaling.jl:199-217)."""
    mu, var = orc.mean_var(X)
    sd = np.sqrt(var)
    expected = sd.copy()
    nc = sd > 0
    expected[nc] = trend(mu[nc])
    metric = orc.standardized_var_clipped(X, mu, expected)
    hvf = np.argsort(-metric, kind="stable")[:n_hvg]
    Y = orc.normalize_cells(X, "lognormalize", 1e4)
    return hvf, metric, orc.scale_features(Y, scale_max=scale_max, features=hvf)


def same_selection(hvf_g, hvf_o, metric_o):
    """The two selections agree as sets; a swap is only tolerated between genes whose metric ties at the cut to 1e-12."""
    sg, so = set(hvf_g.tolist()), set(hvf_o.tolist())
    if sg == so:
        return True
    cut = metric_o[hvf_o[-1]]
    return all(abs(metric_o[j] - cut) <= 1e-12 * abs(cut) for j in sg ^ so)


def fro_norm_centered(So):
    """||A - 1 mu'||_F of the oracle operator without densifying."""
    m = So.shape[0]
    A = So.P
    colsum = np.asarray(A.sum(axis=0)).ravel()
    return np.sqrt(A.data @ A.data - 2.0 * (So.mu @ colsum) + m * (So.mu @ So.mu))


def residual(So, U, s, V):
    """test/test_irlba.jl:30: ||S'U - V Sigma|| / ||S|| with the ORACLE's operator."""
    StU = So.mul(np.asfortranarray(U), trans=True, parallel=True)
    return np.linalg.norm(StU - V * s) / fro_norm_centered(So)


def gapped_block(s_ext, nu, ratio=1.02):
    """Largest k <= nu with sigma_k / sigma_{k+1} > ratio (s_ext holds at least nu+1 values)."""
    r = s_ext[:nu] / s_ext[1:nu + 1]
    idx = np.nonzero(r > ratio)[0]
    return (int(idx.max()) + 1, float(r[idx.max()])) if idx.size else (0, 1.0)


def check_solve(orc, G, O, So, nu, tol, s_ext=None, label=""):
    assert O.info == 0
    np.testing.assert_allclose(G.S, O.S, rtol=1e-6)
    res = residual(So, G.U, G.S, G.V)
    assert res < tol, (label, res)
    if s_ext is not None:
        k, gap = gapped_block(s_ext, nu)
        assert k >= 1
        av = orc.principal_angle(G.V[:, :k], O.V[:, :k])
        au = orc.principal_angle(G.U[:, :k], O.U[:, :k])
        print(f"{label}: residual {res:.2e}, gapped block k* = {k} (sigma_k/sigma_k+1 = {gap:.4f}), angle V {av:.2e} U {au:.2e}, "
              f"max rel dsigma {np.max(np.abs(G.S / O.S - 1)):.2e}, restarts gpu/oracle {G.iters}/{O.iters}")
        assert av < 1e-4 and au < 1e-4, (label, av, au)


def run_config(sv, orc, m, g, nnz, n_hvg, nu, programs, seed, works=(None,)):
    counts = sv.synthetic_counts(m, g, nnz, programs=programs, fold=6.0, seed=seed)
    X = counts.to_host()
    # ---- oracle, from the counts ----
    hvf_o, metric_o, So = oracle_pipeline(orc, X, n_hvg)
    # ---- GPU, from the counts ----
    hvf_g = sv.find_variable_features(counts, n_hvg, expected_std_fn=trend)
    assert same_selection(hvf_g, hvf_o, metric_o)
    hvf = hvf_o                                    # identical sets; one order for both sides so V rows line up
    Y = sv.normalize_cells(counts, method="lognormalize", scale_factor=1e4)
    S = sv.scale_features(Y, scale_max=10.0, features=hvf)
    Y.free()
    np.testing.assert_allclose(S.mu, So.mu, rtol=1e-12)   # log1p differs by <= 1 ulp between CUDA and glibc: not bitwise here
    C = sv.scale_features_counts(counts, scale_factor=1e4, scale_max=10.0, features=hvf)
    counts.free()
    init = np.random.default_rng(seed).standard_normal(n_hvg)
    orc.set_num_threads(os.cpu_count() or 1)
    # the spectrum just past nu (for the gap): one extra oracle solve with a few more values
    ext = orc.irlba(So, nu + 4, init=init, tol=1e-9, parallel=True)
    assert ext.info == 0
    for iw, work in enumerate(works):
        kw = {} if work is None else {"work": work}
        for tol in ((1e-5, 1e-9) if iw == 0 else (1e-9,)):
            O = orc.irlba(So, nu, init=init, tol=tol, parallel=True, **kw)
            for name, op in (("explicit", S), ("counts", C)):
                G = sv.irlba(op, nu, init=init, tol=tol, **kw)
                check_solve(orc, G, O, So, nu, tol, ext.S if tol == 1e-9 else None,
                            label=f"m={m} n={n_hvg} nu={nu} work={work} tol={tol} {name}")
    S.free()
    C.free()


def test_config2_full_pipeline_both_operators(sv, orc):
    # BASELINE.json configs[1]: PBMC-68k-shaped, 2,000 HVGs, 50 PCs
    import bench
    c = bench.CONFIGS["C2"]
    run_config(sv, orc, c["m"], c["g"], c["nnz"], c["n"], c["nu"], c["programs"], seed=20260102)


def test_config4_shaped_work107_and_host_svd_fallback(sv, orc):
    # C4's regime (5,000 HVGs, 100 PCs, work = 107: 8 short of bsvd_kernel's shared-memory capacity) at 65,536 cells, and a
    # work size beyond that capacity, where the host Jacobi of small_svd.cpp takes over
    run_config(sv, orc, 65_536, 12_000, 900.0, 5000, 100, 100, seed=20260104, works=(None, 120))


def test_config4_forced_host_svd_matches_device_svd(sv, orc):
    # the same solve with SVB_HOST_SVD=1 (host SVD of B at every restart) must give the device-SVD result
    counts = sv.synthetic_counts(20_000, 6000, 500.0, programs=40, fold=6.0, seed=5)
    hvf = sv.find_variable_features(counts, 1500, expected_std_fn=trend)
    C = sv.scale_features_counts(counts, scale_factor=1e4, scale_max=10.0, features=hvf)
    init = np.random.default_rng(5).standard_normal(1500)
    G = sv.irlba(C, 40, init=init, tol=1e-9)
    os.environ["SVB_HOST_SVD"] = "1"
    try:
        H = sv.irlba(C, 40, init=init, tol=1e-9)
    finally:
        del os.environ["SVB_HOST_SVD"]
    np.testing.assert_allclose(H.S, G.S, rtol=1e-9)
    assert orc.principal_angle(H.V[:, :20], G.V[:, :20]) < 1e-6
    C.free()
    counts.free()


def test_config3_sample_163840_cells_both_operators(sv, orc):
    # the bench's cpu_baseline sample: the first 163,840 cells of C3 through the bench's own workload builder
    import bench
    cfg = bench.CONFIGS["C3"]
    cells = 163_840
    B, mu, info = bench.build_workload(sv, cfg, 0, 1, rows_total=cells)
    chv, libsize = info["counts_hvg"], info["libsize"]
    n, nu = cfg["n"], cfg["nu"]
    So = orc.CenteredMatrix(B.to_host(), mu)
    S = sv.CenteredMatrix(B, mu)
    # the bench's e2e form of the count-level operator: moments = NULL (internal two-pass moments, as bench.make_counts_operator)
    Cx = sv.CountsCenteredMatrix(chv, libsize, 1e4, bench.SCALE_MAX)
    np.testing.assert_allclose(Cx.mu, mu, rtol=1e-11)
    init = np.random.default_rng(bench.SEED).standard_normal(n)
    orc.set_num_threads(os.cpu_count() or 1)
    ext = orc.irlba(So, nu + 4, init=init, tol=1e-9, parallel=True)
    for tol in (bench.TOL, 1e-9):
        O = orc.irlba(So, nu, init=init, tol=tol, parallel=True)
        for name, op in (("explicit", S), ("counts(two-pass moments)", Cx)):
            G = sv.irlba(op, nu, init=init, tol=tol)
            check_solve(orc, G, O, So, nu, tol, ext.S if tol == 1e-9 else None, label=f"C3 sample {cells} cells tol={tol} {name}")
    S.free()
    Cx.free()
