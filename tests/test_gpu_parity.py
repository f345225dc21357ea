"""GPU parity tests: the CUDA path (through the C ABI / the reference-shaped host API) against the CPU
oracle on the same seeded inputs, against the committed golden fixtures, and through size-independent
properties at larger sizes. Bit-exact where the reference's arithmetic order is defined (normalisation
multiply/divide, Welford moments, scaling); tolerances stated next to every other comparison."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import planted_counts

pytestmark = pytest.mark.gpu

SQRT_EPS = 1.5e-8  # Julia's default rtol for `≈` on Float64 (sqrt(eps)), used by the reference tests


def ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.spacing(np.maximum(np.abs(a), np.abs(b)))


# ---------------------------------------------------------------------------------------------
# matrices
# ---------------------------------------------------------------------------------------------
def test_upload_download_subset_slice_transpose(sv):
    rng = np.random.default_rng(0)
    X = sp.random(700, 90, 0.1, random_state=2, format="csc", data_rvs=lambda k: rng.integers(1, 50, k)).astype(np.int64)
    d = sv.DeviceMatrix.from_host(X)
    assert d.shape == X.shape and d.nnz == X.nnz
    back = d.to_host()
    assert (back != X).nnz == 0 and back.dtype == np.int64
    # Julia-style 1-based Int64 arrays
    dj = sv.DeviceMatrix.from_julia_arrays(700, 90, X.indptr.astype(np.int64) + 1, X.indices.astype(np.int64) + 1, X.data)
    assert (dj.to_host() != X).nnz == 0
    idx = np.array([5, 3, 3, 89, 0])
    assert (d.columns(idx).to_host() != X[:, idx]).nnz == 0
    assert (d.rows(100, 333).to_host() != X[100:333]).nnz == 0
    assert d.rows(10, 10).shape == (0, 90)
    T = d.transpose().to_host()
    assert (T != sp.csc_matrix(X.T)).nnz == 0 and T.has_sorted_indices
    # wide input takes the column-tile transpose path
    Xw = sp.random(40, 70000, 0.002, random_state=3, format="csc")
    Tw = sv.DeviceMatrix.from_host(Xw).transpose().to_host()
    assert abs(Tw - sp.csc_matrix(Xw.T)).max() == 0
    Tw.sort_indices()
    assert np.array_equal(Tw.indices, sp.csc_matrix(Xw.T).indices)
    # empty matrix
    E = sv.DeviceMatrix.from_host(sp.csc_matrix((5, 4), dtype=np.int64))
    assert E.nnz == 0 and E.to_host().nnz == 0


# ---------------------------------------------------------------------------------------------
# filtering.jl
# ---------------------------------------------------------------------------------------------
def test_filter_counts_fixed_matrix_and_random(sv, orc, golden):
    # test/test_input.jl:25-42 through the labelled interface
    X = sp.csc_matrix(np.array(golden["X"], dtype=np.int64))
    C = sv.filter_counts(sv.convert_counts(X), min_features=1, min_cells=2, min_umi=2)
    assert C.shape == (7, 4)
    assert C.names[0] == ["cell-1", "cell-2", "cell-3", "cell-4", "cell-5", "cell-8", "cell-9"]
    assert C.names[1] == ["gene-1", "gene-2", "gene-4", "gene-5"]
    assert (C.array != X[np.ix_(golden["filter_cells"], golden["filter_genes"])]).nnz == 0
    # random count matrices against the oracle, every threshold combination, explicit zeros kept, bit-exact
    rng = np.random.default_rng(3)
    D = rng.poisson(0.25, (4000, 700)).astype(np.int64)
    D[rng.integers(0, 4000, 50), :] = 0                      # empty cells
    D[:, rng.integers(0, 700, 20)] = 0                       # empty genes
    A = sp.csc_matrix(D)
    A.data[::97] = 0                                         # stored zeros: not "detected" (sum(>(0), ...)), but stored
    for kw in (dict(min_cells=3, min_features=150, min_feature_count=0, min_umi=170),
               dict(min_cells=0, min_features=0, min_feature_count=0, min_umi=0),
               dict(min_cells=40, min_features=5, min_feature_count=1, min_umi=0),
               dict(min_cells=10**6), dict(min_features=10**6)):
        G, CI, FI = sv.filter_counts(A, **kw)
        O, CIo, FIo = orc.filter_counts(A, **kw)
        np.testing.assert_array_equal(CI, CIo)
        np.testing.assert_array_equal(FI, FIo)
        assert G.shape == O.shape
        # every stored entry of a kept (cell, feature) pair survives, stored zeros included (Julia's A[CI, FI])
        coo = A.tocoo()
        assert G.nnz == int((CIo[coo.row] & FIo[coo.col]).sum())
        Gs, Os = G.copy(), O.copy()
        Gs.eliminate_zeros(); Os.eliminate_zeros()
        Gs.sort_indices(); Os.sort_indices()
        np.testing.assert_array_equal(Gs.indptr, Os.indptr)
        np.testing.assert_array_equal(Gs.indices, Os.indices)
        np.testing.assert_array_equal(Gs.data, Os.data)
    Gc, CIc = sv.filter_cells(A, min_features=150)
    Oc, CIoc = orc.filter_cells(A, min_features=150)
    np.testing.assert_array_equal(CIc, CIoc)
    assert (Gc != Oc).nnz == 0
    Gf, FIf = sv.filter_features(A, min_cells=60)
    Of, FIof = orc.filter_features(A, min_cells=60)
    np.testing.assert_array_equal(FIf, FIof)
    assert (Gf != Of).nnz == 0
    # device in -> device out, feeding the next stage without a host copy
    dA = sv.DeviceMatrix.from_host(sp.csc_matrix(D))
    dG, _, _ = sv.filter_counts(dA, min_cells=3, min_features=150)
    assert isinstance(dG, sv.DeviceMatrix)
    assert sv.normalize_cells(dG, scale_factor=1e4).shape == dG.shape


# ---------------------------------------------------------------------------------------------
# normalize.jl
# ---------------------------------------------------------------------------------------------
def test_normalize_cells_fixed_matrix(sv, golden):
    X = np.array(golden["X"], dtype=np.int64)[np.ix_(golden["filter_cells"], golden["filter_genes"])]
    C = sv.convert_counts(sp.csc_matrix(X))
    Q = sv.normalize_cells(C, method="relativecounts")
    assert isinstance(Q, sv.NamedArray) and Q.array.dtype == np.float64 and Q.names == C.names
    np.testing.assert_allclose(np.asarray(Q.array.sum(axis=1)).ravel(), 1.0, rtol=SQRT_EPS)
    np.testing.assert_allclose(Q.array.toarray(), np.array(golden["relcounts_dense"]), rtol=1e-15)
    Q10 = sv.normalize_cells(C, method="relativecounts", scale_factor=10.0)
    np.testing.assert_allclose(np.asarray(Q10.array.sum(axis=1)).ravel(), 10.0, rtol=SQRT_EPS)
    Q2 = sv.normalize_cells(C, method="lognormalize")
    np.testing.assert_allclose(Q2.array.toarray(), np.log1p(Q.array.toarray()), rtol=1e-15)
    for meth in ("relativecounts", "lognormalize"):
        assert sv.normalize_cells(C, method=meth, dtype=np.float32).array.dtype == np.float32
    with pytest.raises(ValueError, match="unknown normalization method"):
        sv.normalize_cells(C, method="nope")


def test_normalize_cells_bitwise_and_ulp(sv, orc):
    X = planted_counts(900, 1500, 8, seed=1)
    rel_gpu = sv.normalize_cells(X, method="relativecounts", scale_factor=1e4)
    rel_cpu = orc.normalize_cells(X, method="relativecounts", scale_factor=1e4)
    assert np.array_equal(rel_gpu.indices, rel_cpu.indices)
    np.testing.assert_array_equal(rel_gpu.data, rel_cpu.data)  # sf*x/s: one mul, one div -> bit-identical
    log_gpu = sv.normalize_cells(X, method="lognormalize", scale_factor=1e4)
    # log1p: CUDA, glibc and Julia are each <= 1 ulp but not identical; referee = long double log1p (H2)
    ref = np.log1p(rel_cpu.data.astype(np.longdouble)).astype(np.float64)
    assert ulp_diff(log_gpu.data, ref).max() <= 1.0
    f32 = sv.normalize_cells(X, method="relativecounts", scale_factor=1e4, dtype=np.float32)
    np.testing.assert_array_equal(f32.data, orc.normalize_cells(X, "relativecounts", 1e4, np.float32).data)


# ---------------------------------------------------------------------------------------------
# scaling.jl / variablefeatures.jl sweeps
# ---------------------------------------------------------------------------------------------
def test_mean_var_is_order_exact(sv, orc, golden):
    X = sp.csc_matrix(np.array(golden["X"], dtype=np.int64))
    mu, var = sv.mean_var(X)
    assert [float(v).hex() for v in mu] == golden["welford_mu_hex"]
    assert [float(v).hex() for v in var] == golden["welford_var_hex"]
    Xc = planted_counts(1500, 800, 6, seed=2)
    mu, var = sv.mean_var(Xc)
    mu_o, var_o = orc.mean_var(Xc)
    np.testing.assert_array_equal(mu, mu_o)
    np.testing.assert_array_equal(var, var_o)
    Y = orc.normalize_cells(Xc, "lognormalize", 1e4)
    mu, var = sv.mean_var(Y)
    mu_o, var_o = orc.mean_var(Y)
    np.testing.assert_array_equal(mu, mu_o)
    np.testing.assert_array_equal(var, var_o)
    Y32 = Y.astype(np.float32)
    mu, var = sv.mean_var(Y32)
    mu_o, var_o = orc.mean_var(Y32)
    assert mu.dtype == np.float32
    np.testing.assert_array_equal(mu, mu_o)
    np.testing.assert_array_equal(var, var_o)


def test_standardized_var_clipped(sv, orc):
    X = planted_counts(1200, 600, 6, seed=3)
    mu, sd = orc.mean_std(X)
    sd[7] = 0.0
    for vmax in (None, 2.0):
        got = sv.standardized_var_clipped(X, mu, sd, vmax)
        ref = orc.standardized_var_clipped(X, mu, sd, vmax)
        assert got[7] == 0.0
        assert ulp_diff(got, ref)[ref != 0].max() <= 1.0  # vs the long-double-accumulated value


def test_find_variable_features_sweeps_and_order(sv, orc):
    X = planted_counts(1000, 900, 8, seed=4)
    C = sv.convert_counts(X)
    hvf = sv.find_variable_features(C, 100)
    assert isinstance(hvf, sv.NamedArray) and len(hvf.array) == 100 and hvf.names[0][0] == f"gene-{hvf.array[0] + 1}"
    metric = sv.variance_stabilizing_transformation(X)
    assert np.all(np.diff(metric[hvf.array]) <= 0)  # decreasing metric (variablefeatures.jl:159)
    with pytest.raises(ValueError):
        sv.find_variable_features(C, 10, method="bogus")


def test_scale_features_fixed_matrix_and_bits(sv, orc, golden):
    X = sp.csc_matrix(np.array(golden["X"], dtype=np.int64))
    C = sv.convert_counts(X)
    S = sv.scale_features(C)
    assert S.names == C.names and S.shape == (10, 5)
    np.testing.assert_allclose(S.mu.array, golden["mu_over_std"], rtol=SQRT_EPS)
    assert [float(v).hex() for v in S.mu.array] == golden["welford_mu_over_sd_hex"]
    np.testing.assert_allclose(S.to_dense(), np.array(golden["scaled_dense"]), rtol=SQRT_EPS, atol=1e-15)
    assert sv.scale_features(C, dtype=np.float32).A.array.dtype == np.float32
    # random data: bit-identical to the oracle on identical inputs, with and without clipping / subset
    Xc = planted_counts(1300, 700, 6, seed=5)
    Y = orc.normalize_cells(Xc, "lognormalize", 1e4)
    feats = np.array([5, 100, 7, 650, 3])
    for kw in ({}, {"scale_max": 10.0}, {"scale_max": 0.5, "features": feats}):
        G = sv.scale_features(Y, **kw)
        O = orc.scale_features(Y, **kw)
        np.testing.assert_array_equal(G.A.data, O.P.data)
        np.testing.assert_array_equal(np.asarray(G.mu), O.mu)


# ---------------------------------------------------------------------------------------------
# CenteredMatrix products
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dense", [True, False])
def test_centered_matrix_fixed_vectors(sv, golden, dense):
    op = golden["op"]
    A = np.array(op["A"]) if dense else sp.csc_matrix(np.array(op["A"]))
    C = sv.CenteredMatrix(A, op["mu"])
    assert C.shape == (4, 3)
    Q = C.to_dense()
    r = np.array(op["r"])
    np.testing.assert_allclose(C @ r, Q @ r, rtol=SQRT_EPS, atol=1e-15)
    np.testing.assert_allclose(C @ r, op["Qr"], rtol=SQRT_EPS, atol=1e-15)
    y = np.array(op["y0"])
    np.testing.assert_allclose(C.mul(r, 2.0, 1.0, y), op["y"], rtol=SQRT_EPS)
    np.testing.assert_allclose(C.T @ y, op["Qty"], rtol=SQRT_EPS)
    r0 = r.copy()
    np.testing.assert_allclose(C.mul(y, 2.0, 1.0, r0, trans=True), op["r2"], rtol=SQRT_EPS)
    with pytest.raises(AssertionError):
        sv.CenteredMatrix(A, [0.1, 0.2])


def test_centered_matrix_random_vs_oracle(sv, orc):
    rng = np.random.default_rng(5)
    for (m, n, dens) in ((2000, 400, 0.1), (9000, 333, 0.03), (300, 70001, 0.001)):
        X = sp.random(m, n, dens, random_state=1, format="csc")
        mu = rng.standard_normal(n)
        G, O = sv.CenteredMatrix(X, mu), orc.CenteredMatrix(X, mu)
        v, w = rng.standard_normal(n), rng.standard_normal(m)
        scale = np.linalg.norm(O.mul(v))
        # summation order differs (parallel reduction vs serial scatter): 1e-13 of the vector norm
        assert np.linalg.norm(G @ v - O.mul(v)) <= 1e-13 * scale
        assert np.linalg.norm(G.T @ w - O.mul(w, trans=True)) <= 1e-13 * np.linalg.norm(O.mul(w, trans=True))
        y0 = rng.standard_normal(m)
        assert np.linalg.norm(G.mul(v, -0.5, 2.0, y0.copy()) - O.mul(v, -0.5, 2.0, y0.copy())) <= 1e-13 * (scale + np.linalg.norm(y0))
        # matrix forms (scaling.jl:259-272), adjoint form with the mathematically correct sign (T4)
        Vm, Wm = np.asfortranarray(rng.standard_normal((n, 3))), np.asfortranarray(rng.standard_normal((m, 3)))
        # SpMM kernels: 9 right-hand sides (3 passes of 4), alpha/beta epilogue, against the oracle's column loop
        V9, W9 = np.asfortranarray(rng.standard_normal((n, 9))), np.asfortranarray(rng.standard_normal((m, 9)))
        Y9, Z9 = np.asfortranarray(rng.standard_normal((m, 9))), np.asfortranarray(rng.standard_normal((n, 9)))
        ref = O.mul(V9, 1.5, -0.5, Y9.copy(order="F"))
        assert np.linalg.norm(G.mul(V9, 1.5, -0.5, Y9.copy(order="F")) - ref) <= 1e-12 * np.linalg.norm(ref)
        ref = O.mul(W9, 1.5, -0.5, Z9.copy(order="F"), trans=True)
        assert np.linalg.norm(G.mul(W9, 1.5, -0.5, Z9.copy(order="F"), trans=True) - ref) <= 1e-12 * np.linalg.norm(ref)
        D = O.to_dense() if m * n < 5e6 else None
        if D is not None:
            np.testing.assert_allclose(G @ Vm, D @ Vm, rtol=1e-10, atol=1e-10)
            np.testing.assert_allclose(G.T @ Wm, D.T @ Wm, rtol=1e-10, atol=1e-10)
        G.free()


def test_centered_matrix_adjoint_parent_and_dense(sv, orc):
    rng = np.random.default_rng(6)
    Xc = sp.random(300, 2500, 0.02, random_state=4, format="csc", data_rvs=lambda k: rng.poisson(10, k).astype(float))
    mu = np.asarray(Xc.mean(axis=1)).ravel()
    G = sv.CenteredMatrix(Xc.T, mu)  # lazy adjoint of a CSC, test_irlba.jl:111
    assert G.shape == (2500, 300)
    O = orc.CenteredMatrix(Xc, mu, transposed=True)
    v, w = rng.standard_normal(300), rng.standard_normal(2500)
    np.testing.assert_allclose(G @ v, O.mul(v), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(G.T @ w, O.mul(w, trans=True), rtol=1e-12, atol=1e-11)
    Xd = rng.standard_normal((500, 37))
    Gd = sv.CenteredMatrix(Xd, Xd.mean(axis=0))
    Q = Gd.to_dense()
    np.testing.assert_allclose(Gd @ v[:37], Q @ v[:37], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(Gd.T @ w[:500], Q.T @ w[:500], rtol=1e-12, atol=1e-11)


def test_operator_properties_at_scale(sv):
    # size-independent properties on a device-generated matrix: <S v, w> == <v, S' w>, linearity,
    # shard generation == row slice of the whole, product of shards == shard of product.
    counts = sv.synthetic_counts(200_000, 3000, 150.0, programs=20, seed=11)
    assert counts.shape == (200_000, 3000)
    nnz_per_cell = counts.nnz / 200_000
    assert 120 < nnz_per_cell < 180
    Y = sv.normalize_cells(counts, scale_factor=1e4)
    S = sv.scale_features(Y, scale_max=10.0)
    rng = np.random.default_rng(1)
    v, v2, w = rng.standard_normal(3000), rng.standard_normal(3000), rng.standard_normal(200_000)
    Sv, Stw = S @ v, S.T @ w
    assert abs(Sv @ w - v @ Stw) <= 1e-11 * np.linalg.norm(Sv) * np.linalg.norm(w)
    np.testing.assert_allclose(S @ (2.0 * v + v2), 2.0 * Sv + S @ v2, rtol=1e-10, atol=1e-9)
    # a centred, scaled gene has zero mean where nothing was clipped: S' 1 ~ 0 up to the clipping
    shard = sv.synthetic_counts(200_000, 3000, 150.0, programs=20, seed=11, rows=(50_000, 100_000))
    whole = counts.rows(50_000, 100_000)
    a, b = shard.to_host(), whole.to_host()
    assert (a != b).nnz == 0
    # determinism: same call, same bits
    np.testing.assert_array_equal(S @ v, Sv)
    np.testing.assert_array_equal(S.T @ w, Stw)


# ---------------------------------------------------------------------------------------------
# IRLBA: the reference's own scenarios (test/test_irlba.jl:25-115), same acceptance criteria
# ---------------------------------------------------------------------------------------------
def _relative_error(X, S, k=None):
    k = len(S.S) if k is None else k
    return np.linalg.norm(X - (S.U[:, :k] * S.S[:k]) @ S.Vt[:k])


def _dense_svd(X, k):
    U, s, Vt = np.linalg.svd(X, full_matrices=False)
    return U[:, :k], s[:k], Vt[:k]


@pytest.mark.parametrize("shape", [(50, 50), (100, 50)])
def test_irlba_dense(sv, shape):
    rng = np.random.default_rng(10)
    X = rng.standard_normal(shape)
    S = sv.irlba(X, 20, tol=1e-5, rng=rng)
    U, s, Vt = _dense_svd(X, 20)
    np.testing.assert_allclose(S.S, s, rtol=SQRT_EPS)
    assert np.linalg.norm(X.T @ S.U - S.V * S.S) / np.linalg.norm(X) < 1e-5
    np.testing.assert_allclose(_relative_error(X, S), np.linalg.norm(X - (U * s) @ Vt), rtol=SQRT_EPS)
    assert S.U.shape == (shape[0], 20) and S.Vt.shape == (20, shape[1])


def test_irlba_sparse_and_centered(sv):
    rng = np.random.default_rng(11)
    X = sp.random(2000, 400, 0.1, random_state=12, format="csc")
    Xd = X.toarray()
    S = sv.irlba(X, 2, tol=1e-9, rng=rng)
    U, s, Vt = _dense_svd(Xd, 2)
    np.testing.assert_allclose(S.S, s, rtol=SQRT_EPS)
    np.testing.assert_allclose(_relative_error(Xd, S), np.linalg.norm(Xd - (U * s) @ Vt), rtol=SQRT_EPS)
    assert np.linalg.norm(Xd.T @ S.U - S.V * S.S) / np.linalg.norm(Xd) < 1e-9
    C = sv.CenteredMatrix(X, np.asarray(X.mean(axis=0)).ravel())
    Q = C.to_dense()
    S = sv.irlba(C, 2, tol=1e-9, rng=rng)
    U, s, Vt = _dense_svd(Q, 2)
    np.testing.assert_allclose(S.S, s, rtol=SQRT_EPS)
    np.testing.assert_allclose(_relative_error(Q, S), np.linalg.norm(Q - (U * s) @ Vt), rtol=SQRT_EPS)
    assert np.linalg.norm(Q.T @ S.U - S.V * S.S) / np.linalg.norm(Q) < 1e-9
    Xs = rng.standard_normal((20, 10))
    C = sv.CenteredMatrix(Xs, Xs.mean(axis=0))
    Q = C.to_dense()
    S = sv.irlba(C, 3, rng=rng)
    np.testing.assert_allclose(S.S, _dense_svd(Q, 3)[1], rtol=SQRT_EPS)
    assert np.linalg.norm(Q.T @ S.U - S.V * S.S) / np.linalg.norm(Q) < 1e-9


def test_irlba_restart_signature(sv):
    # test_irlba.jl:52-63: the first solve is checked; the warm restart is @test_broken upstream, so
    # only the call signature / shapes are exercised here.
    rng = np.random.default_rng(12)
    X = rng.standard_normal((20, 10))
    s = np.linalg.svd(X, compute_uv=False)
    S = sv.irlba(X, 2, tol=1e-5, rng=rng)
    np.testing.assert_allclose(S.S, s[:2], rtol=SQRT_EPS)
    S3 = sv.irlba(X, 3, S, tol=1e-5, rng=rng)
    assert S3.U.shape == (20, 3) and S3.S.shape == (3,)
    # upstream marks these two @test_broken; with the start vector orthogonalised against the supplied V they hold
    np.testing.assert_allclose(S3.S, s[:3], rtol=SQRT_EPS)
    U, sd, Vt = np.linalg.svd(X, full_matrices=False)
    np.testing.assert_allclose(_relative_error(X, S3), np.linalg.norm(X - (U[:, :3] * sd[:3]) @ Vt[:3]), rtol=1e-6)


def test_irlba_tall_skinny_and_transpose(sv):
    rng = np.random.default_rng(13)
    X = rng.standard_normal((10000, 10))
    S1 = sv.svd_flip(sv.irlba(X, 2, tol=1e-9, rng=rng), u_based_decision=True)
    S2 = sv.svd_flip(sv.irlba(np.ascontiguousarray(X.T), 2, tol=1e-9, rng=rng), u_based_decision=False)
    np.testing.assert_allclose(S1.S, S2.S, rtol=SQRT_EPS)
    np.testing.assert_allclose(S1.U, S2.V, atol=1e-7)
    np.testing.assert_allclose(S1.V, S2.U, atol=1e-7)


def test_irlba_count_matrix_lazy_adjoint(sv):
    rng = np.random.default_rng(14)
    X = sp.random(3000, 10000, 0.01, random_state=15, format="csc", data_rvs=lambda k: rng.poisson(10, k).astype(float))
    mu = np.asarray(X.mean(axis=1)).ravel()
    C = sv.CenteredMatrix(X.T, mu)
    S = sv.irlba(C, 10, rng=rng)
    np.testing.assert_allclose(S.S, np.linalg.svd(C.to_dense(), compute_uv=False)[:10], rtol=SQRT_EPS)


def test_irlba_errors(sv):
    rng = np.random.default_rng(15)
    X = sp.random(500, 200, 0.05, random_state=1, format="csc")
    with pytest.raises(RuntimeError, match="convergence failed"):  # irlba.jl:73
        sv.irlba(X, 5, tol=1e-12, maxit=1, rng=rng)
    with pytest.raises(sv.SeveroB200Error):
        sv.irlba(X, 500, rng=rng)
    with pytest.raises(ValueError):
        sv.embedding(sv.CenteredMatrix(X, None), 5, method="tsne")
    # rank-deficient input exercises the Lanczos breakdown path (random restart vector, B entry 0)
    u = rng.standard_normal((300, 3))
    v = rng.standard_normal((40, 3))
    Xr = u @ v.T
    S = sv.irlba(Xr, 5, tol=1e-9, rng=rng)
    s = np.linalg.svd(Xr, compute_uv=False)
    np.testing.assert_allclose(S.S[:3], s[:3], rtol=1e-8)
    assert np.all(S.S[3:] < 1e-8 * s[0])


# ---------------------------------------------------------------------------------------------
# the whole path on a PBMC-3k-shaped synthetic input (BASELINE config 1) vs the oracle
# ---------------------------------------------------------------------------------------------
def test_pipeline_config1_against_oracle(sv, orc):
    nu = 50
    counts = sv.synthetic_counts(2700, 32738, 852.0, programs=50, seed=20260101)
    X = counts.to_host()
    assert X.shape == (2700, 32738) and 0.8 * 852 < X.nnz / 2700 < 1.2 * 852
    Y = sv.normalize_cells(counts, method="lognormalize", scale_factor=1e4)
    Yo = orc.normalize_cells(X, "lognormalize", 1e4)
    ref = np.log1p(orc.normalize_cells(X, "relativecounts", 1e4).data.astype(np.longdouble)).astype(np.float64)
    assert ulp_diff(Y.values(), ref).max() <= 1.0
    hvf = sv.find_variable_features(counts, 2000)
    mu_g, var_g = sv.mean_var(counts)
    mu_o, var_o = orc.mean_var(X)
    np.testing.assert_array_equal(mu_g, mu_o)
    np.testing.assert_array_equal(var_g, var_o)
    S = sv.scale_features(Y, scale_max=10.0, features=hvf)
    # oracle on the GPU's normalised values (isolates the scaling sweep from the 1-ulp log1p freedom)
    Yh = Y.to_host()
    So = orc.scale_features(Yh, scale_max=10.0, features=hvf)
    np.testing.assert_array_equal(S.A.values(), So.P.data)
    init = np.random.default_rng(3).standard_normal(2000)
    for tol in (1e-5, 1e-9):
        G = sv.irlba(S, nu, init=init, tol=tol)
        O = orc.irlba(So, nu, init=init, tol=tol)
        assert O.info == 0
        gap = O.S[nu - 1] / np.linalg.svd(So.to_dense(), compute_uv=False)[nu] if tol == 1e-9 else None
        np.testing.assert_allclose(G.S, O.S, rtol=1e-6)            # north_star: singular values rel 1e-6
        Q = So.to_dense()
        assert np.linalg.norm(Q.T @ G.U - G.V * G.S) / np.linalg.norm(Q) < tol   # test_irlba.jl:30 criterion
        if tol == 1e-9:
            # north_star: principal angle < 1e-4. A nu-dimensional subspace is only defined up to
            # residual/gap: at 2,700 cells the 50 planted programs (54 cells each) do not clear the noise
            # floor, sigma_50/sigma_51 ~ 1.001, so the angle is assessed on the leading well-separated block
            # (SURVEY trap T1) and the gap is stated.
            sd = np.linalg.svd(Q, compute_uv=False)
            ratios = sd[:nu] / sd[1:nu + 1]
            kstar = int(np.max(np.nonzero(ratios > 1.05)[0])) + 1
            print(f"config1: sigma_nu/sigma_nu+1 = {gap:.4f}; leading separated block k* = {kstar} (gap {ratios[kstar - 1]:.3f})")
            assert kstar >= 1
            assert orc.principal_angle(G.V[:, :kstar], O.V[:, :kstar]) < 1e-4
            assert orc.principal_angle(G.U[:, :kstar], O.U[:, :kstar]) < 1e-4
            # and both solvers' full nu-subspaces agree with the exact one to residual/gap
            Ud, _, Vtd = np.linalg.svd(Q, full_matrices=False)
            assert orc.principal_angle(G.V, Vtd[:nu].T) < 10 * tol * sd[0] / (sd[nu - 1] - sd[nu])
    em = sv.embedding(S, nu, method="pca", algorithm="irlba", init=init, tol=1e-9)
    Z, stdev, load = orc.pca_post(O.U, O.S, O.V, nu, 2700)
    np.testing.assert_allclose(em.stdev, stdev, rtol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(em.coordinates, axis=0), np.linalg.norm(Z, axis=0), rtol=1e-6)
    assert em.coordinates.shape == (2700, nu) and em.basis.shape == (2000, nu) and load.shape == (2000, nu)


def test_pipeline_planted_gap_subspace(sv, orc):
    # a shape where the planted programs clear the noise floor: the full nu-dimensional subspace is well
    # conditioned and must agree with the oracle to the north-star tolerances (s rel 1e-6, angle < 1e-4)
    nu = 12
    counts = sv.synthetic_counts(40_000, 4000, 400.0, programs=nu, seed=77)
    X = counts.to_host()
    Y = sv.normalize_cells(counts, method="lognormalize", scale_factor=1e4)
    hvf = sv.find_variable_features(counts, 1000)
    S = sv.scale_features(Y, scale_max=10.0, features=hvf)
    So = orc.scale_features(Y.to_host(), scale_max=10.0, features=hvf)
    np.testing.assert_array_equal(S.A.values(), So.P.data)
    init = np.random.default_rng(5).standard_normal(1000)
    G = sv.irlba(S, nu, init=init, tol=1e-9)
    O = orc.irlba(So, nu, init=init, tol=1e-9, parallel=True)
    sd = np.linalg.svd(So.to_dense(), compute_uv=False)
    gap = sd[nu - 1] / sd[nu]
    print(f"planted-gap case: sigma_nu/sigma_nu+1 = {gap:.3f}, restarts gpu/oracle = {G.iters}/{O.iters}")
    assert gap > 1.2
    np.testing.assert_allclose(G.S, sd[:nu], rtol=1e-6)
    np.testing.assert_allclose(G.S, O.S, rtol=1e-6)
    assert orc.principal_angle(G.V, O.V) < 1e-4
    assert orc.principal_angle(G.U, O.U) < 1e-4
    Gs = sv.svd_flip(G)
    Uo, Vo = orc.svd_flip(O.U, O.V)
    np.testing.assert_allclose(Gs.V, Vo, atol=1e-6)


def test_fp32_storage_mode(sv, orc):
    # optional Float32-storage / Float64-accumulate operator (north_star: singular values to rel 1e-4)
    import ctypes
    X = planted_counts(6000, 1200, 10, seed=9, mean_nnz=120)
    Y = orc.normalize_cells(X, "lognormalize", 1e4)
    So = orc.scale_features(Y, scale_max=10.0)
    C64 = sv.CenteredMatrix(So.P, So.mu)
    C32 = sv.CenteredMatrix(So.P, So.mu, storage="f32")
    vb, ib = ctypes.c_int(), ctypes.c_int()
    sv.lib().svb_operator_info(C32._operator(), None, None, None, None, ctypes.byref(vb), ctypes.byref(ib))
    assert (vb.value, ib.value) == (4, 2)
    rng = np.random.default_rng(2)
    v = rng.standard_normal(1200)
    ref = So.mul(v)
    assert np.linalg.norm(C32 @ v - ref) <= 1e-6 * np.linalg.norm(ref)      # fp32 rounding of the stored values
    assert np.linalg.norm(C64 @ v - ref) <= 1e-13 * np.linalg.norm(ref)
    init = rng.standard_normal(1200)
    S64 = sv.irlba(C64, 10, init=init, tol=1e-7)
    S32 = sv.irlba(C32, 10, init=init, tol=1e-7)
    np.testing.assert_allclose(S32.S, S64.S, rtol=1e-4)


_TWO_RANK = r"""
import os, sys, ctypes
import numpy as np, scipy.sparse as sp
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import severo_jl_b200 as sv
from severo_jl_b200 import sharding
from oracle import severo_oracle as orc
mode = sys.argv[2]
os.environ["SVB_P2P"] = "1" if mode == "p2p" else "0"
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
sv.init(local)
sharding.init_comm_from_torch()
rng = np.random.default_rng(0)
m, n, nu = 30011, 500, 8
X = sp.random(m, n, 0.05, random_state=3, format="csc")
u, v = rng.standard_normal((m, nu)), rng.standard_normal((n, nu))
X = sp.csc_matrix(X + sp.csc_matrix((u * np.linspace(3, 1, nu)) @ v.T * (rng.random((m, n)) < 0.02)))
mu = np.asarray(X.mean(axis=0)).ravel()
bounds = sharding.shard_bounds(m, world)
lo, hi = bounds[rank]
C = sv.CenteredMatrix(sp.csc_matrix(X[lo:hi]), mu)
init = rng.standard_normal(n)
S = sv.irlba(C, nu, init=init, tol=1e-9)
O = orc.irlba(orc.CenteredMatrix(X, mu), nu, init=init, tol=1e-9)
assert np.allclose(S.S, O.S, rtol=1e-6), (S.S, O.S)
U = sharding.gather_rows(S.U, bounds, rank)
assert orc.principal_angle(U, O.U) < 1e-4 and orc.principal_angle(S.V, O.V) < 1e-4
# sharded pre-processing: merged moments == moments of the whole matrix
cnt = sp.csc_matrix(rng.poisson(0.3, (4000, 60)).astype(np.int64))
b2 = sharding.shard_bounds(4000, world)
d = sv.DeviceMatrix.from_host(cnt[b2[rank][0]:b2[rank][1]])
mean, var, tot = sharding.merged_mean_var(d)
mo, vo = orc.mean_var(cnt)
assert tot == 4000 and np.allclose(mean, mo, rtol=1e-13) and np.allclose(var, vo, rtol=1e-12)
# order-exact mode: the Welford state is carried from rank to rank -> the bits of one sequential pass
me, ve, tot = sharding.exact_mean_var(d)
assert tot == 4000 and np.array_equal(me, mo) and np.array_equal(ve, vo)
# count-level operator on cell shards: moments over the cells of all ranks, solve == oracle on the whole matrix
lam = np.exp(rng.normal(-1.2, 1.0, 300))
Xc = rng.poisson(np.exp(0.3 * rng.standard_normal(6000))[:, None] * lam[None, :] * 3.0)
Xc[Xc.sum(axis=1) == 0, 0] = 1
Xc = sp.csc_matrix(Xc.astype(np.int64))
b3 = sharding.shard_bounds(6000, world)
Cc = sv.scale_features_counts(sv.DeviceMatrix.from_host(Xc[b3[rank][0]:b3[rank][1]]), scale_factor=1e4, scale_max=10.0, moments="fast")
Soc = orc.scale_features(orc.normalize_cells(Xc, "lognormalize", 1e4), scale_max=10.0)
assert np.allclose(Cc.mu, Soc.mu, rtol=1e-11)
init3 = rng.standard_normal(300)
Sc = sv.irlba(Cc, 6, init=init3, tol=1e-9)
Oc = orc.irlba(Soc, 6, init=init3, tol=1e-9)
assert np.allclose(Sc.S, Oc.S, rtol=1e-6), (Sc.S, Oc.S)
Uc = sharding.gather_rows(Sc.U, b3, rank)
assert orc.principal_angle(Uc, Oc.U) < 1e-4 and orc.principal_angle(Sc.V, Oc.V) < 1e-4
sv.lib().svb_comm_destroy()
dist.destroy_process_group()
print("OK", rank, mode)
"""


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_two_gpu_sharded_irlba(tmp_path, mode):
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "two_rank.py"
    script.write_text(_TWO_RANK)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29644" if mode == "nccl" else "29645", str(script), root, mode],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == 2


def test_welford_carry_across_row_slices(sv, orc):
    # single GPU: chaining svb_welford_carry over three cell ranges reproduces the one-pass Welford bit for bit
    import ctypes
    X = planted_counts(3000, 500, 5, seed=12)
    Y = orc.normalize_cells(X, "lognormalize", 1e4)
    for M in (X, Y):
        d = sv.DeviceMatrix.from_host(M)
        g = M.shape[1]
        count = (M.shape[0] - np.diff(sp.csc_matrix(M).indptr)).astype(np.int64)
        mu, s = np.zeros(g), np.zeros(g)
        for lo, hi in ((0, 1000), (1000, 1004), (1004, 3000)):
            part = d.rows(lo, hi)
            sv._lib.check(sv.lib().svb_welford_carry(part._h, sv._lib.ptr(count), sv._lib.ptr(mu), sv._lib.ptr(s)))
        mo, vo = orc.mean_var(M)
        assert np.all(count == M.shape[0])
        np.testing.assert_array_equal(mu, mo)
        np.testing.assert_array_equal(s / (M.shape[0] - 1.0), vo)


def test_embedding_default_algorithm_is_served(sv, orc):
    # docs/src/pbmc.md:134 calls embedding(S, 15, method=:pca) with the reference default algorithm=:arpack
    X = planted_counts(4000, 900, 9, seed=31, mean_nnz=100)
    Y = sv.normalize_cells(X, scale_factor=1e4)
    hvf = sv.find_variable_features(X, 400)            # also drops never-expressed genes (sigma = 0 -> NaN mu, trap T3)
    S = sv.scale_features(Y, scale_max=10.0, features=hvf)
    init = np.random.default_rng(1).standard_normal(400)
    em = sv.embedding(S, 9, method="pca", init=init)                         # default -> :arpack selector
    em2 = sv.embedding(S, 9, method="pca", algorithm="irlba", init=init, tol=1e-10)
    sd = np.linalg.svd(S.to_dense(), compute_uv=False)[:9]
    np.testing.assert_allclose(em.stdev * np.sqrt(3999), sd, rtol=1e-9)
    np.testing.assert_allclose(em.stdev, em2.stdev, rtol=1e-12)
    assert em.coordinates.shape == (4000, 9) and em.basis.shape == (400, 9)
    with pytest.raises(ValueError):
        sv.embedding(S, 9, algorithm="svd")      # dense LAPACK svd on the host: not served (no CPU fallback)
