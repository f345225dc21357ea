"""GPU edge cases: empty / ragged inputs, zero-variance genes (NaN propagation as upstream, trap T3), shapes
that do not align with the 32-cell blocks and 4096-cell tiles of the device layouts, tiny operators."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def test_empty_rows_columns_and_odd_shapes(sv, orc):
    rng = np.random.default_rng(0)
    for (m, n) in ((1, 1), (3, 2), (33, 7), (4097, 5), (8193, 31), (130, 300)):
        X = sp.random(m, n, 0.3, random_state=m + n, format="lil")
        if m > 2:
            X[1, :] = 0          # an empty cell
        if n > 2:
            X[:, 1] = 0          # an empty gene
        X = sp.csc_matrix(X)
        X.eliminate_zeros()
        mu = rng.standard_normal(n)
        G, O = sv.CenteredMatrix(X, mu), orc.CenteredMatrix(X, mu)
        v, w = rng.standard_normal(n), rng.standard_normal(m)
        np.testing.assert_allclose(G @ v, O.mul(v), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(G.T @ w, O.mul(w, trans=True), rtol=1e-12, atol=1e-12)
        d = sv.DeviceMatrix.from_host(X)
        assert (d.transpose().to_host() != sp.csc_matrix(X.T)).nnz == 0
        G.free()


def test_all_zero_matrix_and_single_entry(sv):
    Z = sp.csc_matrix((50, 8), dtype=np.float64)
    C = sv.CenteredMatrix(Z, np.arange(8.0))
    v = np.ones(8)
    np.testing.assert_allclose(C @ v, -np.full(50, 28.0))
    np.testing.assert_allclose(C.T @ np.ones(50), -50.0 * np.arange(8.0))
    E = sp.csc_matrix(([2.5], ([49], [7])), shape=(50, 8))
    C = sv.CenteredMatrix(E, None)
    y = C @ np.arange(8.0)
    assert y[49] == 17.5 and np.count_nonzero(y) == 1


def test_zero_variance_gene_propagates_nan_like_upstream(sv, orc):
    # scaling.jl:206-207: std == 0 -> mu/std = NaN (constant zero gene) or Inf; the reference does not guard it (T3)
    X = sp.csc_matrix(np.array([[1, 0, 2], [3, 0, 2], [0, 0, 2], [2, 0, 2]], dtype=np.int64))
    G = sv.scale_features(X)
    O = orc.scale_features(X)
    assert np.isnan(np.asarray(G.mu)[1]) and np.isnan(O.mu[1])
    np.testing.assert_array_equal(np.asarray(G.mu), O.mu)           # NaN == NaN positionally, Inf equal
    np.testing.assert_array_equal(G.A.data, O.P.data)
    mu, var = sv.mean_var(X)
    assert var[1] == 0.0 and var[2] == 0.0


def test_irlba_small_and_clamped_work(sv):
    rng = np.random.default_rng(3)
    # work = min(nu + 7, min(m, n)) clamps to the full dimension (irlba.jl:56-58): exact in one sweep
    X = rng.standard_normal((12, 6))
    S = sv.irlba(X, 4, rng=rng)
    np.testing.assert_allclose(S.S, np.linalg.svd(X, compute_uv=False)[:4], rtol=1.5e-8)
    X = sp.random(300, 9, 0.5, random_state=1, format="csc")
    S = sv.irlba(X, 9, rng=rng, tol=1e-9)
    np.testing.assert_allclose(S.S, np.linalg.svd(X.toarray(), compute_uv=False), rtol=1e-8)
    U, s, V = S
    np.testing.assert_allclose(U.T @ U, np.eye(9), atol=1e-8)
    # nu = 1
    S = sv.irlba(sp.random(500, 40, 0.2, random_state=2, format="csc"), 1, rng=rng, tol=1e-9)
    assert S.S.shape == (1,)


def test_reproducible_bits_run_to_run(sv):
    # all reductions are fixed-order: the same call returns the same bits
    rng = np.random.default_rng(4)
    X = sp.random(20000, 300, 0.05, random_state=5, format="csc")
    C = sv.CenteredMatrix(X, np.asarray(X.mean(axis=0)).ravel())
    init = rng.standard_normal(300)
    A = sv.irlba(C, 6, init=init, tol=1e-8)
    B = sv.irlba(C, 6, init=init, tol=1e-8)
    np.testing.assert_array_equal(A.S, B.S)
    np.testing.assert_array_equal(A.U, B.U)
    np.testing.assert_array_equal(A.Vt, B.Vt)


def test_warm_restart_is_correct_and_cheaper(sv):
    # irlba(A, nu, S::SVD) (irlba.jl:87-99): broken upstream (test_irlba.jl:60-62); here the supplied triplets are
    # kept and the new start vector is orthogonalised against them, so more PCs can be added incrementally
    rng = np.random.default_rng(21)
    X = rng.standard_normal((300, 120))
    s = np.linalg.svd(X, compute_uv=False)
    init = rng.standard_normal(120)
    S5 = sv.irlba(X, 5, init=init, tol=1e-8)
    S8 = sv.irlba(X, 8, S5, init=rng.standard_normal(120), tol=1e-8)
    cold = sv.irlba(X, 8, init=init, tol=1e-8)
    np.testing.assert_allclose(S8.S, s[:8], rtol=1.5e-8)
    assert np.linalg.norm(X.T @ S8.U - S8.V * S8.S) / np.linalg.norm(X) < 1e-8
    assert S8.mprod < cold.mprod


def test_remaining_abi_entry_points(sv):
    # svb_row_sums (exact int64 library sizes), svb_synth_normal, device info / version, profile + launch counters,
    # svb_mul_device
    import ctypes
    L = sv._lib
    lib = sv.lib()
    rng = np.random.default_rng(7)
    X = sp.random(5000, 300, 0.05, random_state=9, format="csc", data_rvs=lambda k: rng.integers(1, 2_000_000, k)).astype(np.int64)
    d = sv.DeviceMatrix.from_host(X)
    s = np.zeros(5000, dtype=np.int64)
    L.check(lib.svb_row_sums(d._h, L.ptr(s)))
    np.testing.assert_array_equal(s, np.asarray(X.sum(axis=1)).ravel())
    z = np.zeros(200_000)
    L.check(lib.svb_synth_normal(z.shape[0], 123, L.ptr(z)))
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    z2 = np.zeros(200_000)
    L.check(lib.svb_synth_normal(z2.shape[0], 123, L.ptr(z2)))
    np.testing.assert_array_equal(z, z2)
    sm, mem, ma, mi = ctypes.c_int(), ctypes.c_int64(), ctypes.c_int(), ctypes.c_int()
    L.check(lib.svb_device_info(ctypes.byref(sm), ctypes.byref(mem), ctypes.byref(ma), ctypes.byref(mi)))
    assert sm.value > 0 and mem.value > 0 and ma.value >= 9
    assert b"sm_100a" in lib.svb_version()
    C = sv.CenteredMatrix(sp.random(3000, 200, 0.1, random_state=1, format="csc"), rng.standard_normal(200))
    lib.svb_profile_enable(1)
    lib.svb_profile_reset()
    lib.svb_launch_count_reset()
    v = rng.standard_normal(200)
    y = C @ v
    C.T @ y
    ms, nl, by = (ctypes.c_double * 6)(), (ctypes.c_int64 * 6)(), (ctypes.c_double * 6)()
    lib.svb_profile_get(ms, nl, by)
    lib.svb_profile_enable(0)
    assert nl[0] == 1 and nl[1] == 2 and ms[0] > 0 and by[0] > 0 and lib.svb_launch_count() >= 3
    with pytest.raises(sv.SeveroB200Error):
        L.check(lib.svb_mul(C._operator(), b"N", 1.0, None, 0.0, None, 1))


def test_malformed_csc_is_refused_at_upload(sv):
    # round-1 advice: svb_csc_upload validates colptr / rowval (the kernels index with the rows directly)
    import ctypes
    L = sv._lib
    def upload(nrow, ncol, colptr, rowval, base=0):
        h = ctypes.c_void_p()
        cp, rv = np.asarray(colptr, dtype=np.int64), np.asarray(rowval, dtype=np.int64)
        nz = np.ones(max(len(rv), 1))
        return sv.lib().svb_csc_upload(nrow, ncol, L.ptr(cp), L.ptr(rv), L.SVB_I64, L.ptr(nz), L.SVB_F64, base, ctypes.byref(h)), h
    rc, h = upload(4, 3, [0, 2, 3, 4], [0, 3, 1, 2])
    assert rc == 0
    sv.lib().svb_matrix_free(h)
    assert upload(4, 3, [0, 2, 3, 4], [0, 4, 1, 2])[0] == L.SVB_EDIM        # row index == nrow
    assert upload(4, 3, [0, 2, 3, 4], [3, 0, 1, 2])[0] == L.SVB_EDIM        # rows descend inside a column
    assert upload(4, 3, [0, 2, 2, 4], [0, 0, 1, 2])[0] == L.SVB_EDIM        # duplicate row inside a column
    assert upload(4, 3, [0, 3, 2, 4], [0, 1, 2, 3])[0] == L.SVB_EDIM        # colptr decreases
    assert upload(4, 3, [1, 3, 4, 5], [0, 3, 1, 2], base=1)[0] == L.SVB_EDIM  # 1-based arrays with a 0 row
    assert b"svb_csc_upload" in sv.lib().svb_last_error()


def test_knn_rejects_non_finite_coordinates(sv):
    X = np.random.default_rng(0).standard_normal((300, 5))
    X[17, 2] = np.nan
    with pytest.raises(sv.SeveroB200Error):
        sv.nearest_neighbours(X, 5)


def test_pipelined_upload_moments_bit_identical(sv, orc):
    # svb_csc_upload_lognorm_moments (densest genes first, Welford chains on side streams during the upload) == the plain
    # sequence svb_csc_upload -> svb_normalize_libsize -> svb_mean_var, bit for bit; matrix identical; Julia-style arrays
    import ctypes
    from conftest import planted_counts
    L = sv._lib
    X = planted_counts(5000, 700, 6, seed=4, mean_nnz=90)
    hv = np.arange(0, 700, 3)
    chv = sp.csc_matrix(X[:, hv])
    m, n = chv.shape
    libsize = np.asarray(X.sum(axis=1)).ravel().astype(np.int64)
    colptr = chv.indptr.astype(np.int64) + 1
    rowval = chv.indices.astype(np.int64) + 1
    counts = chv.data.astype(np.int32)
    mean, var = np.zeros(n), np.zeros(n)
    h = ctypes.c_void_p()
    L.check(sv.lib().svb_csc_upload_lognorm_moments(m, n, L.ptr(colptr), L.ptr(rowval), L.SVB_I64, L.ptr(counts), 1, L.ptr(libsize), 1e4,
                                                    L.ptr(mean), L.ptr(var), ctypes.byref(h)))
    d = sv.DeviceMatrix(h)
    assert (d.to_host() != chv).nnz == 0
    Yo = orc.normalize_cells(X, "lognorm" if False else "lognormalize", 1e4)[:, hv]
    Yg = sv.normalize_cells(X, method="lognormalize", scale_factor=1e4)      # device log1p (the oracle's glibc log1p differs by <= 1 ulp)
    mu_ref, var_ref = sv.mean_var(sp.csc_matrix(Yg[:, hv]))
    np.testing.assert_array_equal(mean, mu_ref)
    np.testing.assert_array_equal(var, var_ref)
    np.testing.assert_allclose(mean, orc.mean_var(sp.csc_matrix(Yo))[0], rtol=1e-13)
    # 32-bit 0-based indices take the same path
    h2 = ctypes.c_void_p()
    m2, v2 = np.zeros(n), np.zeros(n)
    L.check(sv.lib().svb_csc_upload_lognorm_moments(m, n, L.ptr(chv.indptr.astype(np.int64)), L.ptr(chv.indices.astype(np.int32)), L.SVB_I32,
                                                    L.ptr(counts), 0, L.ptr(libsize), 1e4, L.ptr(m2), L.ptr(v2), ctypes.byref(h2)))
    np.testing.assert_array_equal(m2, mean)
    sv.lib().svb_matrix_free(h2)
    # malformed input is refused
    bad = rowval.copy()
    bad[3] = m + 5
    h3 = ctypes.c_void_p()
    rc = sv.lib().svb_csc_upload_lognorm_moments(m, n, L.ptr(colptr), L.ptr(bad), L.SVB_I64, L.ptr(counts), 1, L.ptr(libsize), 1e4,
                                                 L.ptr(mean), L.ptr(var), ctypes.byref(h3))
    assert rc == L.SVB_EDIM
    d.free()
