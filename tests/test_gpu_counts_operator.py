"""GPU parity tests of the count-level operator (csrc/factored.cu, svb_operator_create_counts): the operator of
scale_features(normalize_cells(X)[:, hvf]; scale_max) held over the RAW COUNTS — the scaled matrix of
scaling.jl:199-217 is never materialised. Checked against the CPU oracle's explicit scaled matrix (the reference's
own arithmetic) on the same inputs. Tolerance: every entry is t*(1/sd) instead of t/sd, i.e. within 2 ulp of the
reference's, so products agree to ~1e-15 relative (asserted at 1e-13 of the vector norm) and IRLBA results to the
north-star bars (sigma rel 1e-6, principal angle < 1e-4)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import planted_counts

pytestmark = pytest.mark.gpu


def _explicit_oracle(sv, orc, X, hvf, scale_max, sf=1e4):
    """The reference's explicit path on the GPU's normalised values (isolates the <=1-ulp log1p freedom, as the other
    parity tests do): oracle scale_features of Y[:, hvf]."""
    Y = sv.normalize_cells(X, method="lognormalize", scale_factor=sf)
    return orc.scale_features(Y, scale_max=scale_max, features=hvf)


def _check_products(C, So, rng, rtol=1e-13):
    m, n = C.shape
    v, w = rng.standard_normal(n), rng.standard_normal(m)
    ref = So.mul(v)
    assert np.linalg.norm(C @ v - ref) <= rtol * np.linalg.norm(ref)
    reft = So.mul(w, trans=True)
    assert np.linalg.norm(C.T @ w - reft) <= rtol * np.linalg.norm(reft)
    # mul!(y, S, v, alpha, beta) forms (scaling.jl:245-257)
    y0, z0 = rng.standard_normal(m), rng.standard_normal(n)
    got = C.mul(v, 2.0, 1.0, y0.copy())
    np.testing.assert_allclose(got, So.mul(v, 2.0, 1.0, y0.copy()), rtol=0, atol=rtol * np.linalg.norm(ref) * 4)
    got = C.mul(w, -0.5, 3.0, z0.copy(), trans=True)
    np.testing.assert_allclose(got, So.mul(w, -0.5, 3.0, z0.copy(), trans=True), rtol=0, atol=rtol * np.linalg.norm(reft) * 4)


@pytest.mark.parametrize("levels", [0, 4, 8, 16, 32])
def test_counts_operator_products_vs_oracle(sv, orc, levels):
    X = planted_counts(3000, 800, 8, seed=4, mean_nnz=150)   # counts reach the 30s: every level width sees exceptions
    hvf = sv.find_variable_features(X, 300)
    So = _explicit_oracle(sv, orc, X, hvf, 10.0)
    C = sv.scale_features_counts(X, scale_factor=1e4, scale_max=10.0, features=hvf, levels=levels)
    assert C.shape == So.shape
    np.testing.assert_array_equal(C.mu, So.mu)               # exact moments: the stored mean/sd is bit-identical
    info = C.info()
    assert info["nnz_coded"] + info["nnz_exception"] == So.P.nnz
    if levels:
        assert info["levels"] == levels
        big = int((X[:, hvf].data > levels).sum())
        assert info["nnz_exception"] >= big                    # counts above the table become exception chunks
    _check_products(C, So, np.random.default_rng(1))
    # matrix forms (scaling.jl:259-272)
    V = np.asfortranarray(np.random.default_rng(2).standard_normal((300, 3)))
    ref = So.mul(V)
    assert np.linalg.norm(C.mul(V) - ref) <= 1e-13 * np.linalg.norm(ref)
    C.free()


def test_counts_operator_clipped_entries_are_exceptions(sv, orc):
    # a tight scale_max clips many entries (scaling.jl:212): they must keep their exact clipped value
    X = planted_counts(2500, 600, 6, seed=11, mean_nnz=100)
    hvf = sv.find_variable_features(X, 200)
    So = _explicit_oracle(sv, orc, X, hvf, 0.5)
    cols = np.repeat(np.arange(200), np.diff(So.P.indptr))
    nclip = int((So.P.data == (0.5 + So.mu)[cols]).sum())
    assert nclip > 100
    C = sv.scale_features_counts(X, scale_factor=1e4, scale_max=0.5, features=hvf)
    assert C.info()["nnz_exception"] >= nclip
    _check_products(C, So, np.random.default_rng(3))


def test_counts_operator_fast_moments_and_irlba(sv, orc):
    nu = 12
    counts = sv.synthetic_counts(40_000, 4000, 400.0, programs=nu, seed=77)
    hvf = sv.find_variable_features(counts, 1000)
    Y = sv.normalize_cells(counts, method="lognormalize", scale_factor=1e4)
    So = orc.scale_features(Y.to_host(), scale_max=10.0, features=hvf)
    Cx = sv.scale_features_counts(counts, scale_factor=1e4, scale_max=10.0, features=hvf, moments="exact")
    Cf = sv.scale_features_counts(counts, scale_factor=1e4, scale_max=10.0, features=hvf, moments="fast")
    np.testing.assert_array_equal(Cx.mu, So.mu)
    np.testing.assert_allclose(Cf.mu, So.mu, rtol=1e-12)     # parallel two-pass moments vs sequential Welford
    rng = np.random.default_rng(5)
    _check_products(Cx, So, rng)
    _check_products(Cf, So, rng, rtol=1e-11)
    init = rng.standard_normal(1000)
    O = orc.irlba(So, nu, init=init, tol=1e-9, parallel=True)
    for C in (Cx, Cf):
        G = sv.irlba(C, nu, init=init, tol=1e-9)
        np.testing.assert_allclose(G.S, O.S, rtol=1e-6)       # north_star: singular values rel 1e-6
        assert orc.principal_angle(G.V, O.V) < 1e-4            # north_star: principal angle < 1e-4
        assert orc.principal_angle(G.U, O.U) < 1e-4
        Q = So.to_dense()
        assert np.linalg.norm(Q.T @ G.U - G.V * G.S) / np.linalg.norm(Q) < 1e-9   # test_irlba.jl:30 criterion
    em = sv.embedding(Cx, nu, method="pca", algorithm="irlba", init=init, tol=1e-9)
    _, stdev, _ = orc.pca_post(O.U, O.S, O.V, nu, 40_000)
    np.testing.assert_allclose(em.stdev, stdev, rtol=1e-6)


def test_counts_operator_edge_cases(sv, orc):
    rng = np.random.default_rng(8)
    # tiny input, empty cells inside the HVG columns, an explicit zero, counts far above every level
    D = rng.poisson(0.6, (37, 9)).astype(np.int64)
    D[5, :] = 0
    D[5, 0] = 3          # cell 5: library size 3, a single entry
    D[11, :] = 0
    D[11, 8] = 1
    D[20, 2] = 4000      # far above 32 levels -> explicit
    X = sp.csc_matrix(D)
    lib = np.asarray(D.sum(axis=1)).ravel()
    keep = lib > 0
    X = sp.csc_matrix(D[keep])
    hv = np.array([0, 2, 3, 5, 8])
    So = _explicit_oracle(sv, orc, X, hv, 10.0)
    for levels in (0, 4, 32):
        C = sv.scale_features_counts(X, scale_factor=1e4, scale_max=10.0, features=hv, levels=levels)
        np.testing.assert_array_equal(C.mu, So.mu)
        assert C.info()["nnz_exception"] >= 1
        _check_products(C, So, rng)
    # IRLBA through the operator on a small dense-ish problem (work clamps to min(m, n))
    G = sv.irlba(C, 3, init=rng.standard_normal(5), tol=1e-10)
    sd = np.linalg.svd(So.to_dense(), compute_uv=False)
    np.testing.assert_allclose(G.S, sd[:3], rtol=1e-8)
    # a constant gene has zero variance: the reference divides by zero (NaN mu), here the build refuses
    Z = X.tolil()
    Z[:, 1] = 0
    Z = sp.csc_matrix(Z)
    with pytest.raises(sv.SeveroB200Error, match="zero"):
        sv.scale_features_counts(Z, scale_factor=1e4, scale_max=10.0, features=np.array([0, 1, 2]))
    # float input is not a count matrix
    with pytest.raises(TypeError):
        sv.scale_features_counts(sp.csc_matrix(D[keep].astype(np.float64) + 0.5))
    with pytest.raises(sv.SeveroB200Error):
        dm = sv.DeviceMatrix.from_host(X)
        sv.CountsCenteredMatrix(dm, lib[keep], 1e4, 10.0, levels=5)


@pytest.mark.parametrize("n_genes", [4000, 9000])
def test_counts_operator_wide_gene_sets(sv, orc, n_genes):
    # 4000 genes: the gene vector takes the 512-thread forward kernel; 9000 genes: the 16-bit codes are indices
    # (byte offsets no longer fit) — both against the oracle's explicit matrix
    rng = np.random.default_rng(n_genes)
    m = 1500
    lam = np.exp(rng.normal(-2.2, 1.2, n_genes))
    D = rng.poisson(np.exp(0.4 * rng.standard_normal(m))[:, None] * lam[None, :] * 2.0)
    D[D.sum(axis=1) == 0, 0] = 1
    D[:, D.var(axis=0) == 0] += (np.arange(m) % 3 == 0)[:, None]   # no constant genes
    X = sp.csc_matrix(D.astype(np.int64))
    So = _explicit_oracle(sv, orc, X, None, 10.0)
    C = sv.scale_features_counts(X, scale_factor=1e4, scale_max=10.0)
    np.testing.assert_array_equal(C.mu, So.mu)
    _check_products(C, So, rng)
    init = rng.standard_normal(n_genes)
    G = sv.irlba(C, 5, init=init, tol=1e-9)
    O = orc.irlba(So, 5, init=init, tol=1e-9)
    np.testing.assert_allclose(G.S, O.S, rtol=1e-6)


def test_counts_operator_properties_at_scale(sv):
    # 200k cells x 2000 HVGs: adjoint identity <S x, w> = <x, S' w> and agreement with the explicit operator
    counts = sv.synthetic_counts(200_000, 8000, 700.0, programs=20, seed=5)
    hvf = sv.find_variable_features(counts, 2000)
    Y = sv.normalize_cells(counts, method="lognormalize", scale_factor=1e4)
    S = sv.scale_features(Y, scale_max=10.0, features=hvf)
    C = sv.scale_features_counts(counts, scale_factor=1e4, scale_max=10.0, features=hvf)
    info = C.info()
    assert info["nnz_coded"] + info["nnz_exception"] == S.A.nnz
    assert info["nnz_exception"] <= 0.02 * S.A.nnz
    # adjoint tiles fill whole rounds of the grid: 200,000 cells = 2 rounds of <= 1024-cell tiles, a multiple of 16 cells each
    assert info["tile_cells"] % 16 == 0 and 512 < info["tile_cells"] < 1024
    rng = np.random.default_rng(0)
    x, w = rng.standard_normal(2000), rng.standard_normal(200_000)
    Sx, Stw = C @ x, C.T @ w
    assert abs(np.dot(Sx, w) - np.dot(x, Stw)) <= 1e-12 * np.linalg.norm(Sx) * np.linalg.norm(w)
    ex, ext = S @ x, S.T @ w
    assert np.linalg.norm(Sx - ex) <= 1e-13 * np.linalg.norm(ex)
    assert np.linalg.norm(Stw - ext) <= 1e-13 * np.linalg.norm(ext)
    # run-to-run determinism (fixed-order reductions, no floating-point atomics)
    np.testing.assert_array_equal(C @ x, Sx)
    np.testing.assert_array_equal(C.T @ w, Stw)


@pytest.mark.parametrize("levels,log2r,want", [(16, 10, (2, 4)), (8, 11, (2, 3)), (32, 9, (2, 4)), (4, 12, (2, 2))])
def test_counts_operator_replica_tables(sv, orc, monkeypatch, levels, log2r, want):
    """Round 2: bank-shifted replicas of the gathered tables + build-time matching. SVB_FACT_LOG2R forces the one-CTA-per-SM
    adjoint kernel (R*L = 16384 table entries; levels 1-2 in as many replicas as the 227 KB of shared memory hold: `want` =
    (replicated levels, replicas)) that only the 1.3 M-cell configurations reach on their own; the forward replicas (four copies
    of x/sd for <= 2,046 genes) are always on. Products must equal the oracle's, and the assignment must leave (almost) no
    shared-memory conflicts (1.5-1.9 passes per set without it)."""
    X = planted_counts(9000, 900, 8, seed=21, mean_nnz=170)
    hvf = sv.find_variable_features(X, 400)
    So = _explicit_oracle(sv, orc, X, hvf, 10.0)
    monkeypatch.setenv("SVB_FACT_LOG2R", str(log2r))
    C = sv.scale_features_counts(X, scale_factor=1e4, scale_max=10.0, features=hvf, levels=levels)
    info = C.info()
    assert info["tile_cells"] == 1 << log2r and info["levels"] == levels
    assert info["fwd_replicas"] == 4 and (info["adj_replicated_levels"], info["adj_replicas"]) == want
    assert 1.0 <= info["fwd_passes_per_set"] < 1.3
    assert 1.0 <= info["adj_passes_per_set"] < 1.6   # (a 9,000-cell matrix: short segments, sets shared by several of them; 1.10-1.17 at C3)
    _check_products(C, So, np.random.default_rng(5))
    init = np.random.default_rng(6).standard_normal(400)
    G = sv.irlba(C, 8, init=init, tol=1e-9)
    O = orc.irlba(So, 8, init=init, tol=1e-9)
    np.testing.assert_allclose(G.S, O.S, rtol=1e-6)
    assert orc.principal_angle(G.V, O.V) < 1e-4
    C.free()


def test_counts_operator_forward_replicas_by_gene_count(sv, orc):
    # the number of x/sd replicas follows from the 16-bit byte-offset code space: 4 up to 2,046 genes, then 3, 2, 1
    for n_hvg, want in ((500, 4), (2300, 3), (3000, 2), (4500, 1)):
        X = planted_counts(1500, n_hvg + 200, 4, seed=n_hvg, mean_nnz=80)
        hvf = np.arange(n_hvg)
        keep = np.asarray((X[:, hvf] > 0).sum(axis=0)).ravel() > 1
        hvf = hvf[keep]
        C = sv.scale_features_counts(X, scale_factor=1e4, scale_max=10.0, features=hvf)
        nrep = C.info()["fwd_replicas"]
        n = len(hvf)
        stride = n + 1
        while stride % 16 != 5:
            stride += 1
        assert nrep == max(1, min(4, (8191 - n) // stride + 1)), (n, nrep)
        assert abs(n - n_hvg) > 50 or nrep == want
        So = _explicit_oracle(sv, orc, X, hvf, 10.0)
        _check_products(C, So, np.random.default_rng(n_hvg))
        C.free()
