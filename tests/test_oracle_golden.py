"""CPU: pin the oracle against the reference's fixed-input tests (tests/golden) and its solver criteria."""
import numpy as np
import scipy.sparse as sp


def test_welford_bits_match_literal_transcription(orc, golden):
    X = sp.csc_matrix(np.array(golden["X"], dtype=np.int64))
    mu, var = orc.mean_var(X)
    assert [float(v).hex() for v in mu] == golden["welford_mu_hex"]
    assert [float(v).hex() for v in var] == golden["welford_var_hex"]
    # pure-python twin of the same loop
    for j in range(X.shape[1]):
        m2, v2 = orc.mean_var_py(X.data[X.indptr[j]:X.indptr[j + 1]], X.shape[0])
        assert float(m2).hex() == golden["welford_mu_hex"][j] and float(v2).hex() == golden["welford_var_hex"][j]


def test_scale_features_fixed_matrix(orc, golden):
    # test/test_scaling.jl:22-45
    X = sp.csc_matrix(np.array(golden["X"], dtype=np.int64))
    S = orc.scale_features(X)
    np.testing.assert_allclose(S.mu, golden["mu_over_std"], rtol=1.5e-8)
    np.testing.assert_allclose(S.to_dense(), np.array(golden["scaled_dense"]), rtol=1.5e-8, atol=1e-15)
    assert [float(v).hex() for v in S.mu] == golden["welford_mu_over_sd_hex"]


def test_scale_max_clip_rule(orc):
    # scaling.jl:209-212: upper clip at scale_max + mu/std on stored entries only
    X = sp.csc_matrix(np.array([[100, 1], [0, 1], [0, 2], [0, 0], [1, 0], [0, 0]], dtype=np.int64))
    S = orc.scale_features(X, scale_max=1.0)
    D = S.P.toarray()
    mu, sd = orc.mean_std(X)
    assert D[0, 0] == 1.0 + mu[0] / sd[0]
    assert D[4, 0] == 1.0 / sd[0]


def test_centered_operator_fixed_vectors(orc, golden):
    # test/test_scaling.jl:72-112
    op = golden["op"]
    for A in (sp.csc_matrix(np.array(op["A"])), np.array(op["A"])):
        C = orc.CenteredMatrix(A, op["mu"])
        assert C.shape == (4, 3)
        r = np.array(op["r"])
        np.testing.assert_allclose(C.mul(r), op["Qr"], rtol=1.5e-8, atol=1e-15)
        y = C.mul(r, 2.0, 1.0, np.array(op["y0"]))
        np.testing.assert_allclose(y, op["y"], rtol=1.5e-8)
        np.testing.assert_allclose(C.mul(y, trans=True), op["Qty"], rtol=1.5e-8)
        np.testing.assert_allclose(C.mul(y, 2.0, 1.0, r.copy(), trans=True), op["r2"], rtol=1.5e-8)


def test_filter_counts_fixed_matrix(orc, golden):
    # test/test_input.jl:25-42: filter_counts(min_features=1, min_cells=2, min_umi=2) keeps cells 1,2,3,4,5,8,9 and genes 1,2,4,5
    X = sp.csc_matrix(np.array(golden["X"], dtype=np.int64))
    C, CI, FI = orc.filter_counts(X, min_features=1, min_cells=2, min_umi=2)
    assert C.shape == (7, 4)
    assert list(np.nonzero(CI)[0]) == golden["filter_cells"] and list(np.nonzero(FI)[0]) == golden["filter_genes"]
    assert (C != X[np.ix_(golden["filter_cells"], golden["filter_genes"])]).nnz == 0
    # the order matters (filtering.jl:84): gene 3 is detected in 1 cell only once cell 5 ... stays, gene 2 in 2 cells
    C2, FI2 = orc.filter_features(X, min_cells=2)
    assert list(np.nonzero(FI2)[0]) == [0, 1, 3, 4]
    C3, CI3 = orc.filter_cells(X, min_features=1, min_feature_count=2)
    assert list(np.nonzero(CI3)[0]) == [0, 1, 2, 3, 4, 7, 8]     # cells whose largest count exceeds 2


def test_normalize_cells_fixed_matrix(orc, golden):
    # test/test_input.jl:47-75
    X = np.array(golden["X"], dtype=np.int64)[np.ix_(golden["filter_cells"], golden["filter_genes"])]
    A = sp.csc_matrix(X)
    Q = orc.normalize_cells(A, method="relativecounts")
    assert Q.dtype == np.float64
    np.testing.assert_allclose(np.asarray(Q.sum(axis=1)).ravel(), 1.0, rtol=1.5e-8)
    np.testing.assert_allclose(Q.toarray(), np.array(golden["relcounts_dense"]), rtol=1e-15)
    Q2 = orc.normalize_cells(A, method="lognormalize")
    np.testing.assert_array_equal(Q2.data, np.log1p(Q.data))
    Q10 = orc.normalize_cells(A, method="relativecounts", scale_factor=10.0)
    np.testing.assert_allclose(np.asarray(Q10.sum(axis=1)).ravel(), 10.0, rtol=1.5e-8)
    assert orc.normalize_cells(A, method="lognormalize", dtype=np.float32).dtype == np.float32
    # C build == pure-python twin, bit for bit
    np.testing.assert_array_equal(orc.row_norm_py(A, 1e4, True).data, orc.normalize_cells(A, "lognormalize", 1e4).data)


def test_stdvar_clipped_against_dense_formula(orc):
    rng = np.random.default_rng(3)
    X = sp.csc_matrix(rng.poisson(0.3, (500, 40)).astype(np.int64))
    mu, sd = orc.mean_std(X)
    sd[5] = 0.0
    out = orc.standardized_var_clipped(X, mu, sd)
    D = X.toarray().astype(float)
    vmax = np.sqrt(500)
    for j in range(40):
        if sd[j] == 0:
            assert out[j] == 0.0
            continue
        z = np.minimum((D[:, j] - mu[j]) / sd[j], vmax)
        np.testing.assert_allclose(out[j], np.sum(z * z) / 499, rtol=1e-13)


def _rel_err(X, U, s, V):
    return np.linalg.norm(X - (U * s) @ V.T)


def test_irlba_reference_scenarios(orc):
    # the eight scenarios of test/test_irlba.jl:25-115 with seeded inputs, same acceptance criteria
    rng = np.random.default_rng(7)
    for shape in ((50, 50), (100, 50)):
        X = rng.standard_normal(shape)
        R = orc.irlba(X, 20, tol=1e-5, rng=rng)
        sv = np.linalg.svd(X, compute_uv=False)
        assert R.info == 0
        np.testing.assert_allclose(R.S, sv[:20], rtol=1.5e-8)
        assert np.linalg.norm(X.T @ R.U - R.V * R.S) / np.linalg.norm(X) < 1e-5
        Ud, sd_, Vtd = np.linalg.svd(X, full_matrices=False)
        np.testing.assert_allclose(_rel_err(X, R.U, R.S, R.V), _rel_err(X, Ud[:, :20], sd_[:20], Vtd[:20].T), rtol=1.5e-8)
    Xs = sp.random(2000, 400, 0.1, random_state=11, format="csc")
    R = orc.irlba(Xs, 2, tol=1e-9, rng=rng)
    sv = np.linalg.svd(Xs.toarray(), compute_uv=False)
    np.testing.assert_allclose(R.S, sv[:2], rtol=1.5e-8)
    assert np.linalg.norm(Xs.T @ R.U - R.V * R.S) / np.linalg.norm(Xs.toarray()) < 1e-9
    # centred dense / sparse
    Xd = rng.standard_normal((20, 10))
    C = orc.CenteredMatrix(Xd, Xd.mean(axis=0))
    Q = C.to_dense()
    R = orc.irlba(C, 3, rng=rng)
    np.testing.assert_allclose(R.S, np.linalg.svd(Q, compute_uv=False)[:3], rtol=1.5e-8)
    assert np.linalg.norm(Q.T @ R.U - R.V * R.S) / np.linalg.norm(Q) < 1e-9
    C = orc.CenteredMatrix(Xs, np.asarray(Xs.mean(axis=0)).ravel())
    Q = C.to_dense()
    R = orc.irlba(C, 2, tol=1e-9, rng=rng)
    np.testing.assert_allclose(R.S, np.linalg.svd(Q, compute_uv=False)[:2], rtol=1.5e-8)
    assert np.linalg.norm(Q.T @ R.U - R.V * R.S) / np.linalg.norm(Q) < 1e-9
    # tall-skinny and its transpose with svd_flip
    Xt = rng.standard_normal((10000, 10))
    R1 = orc.irlba(Xt, 2, tol=1e-9, rng=rng)
    R2 = orc.irlba(np.ascontiguousarray(Xt.T), 2, tol=1e-9, rng=rng)
    U1, V1 = orc.svd_flip(R1.U, R1.V, True)
    U2, V2 = orc.svd_flip(R2.U, R2.V, False)
    np.testing.assert_allclose(R1.S, R2.S, rtol=1.5e-8)
    np.testing.assert_allclose(U1, V2, atol=1e-7)
    np.testing.assert_allclose(V1, U2, atol=1e-7)
    # count matrix through the lazy adjoint (test_irlba.jl:107-115)
    Xc = sp.random(300, 1000, 0.05, random_state=5, format="csc", data_rvs=lambda k: rng.poisson(10, k).astype(float))
    C = orc.CenteredMatrix(Xc, np.asarray(Xc.mean(axis=1)).ravel(), transposed=True)
    R = orc.irlba(C, 10, rng=rng)
    np.testing.assert_allclose(R.S, np.linalg.svd(C.to_dense(), compute_uv=False)[:10], rtol=1.5e-8)


def test_parallel_forward_product_is_bit_identical(orc):
    rng = np.random.default_rng(2)
    Xs = sp.random(3000, 200, 0.05, random_state=1, format="csc")
    C = orc.CenteredMatrix(Xs, rng.standard_normal(200))
    v = rng.standard_normal(200)
    np.testing.assert_array_equal(C.mul(v), C.mul(v, parallel=True))


def test_gram_and_tssvd_against_dense(orc, golden):
    # scaling.jl:274-296 C'C and embedding.jl:30-44 tssvd have no reference test: pinned on convert(Matrix, S)
    # (scaling.jl:298-303) and the dense SVD, like test_irlba.jl pins the solver.
    op = golden["op"]
    Cs = orc.CenteredMatrix(sp.csc_matrix(np.array(op["A"])), op["mu"])
    Q = np.array(op["A"]) - np.array(op["mu"])[None, :]
    np.testing.assert_allclose(orc.gram(Cs), Q.T @ Q, rtol=1e-14, atol=1e-15)
    rng = np.random.default_rng(11)
    X = sp.random(700, 90, 0.08, random_state=5, format="csc")
    mu = rng.standard_normal(90)
    for C in (orc.CenteredMatrix(X, mu), orc.CenteredMatrix(X, None), orc.CenteredMatrix(X.toarray(), mu),
              orc.CenteredMatrix(sp.csc_matrix(X.T), mu, transposed=True)):
        D = C.to_dense()
        G = orc.gram(C)
        assert np.abs(G - D.T @ D).max() <= 1e-13 * np.abs(G).max()
    C = orc.CenteredMatrix(X, np.asarray(X.mean(axis=0)).ravel())
    T = orc.tssvd(C, 6)
    U, s, Vt = np.linalg.svd(C.to_dense(), full_matrices=False)
    np.testing.assert_allclose(T.S, s[:6], rtol=1e-12)
    assert orc.principal_angle(T.V, Vt[:6].T) < 1e-8 and orc.principal_angle(T.U, U[:, :6]) < 1e-8


def test_knn_restatement_meets_the_reference_criterion(orc):
    # test/test_nn.jl:22-56: Jaccard overlap with partialsortperm of the pairwise distances, 30 % quantile == 1.0. The oracle is
    # that exact search (k nearest by Distances.Euclidean / CosineDist), checked here against scipy's pairwise distances.
    from scipy.spatial.distance import cdist
    X = np.random.default_rng(3).random((100, 10))
    for metric in ("euclidean", "cosine"):
        D = cdist(X, X, metric)
        idx, dist = orc.knn(X, 4, metric, include_self=True)
        j = []
        for i in range(100):
            nn = set(np.argsort(D[i], kind="stable")[:4])
            x = len(nn & set(idx[i]))
            j.append(x / (4 + (4 - x)))
        assert np.quantile(j, 0.3) == 1.0 and min(j) == 1.0
        np.testing.assert_allclose(dist, np.take_along_axis(D, idx.astype(np.int64), axis=1), atol=1e-14)
        idx2, _ = orc.knn(X, 4, metric, include_self=False)
        assert not np.any(idx2 == np.arange(100)[:, None])
    nn = orc.nearest_neighbours(X, 4)
    assert nn.shape == (100, 100) and nn.nnz == 400 and np.all(np.asarray(nn.sum(axis=0)).ravel() == 4)   # k rows per cell column


def _device_snn_algorithm(indptr, indices, n, k, prune, T):
    """Pure-Python twin of csrc/snn.cu: reverse lists, a pair (i, j) emitted only at its smallest common neighbour (with an
    early exit on a common element below it), survivors rank-sorted inside the column. Validates the ALGORITHM on the CPU."""
    N = [list(indices[indptr[i]:indptr[i + 1]]) for i in range(n)]
    R = [[] for _ in range(n)]
    for j in range(n):
        for p in N[j]:
            R[p].append(j)
    R = [r[::-1] for r in R]                     # any order must do
    cols = []
    for j in range(n):
        kk = T(k if k is not None else len(N[j]))
        surv = []
        for p in N[j]:
            for i in R[p]:
                c, first = 0, True
                for q in N[i]:
                    if q in N[j]:
                        if q < p:
                            first = False
                            break
                        c += 1
                if first:
                    x = T(c)
                    v = x / (kk + (kk - x))
                    if not abs(v) <= T(prune):
                        surv.append((i, v))
        keys = [s[0] for s in surv]
        assert len(set(keys)) == len(keys)       # every pair is emitted exactly once
        out = [None] * len(surv)
        for (i, v) in surv:
            out[sum(1 for u in keys if u < i)] = (i, v)
        cols.append(out)
    return cols


def test_jaccard_index_restatement(orc):
    # neighbours.jl:88-110. Hand-computed case: N(0)={0,1,2}, N(1)={0,1,2}, N(2)={1,2,3}, N(3)={2,3,4}, N(4)={0,3,4}
    nbrs = [[0, 1, 2], [0, 1, 2], [1, 2, 3], [2, 3, 4], [0, 3, 4]]
    nn = sp.csc_matrix((np.ones(15, dtype=bool), np.concatenate(nbrs), np.arange(0, 16, 3)), shape=(5, 5))
    j = {3: 1.0, 2: 0.5, 1: 0.2}                                   # x / (3 + (3 - x))
    shared = np.array([[3, 3, 2, 1, 1], [3, 3, 2, 1, 1], [2, 2, 3, 2, 1], [1, 1, 2, 3, 2], [1, 1, 1, 2, 3]])
    expect = np.vectorize(j.get)(shared)
    S = orc.jaccard_index(nn, 3, prune=1.0 / 15.0)
    assert np.array_equal(S.toarray(), expect) and S.nnz == 25
    S = orc.jaccard_index(nn, 3, prune=0.2)                         # droptol!: abs(x) <= tol is dropped, 0.2 itself goes
    assert np.array_equal(S.toarray(), np.where(expect > 0.2, expect, 0.0)) and S.nnz == 15
    assert np.array_equal(orc.jaccard_index(nn, None, prune=1.0 / 15.0).toarray(), expect)   # diag(snn) = 3 everywhere
    # ragged neighbourhoods: the form without k divides by the size of the COLUMN's neighbourhood (neighbours.jl:103-106)
    nbrs = [[0, 1], [0, 1, 2, 3], [2], [], [1, 2, 4]]
    ptr = np.concatenate([[0], np.cumsum([len(v) for v in nbrs])])
    nn = sp.csc_matrix((np.ones(ptr[-1], dtype=bool), np.concatenate(nbrs).astype(np.int64), ptr), shape=(5, 5))
    S = orc.jaccard_index(nn, None, prune=0.0).toarray()
    for jcol in range(5):
        for i in range(5):
            x = len(set(nbrs[i]) & set(nbrs[jcol]))
            kj = len(nbrs[jcol])
            assert S[i, jcol] == (x / (kj + (kj - x)) if x else 0.0)
    # the device algorithm (emit at the smallest common neighbour, rank sort) gives the same entries, both element types
    rng = np.random.default_rng(3)
    X = rng.random((70, 4))
    for k, T in ((5, np.float64), (5, np.float32), (None, np.float64)):
        nnk = orc.nearest_neighbours(X, 5)
        if k is None:                                              # make it ragged
            nnk = sp.csc_matrix(nnk.multiply(sp.csc_matrix(rng.random((70, 70)) < 0.8)))
            nnk.eliminate_zeros()
            nnk.sort_indices()
        S = orc.jaccard_index(nnk, k, 1.0 / 15.0, T)
        assert S.dtype == T
        cols = _device_snn_algorithm(nnk.indptr, nnk.indices, 70, k, 1.0 / 15.0, T)
        for jcol in range(70):
            a, b = S.indptr[jcol], S.indptr[jcol + 1]
            assert [c[0] for c in cols[jcol]] == list(S.indices[a:b])
            assert [c[1] for c in cols[jcol]] == list(S.data[a:b])
    Sn = orc.shared_nearest_neighbours(X.astype(np.float32), 5)
    assert Sn.dtype == np.float32 and np.array_equal(Sn.toarray(), orc.jaccard_index(orc.nearest_neighbours(X.astype(np.float32), 5), 5, 1 / 15, np.float32).toarray())


def test_variable_feature_selectors_host_arithmetic(orc):
    # variablefeatures.jl:52-103. The product's gene-length host arithmetic (vectorised) against the oracle's literal
    # loop restatement, both fed with the oracle's moments of row_norm(counts, 1). No device involved.
    import severo_jl_b200 as sv
    from conftest import planted_counts
    api = sv.api
    X = planted_counts(400, 900, 5, seed=21, mean_nnz=40.0)
    X = sp.csc_matrix(X)
    X[:, 17] = 0                                                    # an undetected gene: mu = var = 0 -> NaN -> 0
    X.eliminate_zeros()
    norm = orc.row_norm(X, 1.0)
    mu, var = orc.mean_var(norm)
    assert np.array_equal(api._dispersion_metric(mu, var), orc.select_dispersion(norm))
    assert api._dispersion_metric(mu, var)[17] == 0.0
    for nb in (20, 7):
        assert np.array_equal(api._meanvarplot_metric(mu, var, nb), orc.select_meanvarplot(norm, nb))
    brk, lab = api._cut(np.log1p(mu), 20)
    obrk, olab = orc.cut_width(np.log1p(mu), 20)
    assert np.array_equal(brk, obrk) and np.array_equal(lab, olab) and lab.min() >= 1 and lab.max() == 20
    # a value exactly on an inner break belongs to the bin on its left (right = true, utils.jl:120-122)
    v = np.array([0.0, 0.25, 0.5, 1.0])
    assert list(api._cut(v, 4)[1]) == [1, 1, 2, 4] and list(orc.cut_width(v, 4)[1]) == [1, 1, 2, 4]
    trx = np.asarray(X.sum(axis=1)).ravel()
    for alpha in (0.1, 0.99):
        assert np.array_equal(api._saunders_metric(mu, var, trx, X.shape[0], alpha), orc.select_features_saunders(X, norm, alpha))
    assert np.count_nonzero(orc.select_features_saunders(X, norm, 0.1)) > 10
    # :frequency binning puts the smallest mean outside every bin; the reference then indexes out of bounds
    import pytest
    with pytest.raises(IndexError):
        api._meanvarplot_metric(mu, var, 20, "frequency")
    with pytest.raises(ValueError):
        api._cut(mu, 20, "nope")
