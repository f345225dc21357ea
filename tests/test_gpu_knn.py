"""GPU parity tests of SURVEY §8f #4 (second half): exact k-nearest neighbours on the device (``svb_knn`` /
``svb_knn_result``) against the oracle's brute-force search, with the reference's own acceptance criterion
(test/test_nn.jl: Jaccard overlap with ``partialsortperm`` of the pairwise distances, 30 % quantile == 1.0)."""
import ctypes

import numpy as np
import pytest

from conftest import planted_counts

pytestmark = pytest.mark.gpu


def _same_neighbour_sets(idx, ref):
    return np.array_equal(np.sort(idx, axis=1), np.sort(ref, axis=1))


@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
def test_ann_reference_cases(sv, orc, metric):
    # test/test_nn.jl:22-56 (100 x 10, k = 4), :58-77 (a column view), :79-96 (Float32 in -> Float32 distances)
    rng = np.random.default_rng(7)
    X = rng.random((100, 20))
    for Z in (X[:, :10], np.asfortranarray(X)[:, :10], X):
        idx, dist = sv.ann(Z, 4, metric=metric)
        assert idx.dtype == np.int32 and dist.dtype == np.float64 and idx.shape == (100, 4)
        ref_idx, ref_dist = orc.knn(Z, 4, metric, include_self=True)
        j = [len(set(a) & set(b)) / (4 + (4 - len(set(a) & set(b)))) for a, b in zip(idx, ref_idx)]
        assert np.quantile(j, 0.3) == 1.0           # the reference's criterion
        assert np.array_equal(idx, ref_idx)         # and the exact answer, order included
        np.testing.assert_allclose(dist, ref_dist, rtol=1e-12, atol=1e-14)
    X32 = rng.random((100, 20)).astype(np.float32)
    idx, dist = sv.ann(X32, 4, metric=metric)
    assert dist.dtype == np.float32
    ref_idx, ref_dist = orc.knn(X32.astype(np.float64), 4, metric, include_self=True)
    assert _same_neighbour_sets(idx, ref_idx)
    np.testing.assert_allclose(dist, ref_dist, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("case", [(3000, 7, 5), (3000, 50, 20), (2049, 10, 20), (700, 64, 64), (1500, 33, 1), (300, 80, 4), (400, 100, 10), (130, 128, 3), (500, 11, 5), (640, 19, 6)])
def test_knn_exact_against_oracle(sv, orc, case):
    # every kind of padded width of the kernel family (8, 50, 10, 64, 40, 96, 128-with-spills, 12, 20: knn_padded_dims), n not a multiple of
    # the tile or CTA size
    n, d, k = case
    rng = np.random.default_rng(n + d)
    centres = rng.standard_normal((8, d)) * 3.0
    X = centres[rng.integers(0, 8, n)] + rng.standard_normal((n, d))
    for metric in ("euclidean", "cosine"):
        for include_self in (True, False):
            idx, dist = sv.ann(X, k, metric=metric, include_self=include_self)
            ref_idx, ref_dist = orc.knn(X, k, metric, include_self)
            assert _same_neighbour_sets(idx, ref_idx), (metric, include_self)
            assert np.all(np.diff(dist, axis=1) >= -1e-12)
            np.testing.assert_allclose(np.sort(dist, axis=1), ref_dist, rtol=1e-11, atol=1e-13)
            if include_self:
                assert np.array_equal(idx[:, 0], np.arange(n)) and np.all(dist[:, 0] == 0.0)
            else:
                assert not np.any(idx == np.arange(n)[:, None])


def test_knn_ties_strides_and_errors(sv, orc):
    L = sv._lib
    rng = np.random.default_rng(2)
    # duplicated cells: ties resolved by the lower index, like a stable partialsortperm
    base = rng.random((50, 6))
    X = np.vstack([base, base, base])
    idx, dist = sv.ann(X, 5, include_self=False)
    ref_idx, ref_dist = orc.knn(X, 5, "euclidean", include_self=False)
    assert np.array_equal(idx, ref_idx)
    np.testing.assert_allclose(dist, ref_dist, rtol=1e-12, atol=1e-14)
    # raw ABI with a column stride larger than n (stride(X,2) of a row-sliced parent, neighbours.jl:35-41) and 1-based indices
    parent = np.asfortranarray(rng.random((260, 9)))
    n, d, k = 200, 9, 6
    nn = np.zeros((n, k), dtype=np.int32, order="F")
    dd = np.zeros((n, k), order="F")
    L.check(sv.lib().svb_knn(L.ptr(parent), L.SVB_F64, n, d, 260, k, L.METRIC_EUCLIDEAN, 1, 1, L.ptr(nn), L.ptr(dd)))
    ref_idx, ref_dist = orc.knn(parent[:n], k, "euclidean", True)
    assert np.array_equal(nn, ref_idx + 1)
    np.testing.assert_allclose(dd, ref_dist, rtol=1e-12, atol=1e-14)
    # a zero vector under the cosine metric: distance 1 to everything (documented deviation: Distances.jl gives NaN)
    Xz = rng.random((40, 5))
    Xz[3] = 0.0
    idx, dist = sv.ann(Xz, 3, metric="cosine", include_self=False)
    assert np.all(dist[3] == 1.0) and np.all(np.isfinite(dist))
    for bad in (dict(k=65), dict(k=0), dict(k=41)):
        with pytest.raises(sv.SeveroB200Error):
            sv.ann(Xz, bad["k"])
    with pytest.raises(sv.SeveroB200Error):
        sv.ann(rng.random((10, 129)), 2)
    with pytest.raises(KeyError):
        sv.ann(Xz, 3, metric="manhattan")


def test_nearest_neighbours_of_an_embedding(sv, orc):
    # docs/src/pbmc.md:157: nn = nearest_neighbours(em, 20, dims=1:10) right after embedding(...)
    X = planted_counts(3000, 600, 8, seed=5, mean_nnz=80)
    Y = sv.normalize_cells(X, scale_factor=1e4)
    hvf = sv.find_variable_features(X, 300)
    S = sv.scale_features(Y, scale_max=10.0, features=hvf)
    init = np.random.default_rng(1).standard_normal(300)
    em = sv.embedding(S, 12, method="pca", algorithm="irlba", init=init, tol=1e-10)
    nn = sv.nearest_neighbours(em, 20, dims=slice(0, 10))
    Z = em.coordinates.array if hasattr(em.coordinates, "array") else em.coordinates
    ref = orc.nearest_neighbours(Z, 20, dims=slice(0, 10))
    A = nn.array if hasattr(nn, "array") else nn
    assert A.shape == (3000, 3000) and A.nnz == 60000
    assert (A != ref).nnz == 0
    # the same graph straight from the solve that is still in HBM (svb_knn_result): no host round trip of the coordinates
    L = sv._lib
    r = ctypes.c_void_p()
    L.check(sv.lib().svb_irlba_solve(S._operator(), 12, 0, 1000, 0, 1e-10, 1e-10, L.ptr(init), None, None, None, ctypes.byref(r)))
    idx = np.zeros((3000, 20), dtype=np.int32, order="F")
    dist = np.zeros((3000, 20), order="F")
    L.check(sv.lib().svb_knn_result(r, 10, 20, L.METRIC_EUCLIDEAN, 1, 0, L.ptr(idx), L.ptr(dist)))
    sv.lib().svb_result_free(r)
    ref_idx, ref_dist = orc.knn(Z[:, :10], 20, "euclidean", True)
    # the two solves are the same computation: coordinates agree to rounding, so do the neighbour sets up to exact-tie flips
    same = np.mean([set(a) == set(b) for a, b in zip(idx, ref_idx)])
    assert same >= 0.999
    np.testing.assert_allclose(np.sort(dist, axis=1), ref_dist, rtol=1e-8, atol=1e-10)


_Q2 = r"""
import sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import severo_jl_b200 as sv
from oracle import severo_oracle as orc
sv.init(0)
for (n, d, k) in ((3000, 7, 5), (2049, 10, 20), (1500, 30, 4), (257, 32, 64)):
    rng = np.random.default_rng(n + d)
    X = rng.standard_normal((8, d))[rng.integers(0, 8, n)] * 3.0 + rng.standard_normal((n, d))
    for metric in ("euclidean", "cosine"):
        for include_self in (True, False):
            idx, dist = sv.ann(X, k, metric=metric, include_self=include_self)
            ref_idx, ref_dist = orc.knn(X, k, metric, include_self)
            assert np.array_equal(np.sort(idx, axis=1), np.sort(ref_idx, axis=1)), (n, d, k, metric, include_self)
            np.testing.assert_allclose(np.sort(dist, axis=1), ref_dist, rtol=1e-11, atol=1e-13)
print("OK")
"""


def test_knn_two_queries_per_thread_variant(tmp_path):
    # the D <= 32 kernels with two query cells per thread (SVB_KNN_Q=2 forces them at any n)
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "knn_q2.py"
    script.write_text(_Q2)
    r = subprocess.run([sys.executable, str(script), root], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, SVB_KNN_Q="2"))
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
