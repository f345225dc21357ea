"""GPU parity tests of the consumer of the kNN graph: ``jaccard_index`` / ``shared_nearest_neighbours`` on the device
(``svb_jaccard_index``, csrc/snn.cu) against the oracle's restatement of neighbours.jl:88-110 (``nn' * nn`` with the generic
sparse product, ``x / (k + (k - x))``, ``droptol!``). Integer work plus one rounded division per entry: bit-exact."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _same_csc(a, b):
    a, b = sp.csc_matrix(a), sp.csc_matrix(b)
    return (a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.indptr, b.indptr)
            and np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data))


def _clustered(n, d, seed):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((6, d)) * 3.0
    return centres[rng.integers(0, 6, n)] + rng.standard_normal((n, d))


@pytest.fixture(params=["hash", "general"])
def snn_path(request, monkeypatch):
    """Both device paths of csrc/snn.cu: per-warp hash tables (default; hub columns fall back inside the kernel) and the
    general smallest-common-neighbour enumeration for every column (SVB_SNN_HASH=0, read per call)."""
    monkeypatch.setenv("SVB_SNN_HASH", "1" if request.param == "hash" else "0")
    return request.param


def test_jaccard_hand_computed_case(sv, snn_path):
    # same case as tests/test_oracle_golden.py::test_jaccard_index_restatement
    nbrs = [[0, 1, 2], [0, 1, 2], [1, 2, 3], [2, 3, 4], [0, 3, 4]]
    nn = sp.csc_matrix((np.ones(15, dtype=bool), np.concatenate(nbrs), np.arange(0, 16, 3)), shape=(5, 5))
    shared = np.array([[3, 3, 2, 1, 1], [3, 3, 2, 1, 1], [2, 2, 3, 2, 1], [1, 1, 2, 3, 2], [1, 1, 1, 2, 3]])
    expect = np.vectorize({3: 1.0, 2: 0.5, 1: 0.2}.get)(shared)
    S = sv.jaccard_index(nn, 3)
    assert S.has_sorted_indices and np.array_equal(S.toarray(), expect) and S.nnz == 25
    S = sv.jaccard_index(nn, 3, prune=0.2)
    assert np.array_equal(S.toarray(), np.where(expect > 0.2, expect, 0.0)) and S.nnz == 15
    assert np.array_equal(sv.jaccard_index(nn, prune=1.0 / 15.0).toarray(), expect)


@pytest.mark.parametrize("case", [(300, 5, 4), (2000, 10, 20), (1500, 8, 33), (4097, 6, 15)])
def test_jaccard_index_against_oracle(sv, orc, case, snn_path):
    n, d, k = case
    X = _clustered(n, d, n + k)
    nn = orc.nearest_neighbours(X, k)
    for kk, prune, T in ((k, 1.0 / 15.0, np.float64), (k, 1.0 / 15.0, np.float32), (k, 0.0, np.float64), (None, 1.0 / 15.0, np.float64),
                         (k, 0.35, np.float32)):
        S = sv.jaccard_index(nn, kk, prune=prune, dtype=T)
        assert _same_csc(S, orc.jaccard_index(nn, kk, prune, T)), (kk, prune, T)
    # symmetric (every cell has k neighbours) with a unit diagonal
    S = sv.jaccard_index(nn, k)
    assert (abs(S - S.T)).nnz == 0 and np.all(S.diagonal() == 1.0)


def test_jaccard_ragged_hubs_and_empty_columns(sv, orc, snn_path):
    rng = np.random.default_rng(11)
    n = 900
    # ragged neighbourhoods, some empty, and a hub that a third of the cells point at (long reverse list: several
    # 32-lane rounds per neighbour, many survivors per column)
    M = sp.random(n, n, density=0.01, random_state=5, format="csc", dtype=np.float64)
    M.data[:] = 1.0
    M = sp.lil_matrix(M)
    M[7, rng.choice(n, n // 3, replace=False)] = 1.0
    M[:, 100:110] = 0.0
    nn = sp.csc_matrix(M).astype(bool)
    nn.eliminate_zeros()
    nn.sort_indices()
    assert np.diff(nn.indptr).min() == 0 and np.diff(sp.csr_matrix(nn).indptr).max() >= n // 3 - 10
    for kk, prune, T in ((None, 1.0 / 15.0, np.float64), (None, 0.0, np.float32), (12, 1.0 / 15.0, np.float64), (2, 0.1, np.float64)):
        # k = 2 is smaller than many neighbourhoods: zero / negative denominators follow IEEE like the reference (inf, < 0 kept)
        S = sv.jaccard_index(nn, kk, prune=prune, dtype=T)
        ref = orc.jaccard_index(nn, kk, prune, T)
        assert np.array_equal(S.indptr, ref.indptr) and np.array_equal(S.indices, ref.indices)
        assert np.array_equal(S.data, ref.data, equal_nan=True), (kk, prune, T)
    # stored `false` entries are not neighbours
    withfalse = nn.copy().astype(bool)
    withfalse.data[::3] = False
    ref = nn.copy()
    ref.data = withfalse.data.copy()
    ref.eliminate_zeros()
    assert _same_csc(sv.jaccard_index(withfalse, 10), orc.jaccard_index(ref, 10))
    # empty graph
    E = sv.jaccard_index(sp.csc_matrix((50, 50), dtype=bool), 5)
    assert E.shape == (50, 50) and E.nnz == 0


def test_jaccard_hub_columns_leave_the_hash_path(sv, orc):
    # a cell that 2,500 others point at: every column holding it has more candidate occurrences than a warp's hash table
    # takes (1,536) and is enumerated by the general path inside the hash kernel; the other columns stay on the hash path
    rng = np.random.default_rng(5)
    n, k = 5000, 8
    idx = np.array([rng.choice(n, k, replace=False) for _ in range(n)])
    hub_cols = rng.choice(n, 2500, replace=False)
    idx[hub_cols, 0] = 42
    idx = np.array([np.unique(r) for r in idx], dtype=object)
    ptr = np.concatenate([[0], np.cumsum([len(r) for r in idx])])
    nn = sp.csc_matrix((np.ones(ptr[-1], dtype=bool), np.concatenate(idx).astype(np.int64), ptr), shape=(n, n))
    assert np.bincount(nn.indices, minlength=n)[42] >= 2500
    for kk, prune in ((k, 1.0 / 15.0), (None, 0.0)):
        assert _same_csc(sv.jaccard_index(nn, kk, prune=prune), orc.jaccard_index(nn, kk, prune))


def test_shared_nearest_neighbours_and_errors(sv, orc):
    L = sv._lib
    X = _clustered(1200, 12, 3)
    for Z in (X, X.astype(np.float32)):
        S = sv.shared_nearest_neighbours(Z, 10, dims=slice(0, 8))
        assert S.dtype == Z.dtype
        nn = sv.nearest_neighbours(Z, 10, dims=slice(0, 8))
        assert _same_csc(S, orc.jaccard_index(nn, 10, 1.0 / 15.0, Z.dtype))      # the Jaccard step alone is bit-exact
    # labelled input -> labelled output (neighbours.jl:130,269)
    names = ["c%d" % i for i in range(1200)]
    em = sv.NamedArray(X, (names, ["PC-%d" % i for i in range(12)]), ("cells", "latent"))
    Sn = sv.shared_nearest_neighbours(em, 10)
    assert isinstance(Sn, sv.NamedArray) and Sn.names == (names, names) and Sn.dimnames == ("cells", "cells")
    assert _same_csc(Sn.array, sv.shared_nearest_neighbours(X, 10))
    # raw ABI: 1-based Julia arrays in, 1-based out; unsorted rows and non-square graphs are refused
    nn = orc.nearest_neighbours(X[:200], 6)
    dN = sv.DeviceMatrix.from_julia_arrays(200, 200, nn.indptr.astype(np.int64) + 1, nn.indices.astype(np.int64) + 1, np.ones(nn.nnz, dtype=np.int32))
    h = ctypes.c_void_p()
    L.check(sv.lib().svb_jaccard_index(dN._h, 6, 1.0 / 15.0, L.SVB_F64, ctypes.byref(h)))
    out = sv.DeviceMatrix(h)
    ref = orc.jaccard_index(nn, 6)
    colptr, rowval, nz = np.zeros(201, dtype=np.int64), np.zeros(out.nnz, dtype=np.int64), np.zeros(out.nnz)
    L.check(sv.lib().svb_matrix_download(out._h, L.ptr(colptr), L.ptr(rowval), L.ptr(nz), L.SVB_F64, 1))
    assert np.array_equal(colptr, ref.indptr + 1) and np.array_equal(rowval, ref.indices + 1) and np.array_equal(nz, ref.data)
    out.free()
    assert sv.lib().svb_jaccard_index(dN._h, 6, 1.0 / 15.0, L.SVB_I32, ctypes.byref(h)) == L.SVB_EARG
    dN.free()
    with pytest.raises(sv.SeveroB200Error):   # unsorted rows are refused at the upload already (svb_csc_upload validates: round 2)
        sv.DeviceMatrix.from_julia_arrays(3, 3, np.array([1, 3, 4, 5]), np.array([2, 1, 3, 1]), np.ones(4, dtype=np.int32))
    rect = sv.DeviceMatrix.from_julia_arrays(4, 3, np.array([1, 2, 3, 4]), np.array([1, 2, 4]), np.ones(3, dtype=np.int32))
    assert sv.lib().svb_jaccard_index(rect._h, 2, 0.0, L.SVB_F64, ctypes.byref(h)) == L.SVB_EDIM
    rect.free()
    with pytest.raises(ValueError):
        sv.jaccard_index(sp.csc_matrix((4, 4), dtype=bool), 0)
