"""GPU: the single-process multi-GPU boundary (csrc/multi.cu): ONE call from ONE host thread takes the whole host
SparseMatrixCSC, the library shards it by cells over its worker threads (one per GPU, ncclCommInitAll, peer-access mailboxes)
and returns U, s, V — the call shape of src/irlba.jl:66-71. Checked against the oracle on the whole matrix (sigma rel 1e-6,
principal angle < 1e-4). Uses every GPU of the box (2+ on the multi-GPU boxes; the executor path itself also runs on one)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import planted_counts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def group(sv):
    import torch
    n = min(torch.cuda.device_count(), 8)
    sv.init_devices(n)
    info = sv.devices_info()
    assert info["ndev"] == n and info["devices"] == list(range(n))
    assert info["peer_mailboxes"] == (n > 1)          # NVLink peer access between the GPUs of one box
    yield info
    sv.shutdown_devices()


def test_irlba_whole_matrix_one_call(sv, orc, group):
    rng = np.random.default_rng(0)
    m, n, nu = 30011, 500, 8
    X = sp.random(m, n, 0.05, random_state=3, format="csc")
    u, v = rng.standard_normal((m, nu)), rng.standard_normal((n, nu))
    X = sp.csc_matrix(X + sp.csc_matrix((u * np.linspace(3, 1, nu)) @ v.T * (rng.random((m, n)) < 0.02)))
    mu = np.asarray(X.mean(axis=0)).ravel()
    init = rng.standard_normal(n)
    S = sv.irlba_devices(X, nu, mu=mu, init=init, tol=1e-9)
    O = orc.irlba(orc.CenteredMatrix(X, mu), nu, init=init, tol=1e-9)
    np.testing.assert_allclose(S.S, O.S, rtol=1e-6)
    assert S.U.shape == (m, nu) and orc.principal_angle(S.U, O.U) < 1e-4 and orc.principal_angle(S.V, O.V) < 1e-4
    # the reference's acceptance criterion with the oracle's operator (test_irlba.jl:30)
    Co = orc.CenteredMatrix(X, mu)
    StU = Co.mul(np.asfortranarray(S.U), trans=True)
    assert np.linalg.norm(StU - S.V * S.S) / np.linalg.norm(Co.to_dense()) < 1e-9
    # Julia-style arrays through the raw entry point: 1-based Int64 indices
    L = sv._lib
    colptr = X.indptr.astype(np.int64) + 1
    rowval = X.indices.astype(np.int64) + 1
    U = np.zeros((m, nu), order="F"); s = np.zeros(nu); V = np.zeros((n, nu), order="F")
    import ctypes
    it, mp = ctypes.c_int64(), ctypes.c_int64()
    L.check(L.lib().svb_irlba_csc_devices(m, n, L.ptr(colptr), L.ptr(rowval), L.SVB_I64, L.ptr(np.ascontiguousarray(X.data)), L.SVB_F64,
                                          1, L.ptr(mu), nu, nu + 7, 1000, 1e-9, 1e-9, L.ptr(init), L.ptr(s), L.ptr(U), L.ptr(V),
                                          ctypes.byref(it), ctypes.byref(mp)))
    np.testing.assert_array_equal(s, S.S)              # same shards, same arithmetic: identical bits
    np.testing.assert_array_equal(U, S.U)


def test_pca_counts_whole_matrix_one_call(sv, orc, group):
    X = planted_counts(12000, 700, 6, seed=8, mean_nnz=120)
    hvf = sv.find_variable_features(X, 300)
    libsize = np.asarray(X.sum(axis=1)).ravel().astype(np.int64)
    chv = sp.csc_matrix(X[:, hvf])
    init = np.random.default_rng(2).standard_normal(300)
    S, mu = sv.pca_counts_devices(chv, libsize, 6, scale_factor=1e4, scale_max=10.0, init=init, tol=1e-9)
    So = orc.scale_features(orc.normalize_cells(X, "lognormalize", 1e4), scale_max=10.0, features=hvf)
    np.testing.assert_allclose(mu, So.mu, rtol=1e-11)
    O = orc.irlba(So, 6, init=init, tol=1e-9)
    np.testing.assert_allclose(S.S, O.S, rtol=1e-6)
    assert orc.principal_angle(S.U, O.U) < 1e-4 and orc.principal_angle(S.V, O.V) < 1e-4


def test_devices_errors_and_single_gpu_context_coexist(sv, orc, group):
    # a bad argument comes back as a code + message, the workers stay usable; the one-GPU context of svb_init keeps working
    with pytest.raises(sv.SeveroB200Error):
        sv.irlba_devices(sp.random(50, 20, 0.2, format="csc"), 40)            # nu > min(m, n)
    X = sp.random(4000, 120, 0.1, random_state=1, format="csc")
    a = sv.irlba_devices(X, 5, init=np.ones(120), tol=1e-9)
    b = sv.irlba(X, 5, init=np.ones(120), tol=1e-9)                           # svb_init context, device 0
    np.testing.assert_allclose(a.S, b.S, rtol=1e-9)
    np.testing.assert_allclose(a.S, np.linalg.svd(X.toarray(), compute_uv=False)[:5], rtol=1e-8)


def test_devices_refuse_malformed_csc_and_stay_usable(sv, group):
    """The library cuts the caller's columns by binary search, which is only right for ascending rows: unsorted or out-of-range
    row indices must come back as SVB_EDIM from the phase that has no collective in it (no worker left waiting), and the
    group must solve the next, well-formed matrix."""
    import ctypes
    L = sv._lib
    m, n, nu = 6000, 60, 4
    X = sp.random(m, n, 0.05, random_state=2, format="csc")
    X.sort_indices()
    colptr = X.indptr.astype(np.int64)
    data = np.ascontiguousarray(X.data)
    init = np.ones(n)

    def call(rowval):
        U = np.zeros((m, nu), order="F"); s = np.zeros(nu); V = np.zeros((n, nu), order="F")
        it, mp = ctypes.c_int64(), ctypes.c_int64()
        return L.lib().svb_irlba_csc_devices(m, n, L.ptr(colptr), L.ptr(rowval), L.SVB_I64, L.ptr(data), L.SVB_F64, 0, None, nu, nu + 7,
                                             1000, 1e-9, 1e-9, L.ptr(init), L.ptr(s), L.ptr(U), L.ptr(V), ctypes.byref(it), ctypes.byref(mp)), s

    good = X.indices.astype(np.int64)
    j = int(np.argmax(np.diff(colptr)))                       # the longest column
    b, e = int(colptr[j]), int(colptr[j + 1])
    swapped = good.copy()
    swapped[b], swapped[e - 1] = good[e - 1], good[b]         # first and last row of the column exchanged: not ascending
    rc, _ = call(swapped)
    assert rc == L.SVB_EDIM
    beyond = good.copy()
    beyond[e - 1] = m + 5                                     # a row index outside [0, m)
    rc, _ = call(beyond)
    assert rc == L.SVB_EDIM
    rc, s = call(good)
    assert rc == 0
    np.testing.assert_allclose(s, np.linalg.svd(X.toarray(), compute_uv=False)[:nu], rtol=1e-8)
