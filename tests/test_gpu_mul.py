"""GPU: the stand-alone sparse products of src/mul.jl — SpMSpV (mul.jl:50-77) and CSC x CSC -> dense
(mul.jl:79-114) — against the oracle's literal restatement of the reference loops. The device kernels add the same terms in
the same order with the same roundings, so every comparison is BIT-exact (assert_array_equal), including alpha / beta forms,
stored zeros in x, empty columns, and the Transpose / Adjoint methods."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _rand_sparse(m, n, density, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    A = sp.random(m, n, density, random_state=seed, format="csc", data_rvs=lambda k: rng.standard_normal(k))
    return A.astype(dtype)


@pytest.mark.parametrize("m,n,da,dx,alpha,beta", [
    (500, 300, 0.05, 0.2, 1.0, 0.0), (500, 300, 0.05, 0.2, 2.0, 1.0), (4000, 1200, 0.02, 0.05, -0.5, 3.0),
    (64, 50, 0.3, 1.0, 1.0, 1.0), (300, 40, 0.1, 0.0, 2.0, 0.5), (300, 40, 0.1, 0.3, 0.0, 2.0)])
def test_spmspv_bit_exact(sv, orc, m, n, da, dx, alpha, beta):
    A = _rand_sparse(m, n, da, 1)
    x = _rand_sparse(n, 1, dx, 2) if dx > 0 else sp.csc_matrix((n, 1))
    y0 = np.random.default_rng(3).standard_normal(m)
    want = orc.spmspv(y0.copy(), A, x, alpha, beta)
    got = sv.mul_sparse_vector(y0.copy(), A, x, alpha, beta)
    np.testing.assert_array_equal(got, want)
    dense = beta * y0 + alpha * (A @ x.toarray().ravel())
    np.testing.assert_allclose(got, dense, rtol=1e-12, atol=1e-12)


def test_spmspv_stored_zero_and_types(sv, orc):
    # a stored zero of x takes part (Inf * 0 = NaN propagates exactly as in the reference); Int32 / Float32 A are promoted
    A = sp.csc_matrix(np.array([[1.0, np.inf], [2.0, 0.0], [0.0, 3.0]]))
    x = sp.csc_matrix((np.array([5.0, 0.0]), np.array([0, 1]), np.array([0, 2])), shape=(2, 1))
    want = orc.spmspv(np.zeros(3), A, x)
    got = sv.mul_sparse_vector(np.zeros(3), A, x)
    np.testing.assert_array_equal(got, want)
    assert np.isnan(got[0]) and got[1] == 10.0
    Ai = sp.random(200, 80, 0.1, random_state=5, format="csc", data_rvs=lambda k: np.random.default_rng(5).integers(1, 9, k)).astype(np.int64)
    xv = _rand_sparse(80, 1, 0.4, 6)
    np.testing.assert_array_equal(sv.mul_sparse_vector(np.zeros(200), Ai, xv, 1.5, 0.0), orc.spmspv(np.zeros(200), Ai, xv, 1.5, 0.0))
    Af = _rand_sparse(200, 80, 0.1, 7, np.float32)
    np.testing.assert_array_equal(sv.mul_sparse_vector(np.zeros(200), Af, xv), orc.spmspv(np.zeros(200), Af.astype(np.float64), xv))
    with pytest.raises(ValueError):
        sv.mul_sparse_vector(np.zeros(200), Af, _rand_sparse(81, 1, 0.4, 6))


@pytest.mark.parametrize("m,k,p,alpha,beta,trans", [
    (300, 200, 17, 1.0, 0.0, False), (300, 200, 17, 2.0, 1.0, False), (1000, 400, 64, -1.5, 0.25, False),
    (300, 200, 9, 1.0, 0.0, True), (50, 30, 1, 3.0, 1.0, True)])
def test_spgemm_dense_bit_exact(sv, orc, m, k, p, alpha, beta, trans):
    A = _rand_sparse(k, m, 0.05, 11) if trans else _rand_sparse(m, k, 0.05, 11)
    B = _rand_sparse(k, p, 0.1, 12)
    C0 = np.asfortranarray(np.random.default_rng(13).standard_normal((m, p)))
    want = orc.spgemm_dense(C0.copy(order="F"), A, B, alpha, beta, transpose_a=trans)
    got = sv.mul_sparse_dense(C0.copy(order="F"), sp.csr_matrix(A.T) if trans else A, B, alpha, beta)
    np.testing.assert_array_equal(got, want)
    Ad = A.T.toarray() if trans else A.toarray()
    np.testing.assert_allclose(got, beta * C0 + alpha * (Ad @ B.toarray()), rtol=1e-11, atol=1e-11)


def test_spgemm_dense_is_the_product_under_CtC(sv, orc):
    # scaling.jl:274-296 builds C'C on top of this product: A'A through mul_sparse_dense equals the oracle's gram without centring
    A = _rand_sparse(2000, 60, 0.1, 21)
    G = sv.mul_sparse_dense(np.zeros((60, 60), order="F"), sp.csr_matrix(A.T), A)
    np.testing.assert_allclose(G, (A.T @ A).toarray(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(G, orc.gram(orc.CenteredMatrix(A, None)), rtol=1e-12, atol=1e-12)
