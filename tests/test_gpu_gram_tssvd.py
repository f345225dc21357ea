"""GPU parity tests of SURVEY §8f #1: the Gram matrix ``C'C`` (scaling.jl:274-296 over mul.jl:82-114, ``svb_gram``) and
``tssvd`` (embedding.jl:30-44, ``svb_tssvd``) against the oracle / the dense matrix. The reference has no test for
either, so — as test_irlba.jl does for the solver — results are pinned on ``convert(Matrix, S)`` and the dense SVD."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import planted_counts

pytestmark = pytest.mark.gpu


def _check_gram(G, D, rtol):
    ref = D.T @ D
    # the fused kernel sums every entry in tile order, the centring terms are rounded separately: G is exactly symmetric
    assert np.array_equal(G, G.T)
    assert np.abs(G - ref).max() <= rtol * np.abs(ref).max()


@pytest.mark.parametrize("shape", [(2000, 400, 0.1), (9000, 333, 0.03), (4000, 50, 0.5), (300, 3, 0.4), (13000, 7, 0.9)])
def test_gram_sparse_matches_oracle_and_dense(sv, orc, shape):
    # 333 and 7 genes: a last block of 1 / 3 genes; 9000 and 13000 cells: 3 / 4 cell tiles of 4096; density 0.5 / 0.9 with
    # one tile: (tile, gene) segments of >= 64 entries (the 32-lane variant of the kernel)
    m, n, dens = shape
    rng = np.random.default_rng(m + n)
    X = sp.random(m, n, dens, random_state=m, format="csc")
    mu = rng.standard_normal(n)
    for mu_ in (mu, None):
        C, O = sv.CenteredMatrix(X, mu_), orc.CenteredMatrix(X, mu_)
        G = sv.gram(C)
        _check_gram(G, O.to_dense(), 1e-13)
        assert np.abs(G - orc.gram(O)).max() <= 1e-13 * np.abs(G).max()
        C.free()


def test_gram_other_operator_kinds(sv, orc):
    rng = np.random.default_rng(8)
    # lazy adjoint parent (test_irlba.jl:111): genes x cells CSC
    Xc = sp.random(60, 5000, 0.05, random_state=4, format="csc", data_rvs=lambda k: rng.poisson(10, k).astype(float))
    mu = rng.standard_normal(60)
    C = sv.CenteredMatrix(Xc.T, mu)
    assert C.shape == (5000, 60)
    _check_gram(sv.gram(C), orc.CenteredMatrix(Xc, mu, transposed=True).to_dense(), 1e-13)
    # Float32 storage of the values, Float64 accumulation
    X = sp.random(6000, 120, 0.1, random_state=2, format="csc")
    mu = np.asarray(X.mean(axis=0)).ravel()
    C32 = sv.CenteredMatrix(X, mu, storage="f32")
    _check_gram(sv.gram(C32), orc.CenteredMatrix(X, mu).to_dense(), 1e-6)
    # dense operator: column by column through the vector products, then symmetrised
    Xd = rng.standard_normal((500, 37))
    Cd = sv.CenteredMatrix(Xd, Xd.mean(axis=0))
    _check_gram(sv.gram(Cd), Cd.to_dense(), 1e-13)
    # count-level operator (scaled matrix never materialised): against the oracle's explicit scaled matrix
    cnt = planted_counts(3000, 260, 5, seed=3, mean_nnz=40)
    keep = np.nonzero(np.asarray((cnt > 0).sum(axis=0)).ravel() > 1)[0]
    Cc = sv.scale_features_counts(cnt, scale_factor=1e4, scale_max=10.0, features=keep)
    So = orc.scale_features(orc.normalize_cells(cnt, "lognormalize", 1e4), scale_max=10.0, features=keep)
    _check_gram(sv.gram(Cc), So.to_dense(), 1e-12)


def test_tssvd_against_dense_svd(sv, orc):
    X = planted_counts(4000, 900, 9, seed=31, mean_nnz=100)
    Y = sv.normalize_cells(X, scale_factor=1e4)
    hvf = sv.find_variable_features(X, 400)
    S = sv.scale_features(Y, scale_max=10.0, features=hvf)
    init = np.random.default_rng(1).standard_normal(400)
    T = sv.tssvd(S, nsv=9, init=init)                       # upstream defaults: tol = 0.0, ncv = 2*nsv
    D = S.to_dense()
    U, s, Vt = np.linalg.svd(D, full_matrices=False)
    # sigma = sqrt(lambda): relative error ~ eps*(sigma_1/sigma_i)^2; 1e-10 stated, ~1e-14 expected
    np.testing.assert_allclose(T.S, s[:9], rtol=1e-10)
    assert T.U.shape == (4000, 9) and T.Vt.shape == (9, 400)
    assert orc.principal_angle(T.V, Vt[:9].T) < 1e-6 and orc.principal_angle(T.U, U[:, :9]) < 1e-6
    assert np.abs(T.U.T @ T.U - np.eye(9)).max() < 1e-8       # U = A*phi/Sigma is orthonormal only through phi's accuracy
    assert np.linalg.norm(D.T @ T.U - T.V * T.S[None, :]) < 1e-8 * s[0]   # the reference's own criterion, test_irlba.jl:30
    # the oracle's restatement (LAPACK eigh of the same Gram matrix)
    O = orc.tssvd(orc.scale_features(Y, scale_max=10.0, features=hvf), 9)
    np.testing.assert_allclose(T.S, O.S, rtol=1e-10)
    # ritzvec = false: an m x 0 U (embedding.jl:40-42)
    T0 = sv.tssvd(S, nsv=5, ritzvec=False, init=init)
    assert T0.U.shape == (4000, 0)
    np.testing.assert_allclose(T0.S, s[:5], rtol=1e-10)
    # embedding(...; algorithm=:tssvd) — embedding.jl:58-59 — against the IRLBA path
    em = sv.embedding(S, 9, method="pca", algorithm="tssvd", init=init)
    em2 = sv.embedding(S, 9, method="pca", algorithm="irlba", init=init, tol=1e-10)
    np.testing.assert_allclose(em.stdev, em2.stdev, rtol=1e-9)
    assert em.coordinates.shape == (4000, 9) and em.basis.shape == (400, 9)
    # count-level operator through the same call
    Cc = sv.scale_features_counts(X, scale_factor=1e4, scale_max=10.0, features=hvf)
    Tc = sv.tssvd(Cc, nsv=9, init=init)
    np.testing.assert_allclose(Tc.S, s[:9], rtol=1e-9)


def test_tssvd_small_gaps_and_errors(sv):
    # unstructured sparse matrix: clustered singular values, many restarts of the n x n eigen-solve
    X = sp.random(3000, 300, 0.05, random_state=1, format="csc")
    mu = np.asarray(X.mean(axis=0)).ravel()
    C = sv.CenteredMatrix(X, mu)
    s = np.linalg.svd(X.toarray() - mu[None, :], compute_uv=False)
    T = sv.tssvd(C, nsv=10, init=np.random.default_rng(0).standard_normal(300))
    np.testing.assert_allclose(T.S, s[:10], rtol=1e-10)
    with pytest.raises(ValueError):
        sv.tssvd(C, nsv=10, init=np.zeros(7))
    with pytest.raises(sv.SeveroB200Error):
        sv.tssvd(C, nsv=301, init=np.ones(300))
    with pytest.raises(RuntimeError):
        sv.tssvd(C, nsv=10, maxiter=1, tol=1e-14, init=np.random.default_rng(0).standard_normal(300))


_TWO_RANK_GRAM = r"""
import os, sys
import numpy as np, scipy.sparse as sp
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import severo_jl_b200 as sv
from severo_jl_b200 import sharding
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
sv.init(local)
sharding.init_comm_from_torch()
rng = np.random.default_rng(0)
m, n = 30011, 300
X = sp.random(m, n, 0.05, random_state=3, format="csc")
mu = np.asarray(X.mean(axis=0)).ravel()
bounds = sharding.shard_bounds(m, world)
lo, hi = bounds[rank]
C = sv.CenteredMatrix(sp.csc_matrix(X[lo:hi]), mu)
D = X.toarray() - mu[None, :]
G = sv.gram(C)
ref = D.T @ D
assert np.array_equal(G, G.T) and np.abs(G - ref).max() <= 1e-13 * np.abs(ref).max()
T = sv.tssvd(C, nsv=8, init=rng.standard_normal(n))
s = np.linalg.svd(D, compute_uv=False)
assert np.allclose(T.S, s[:8], rtol=1e-10), (T.S, s[:8])
U = sharding.gather_rows(T.U, bounds, rank)
assert np.linalg.norm(D.T @ U - T.V * T.S[None, :]) < 1e-8 * s[0]
sv.lib().svb_comm_destroy()
dist.destroy_process_group()
print("OK", rank)
"""


def test_two_gpu_gram_and_tssvd(tmp_path):
    import os
    import subprocess
    import sys
    try:
        ngpu = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout.count("GPU ")
    except Exception:
        ngpu = 0
    if ngpu < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "two_rank_gram.py"
    script.write_text(_TWO_RANK_GRAM)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29646", str(script), root], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == 2
