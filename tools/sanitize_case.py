"""A small pass through every kernel family of the path, for compute-sanitizer (tools/gpu_session.sh sanitize):
pre-processing sweeps, explicit and count-level operators (replica tables forced on), IRLBA with the device SVD of B,
the mul.jl products, C'C, kNN. Sizes are tiny: memcheck / racecheck slow kernels down 10-100x."""
import os
import sys
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import severo_jl_b200 as sv
from conftest import planted_counts

sv.init(0)
rng = np.random.default_rng(0)
X = planted_counts(1300, 260, 4, seed=1, mean_nnz=60)
Y = sv.normalize_cells(X, method="lognormalize", scale_factor=1e4)
hvf = sv.find_variable_features(X, 120)
S = sv.scale_features(Y, scale_max=10.0, features=hvf)
init = rng.standard_normal(120)
G = sv.irlba(S, 6, init=init, tol=1e-8)
os.environ["SVB_FACT_LOG2R"] = "10"          # the one-CTA-per-SM adjoint kernel with replica tables
C = sv.scale_features_counts(X, scale_factor=1e4, scale_max=10.0, features=hvf, levels=16)
Gc = sv.irlba(C, 6, init=init, tol=1e-8)
assert np.allclose(G.S, Gc.S, rtol=1e-6)
del os.environ["SVB_FACT_LOG2R"]
C2 = sv.scale_features_counts(X, scale_factor=1e4, scale_max=10.0, features=hvf, moments="fast")
v = rng.standard_normal(120)
assert np.allclose(C2 @ v, S @ v, rtol=1e-9, atol=1e-9)
sv.gram(S)
A = sp.random(300, 80, 0.1, random_state=2, format="csc")
x = sp.random(80, 1, 0.3, random_state=3, format="csc")
sv.mul_sparse_vector(np.zeros(300), A, x, 2.0, 0.0)
sv.mul_sparse_dense(np.zeros((300, 5), order="F"), A, sp.random(80, 5, 0.2, random_state=4, format="csc"))
sv.nearest_neighbours(G.U * G.S, 5)
print("sanitize case ok: sigma_1 = %.6f" % G.S[0])
