"""Time of the order-exact Welford kernel (svb_mean_var) on dense genes: 1,306,127 cells x 64 genes, 95 % stored values —
the chain of the densest gene bounds the pre-processing sweeps and the moments computed during the upload.
usage: python tools/welford_timing.py"""
import os, sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import severo_jl_b200 as sv

m, n, dens = 1_306_127, 64, 0.95
rng = np.random.default_rng(1)
cols = []
indptr = [0]
rows_all, vals_all = [], []
for j in range(n):
    keep = np.flatnonzero(rng.random(m) < dens).astype(np.int32)
    rows_all.append(keep)
    vals_all.append(rng.poisson(3.0, keep.size).astype(np.int32) + 1)
    indptr.append(indptr[-1] + keep.size)
A = sp.csc_matrix((np.concatenate(vals_all), np.concatenate(rows_all), np.array(indptr, dtype=np.int64)), shape=(m, n))
sv.init(0)
dA = sv.DeviceMatrix.from_host(A) if hasattr(sv.DeviceMatrix, "from_host") else sv.api._to_device(A)[0]
for kind, M in (("int32 counts", dA), ("float64 log-normalised", sv.normalize_cells(dA, scale_factor=1e4))):
    for i in range(3):
        sv.lib().svb_synchronize()
        t0 = time.perf_counter()
        mu, var = sv.mean_var(M)
        sv.lib().svb_synchronize()
        dt = time.perf_counter() - t0
    print(f"{kind}: {1e3 * dt:.2f} ms = {1e9 * dt / (dens * m):.1f} ns per element of a gene")
