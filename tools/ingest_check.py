import sys, time, ctypes
sys.path.insert(0, '/root/repo')
import numpy as np
import severo_jl_b200 as sv
import bench
sv.init(0)
cfg = bench.CONFIGS["C3"]
counts = sv.synthetic_counts(cfg["m"], cfg["g"], cfg["nnz"], programs=cfg["programs"], fold=6.0, seed=bench.SEED)
sv.lib().svb_synchronize()
print("counts", counts.shape, counts.nnz)
mu0, var0 = sv.mean_var(counts)
for rep in range(2):
    t0 = time.perf_counter(); T = counts.transpose(); sv.lib().svb_synchronize(); t1 = time.perf_counter()
    print(f"cells x genes -> genes x cells (10X-native): {t1-t0:.3f} s  ({counts.nnz*12*2/1e9/(t1-t0):.0f} GB/s of in+out)")
    t0 = time.perf_counter(); B = T.transpose(); sv.lib().svb_synchronize(); t1 = time.perf_counter()
    print(f"genes x cells -> cells x genes (the copy(X') of input.jl): {t1-t0:.3f} s")
    T.free()
    mu1, var1 = sv.mean_var(B)
    assert B.shape == counts.shape and B.nnz == counts.nnz
    assert np.array_equal(mu0, mu1) and np.array_equal(var0, var1), "round trip changed the order/values"
    s0 = np.zeros(cfg["m"], dtype=np.int64); s1 = np.zeros(cfg["m"], dtype=np.int64)
    sv._lib.check(sv.lib().svb_row_sums(counts._h, sv._lib.ptr(s0))); sv._lib.check(sv.lib().svb_row_sums(B._h, sv._lib.ptr(s1)))
    assert np.array_equal(s0, s1)
    B.free()
print("round trip bit-identical (per-gene Welford moments + per-cell sums)")
