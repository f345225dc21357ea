"""tools/gram_check.py — full-size timing + parity of C'C (svb_gram) and tssvd (svb_tssvd) on the explicit operator (1 GPU).
usage: python tools/gram_check.py [C3|C2|C1] [cells]
Prints one JSON line: Gram / tssvd wall times (device synchronised on both sides), the algorithmic bytes of the Gram pass
(10 B per nonzero per block of 4 genes, lower triangle only => z*10*(n/4)/2 + 8*n*n), symmetry, and the
singular values of tssvd against the device IRLBA on the same operator."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import severo_jl_b200 as sv  # noqa: E402
from bench import CONFIGS, SCALE_MAX, SEED, solve_device  # noqa: E402
from severo_jl_b200 import sharding  # noqa: E402

L = sv._lib


def main():
    cfg = dict(CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"])
    if len(sys.argv) > 2:
        cfg["m"] = int(sys.argv[2])
    lib = sv.init(0)
    out = {"config": cfg}
    m, n, nu = cfg["m"], cfg["n"], cfg["nu"]
    counts = sv.synthetic_counts(m, cfg["g"], cfg["nnz"], programs=cfg["programs"], fold=6.0, seed=SEED)
    Y = sv.normalize_cells(counts, method="lognormalize", scale_factor=1e4)
    metric = sharding.sharded_vst_metric(counts)
    hvf = np.argsort(-metric, kind="stable")[:n]
    counts.free()
    B, mu = sharding.sharded_scale_features(Y, scale_max=SCALE_MAX, features=hvf)
    Y.free()
    z = int(B.nnz)
    out["hvg_nnz"] = z
    h = ctypes.c_void_p()
    L.check(lib.svb_operator_create_ex(B._h, L.ptr(np.ascontiguousarray(mu)), 0, 0, ctypes.byref(h)))
    B.free()
    G = np.zeros((n, n), order="F")
    L.check(lib.svb_gram(h, L.ptr(G)))  # warm-up (allocations, function attributes)
    lib.svb_synchronize()
    t0 = time.perf_counter()
    L.check(lib.svb_gram(h, L.ptr(G)))
    lib.svb_synchronize()
    dt = time.perf_counter() - t0
    alg = z * 10.0 * (n / 4.0) / 2.0 + 8.0 * n * n
    out["gram_s"] = round(dt, 5)
    out["gram_algorithmic_GB"] = round(alg / 1e9, 2)
    out["gram_GBps"] = round(alg / dt / 1e9, 1)
    out["gram_symmetric"] = bool(np.array_equal(G, G.T))
    # spot parity: G v against S'(S v) through the vector products
    v = np.random.default_rng(1).standard_normal(n)
    y = np.zeros(m)
    z2 = np.zeros(n)
    L.check(lib.svb_mul(h, b"N", 1.0, L.ptr(v), 0.0, L.ptr(y), 1))
    L.check(lib.svb_mul(h, b"T", 1.0, L.ptr(y), 0.0, L.ptr(z2), 1))
    out["gram_matvec_rel_diff"] = float(np.linalg.norm(G @ v - z2) / np.linalg.norm(z2))
    init = np.random.default_rng(SEED).standard_normal(n)
    r = ctypes.c_void_p()
    L.check(lib.svb_tssvd(h, nu, 2 * nu, 1000, 0.0, L.ptr(init), ctypes.byref(r)))
    lib.svb_result_free(r)
    lib.svb_synchronize()
    t0 = time.perf_counter()
    L.check(lib.svb_tssvd(h, nu, 2 * nu, 1000, 0.0, L.ptr(init), ctypes.byref(r)))
    lib.svb_synchronize()
    out["tssvd_s"] = round(time.perf_counter() - t0, 5)
    it, mp, info = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
    lib.svb_result_info(r, None, None, None, ctypes.byref(it), ctypes.byref(mp), ctypes.byref(info))
    s_t = np.zeros(nu)
    L.check(lib.svb_result_download(r, L.ptr(s_t), None, None, 0))
    lib.svb_result_free(r)
    out["tssvd"] = {"restarts": it.value, "matvecs": mp.value, "info": info.value}
    r2, it2, mp2, info2 = solve_device(sv, h, nu, init)
    lib.svb_synchronize()
    t0 = time.perf_counter()
    lib.svb_result_free(r2)
    r2, it2, mp2, info2 = solve_device(sv, h, nu, init)
    lib.svb_synchronize()
    out["irlba_s"] = round(time.perf_counter() - t0, 5)
    s_i = np.zeros(nu)
    L.check(lib.svb_result_download(r2, L.ptr(s_i), None, None, 0))
    lib.svb_result_free(r2)
    out["sigma_max_rel_diff_vs_irlba_tol1e-5"] = float(np.max(np.abs(s_t / s_i - 1)))
    lib.svb_operator_free(h)
    line = json.dumps(out)
    print(line, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gram_check_%s.json" % (sys.argv[1] if len(sys.argv) > 1 else "C2")), "w") as f:
        f.write(line + "\n")


if __name__ == "__main__":
    main()
