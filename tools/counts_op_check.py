"""tools/counts_op_check.py — full-size check + timing of the count-level operator against the explicit one (1 GPU).
usage: python tools/counts_op_check.py [C3|C2|C1] [cells]
Prints one JSON line: build times, per-product device times (CUDA events on the library stream), solve times, parity."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import severo_jl_b200 as sv  # noqa: E402
from bench import CONFIGS, SCALE_MAX, SEED, TOL, solve_device  # noqa: E402
from severo_jl_b200 import sharding  # noqa: E402

L = sv._lib


def timed_products(lib, op, trans, dx, dy, stream, reps=20):
    for _ in range(3):
        L.check(lib.svb_mul_device(op, trans, 1.0, dx, 0.0, dy))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        L.check(lib.svb_mul_device(op, trans, 1.0, dx, 0.0, dy))
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    cfg = dict(CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"])
    if len(sys.argv) > 2:
        cfg["m"] = int(sys.argv[2])
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    lib = sv.init(0)
    out = {"config": cfg}
    with torch.cuda.stream(stream):
        L.check(lib.svb_set_stream(ctypes.c_void_p(stream.cuda_stream)))
        m, n, nu = cfg["m"], cfg["n"], cfg["nu"]
        counts = sv.synthetic_counts(m, cfg["g"], cfg["nnz"], programs=cfg["programs"], fold=6.0, seed=SEED)
        libsize = np.empty(m, dtype=np.int64)
        L.check(lib.svb_row_sums(counts._h, L.ptr(libsize)))
        Y = sv.normalize_cells(counts, method="lognormalize", scale_factor=1e4)
        metric = sharding.sharded_vst_metric(counts)
        hvf = np.argsort(-metric, kind="stable")[:n]
        chv = counts.columns(hvf)
        counts.free()
        B, mu = sharding.sharded_scale_features(Y, scale_max=SCALE_MAX, features=hvf)
        Y.free()
        vals = chv.values()
        out["hvg_nnz"] = int(chv.nnz)
        out["count_hist"] = {str(k): float((vals > k).mean()) for k in (1, 2, 4, 8, 16, 32)}
        del vals
        # explicit operator
        t0 = time.perf_counter()
        h = ctypes.c_void_p()
        L.check(lib.svb_operator_create_ex(B._h, L.ptr(np.ascontiguousarray(mu)), 0, 0, ctypes.byref(h)))
        lib.svb_synchronize()
        out["explicit_build_s"] = round(time.perf_counter() - t0, 4)
        B.free()
        # count-level operator, parallel moments
        t0 = time.perf_counter()
        C = sv.CountsCenteredMatrix(chv, libsize, 1e4, SCALE_MAX, None, None, int(os.environ.get("LEVELS", "0")))
        lib.svb_synchronize()
        out["counts_build_s"] = round(time.perf_counter() - t0, 4)
        t0 = time.perf_counter()
        C2 = sv.CountsCenteredMatrix(chv, libsize, 1e4, SCALE_MAX, None, None, int(os.environ.get("LEVELS", "0")))
        lib.svb_synchronize()
        out["counts_build_warm_s"] = round(time.perf_counter() - t0, 4)
        C2.free()
        out["counts_info"] = C.info()
        out["mu_max_rel_diff"] = float(np.max(np.abs(C.mu - mu) / np.maximum(np.abs(mu), 1e-300)))
        # products
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        w = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
        y1, y2 = torch.empty(m, dtype=torch.float64, device="cuda"), torch.empty(m, dtype=torch.float64, device="cuda")
        z1, z2 = torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")
        px, pw = ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(w.data_ptr())
        out["explicit_fwd_ms"] = round(timed_products(lib, h, b"N", px, ctypes.c_void_p(y1.data_ptr()), stream), 4)
        out["explicit_adj_ms"] = round(timed_products(lib, h, b"T", pw, ctypes.c_void_p(z1.data_ptr()), stream), 4)
        out["counts_fwd_ms"] = round(timed_products(lib, C._op, b"N", px, ctypes.c_void_p(y2.data_ptr()), stream), 4)
        out["counts_adj_ms"] = round(timed_products(lib, C._op, b"T", pw, ctypes.c_void_p(z2.data_ptr()), stream), 4)
        out["fwd_rel_diff"] = float((y1 - y2).norm() / y1.norm())
        out["adj_rel_diff"] = float((z1 - z2).norm() / z1.norm())
        # solves
        init = np.random.default_rng(SEED).standard_normal(n)
        res = {}
        for name, op in (("explicit", h), ("counts", C._op)):
            for _ in range(2):
                r, it, mp, info = solve_device(sv, op, nu, init)
                lib.svb_result_free(r)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r, it, mp, info = solve_device(sv, op, nu, init)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            s = np.zeros(nu)
            L.check(lib.svb_result_download(r, L.ptr(s), None, None, 0))
            lib.svb_result_free(r)
            res[name] = s
            out[name + "_solve_s"] = round(dt, 5)
            out[name + "_solve"] = {"restarts": it, "matvecs": mp, "info": info}
        out["sigma_max_rel_diff"] = float(np.max(np.abs(res["counts"] / res["explicit"] - 1)))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
