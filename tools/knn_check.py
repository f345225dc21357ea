"""tools/knn_check.py — timing of the exact device k-nearest-neighbour search (svb_knn) at scale, 1 GPU.
usage: python tools/knn_check.py n d k [metric]
Prints one JSON line: wall time of the call (host buffers in, host buffers out), n^2*D fp64 FMAs per second against the
fp64 pipe, and a spot check of 64 random queries against a numpy brute-force search."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import severo_jl_b200 as sv  # noqa: E402


def main():
    n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    metric = sys.argv[4] if len(sys.argv) > 4 else "euclidean"
    sv.init(0)
    rng = np.random.default_rng(1)
    centres = rng.standard_normal((64, d)) * 3.0
    X = np.asfortranarray(centres[rng.integers(0, 64, n)] + rng.standard_normal((n, d)))
    import ctypes
    L = sv._lib
    lib = sv.lib()
    sv.ann(X[:4096], k, metric=metric)  # warm-up: module load, allocations
    def padded(d, mma, exact=True):      # knn_padded_dims (csrc/knn.cu)
        if exact and not mma and d in (9, 10, 11, 12, 17, 18, 19, 20, 49, 50):
            return {9: 10, 10: 10, 11: 12, 12: 12, 49: 50, 50: 50}.get(d, 20)
        return (d + 7) // 8 * 8 if d <= 64 else (96 if d <= 96 else 128)

    D = padded(d, False)
    out = {"n": n, "d": d, "D": D, "k": k, "metric": metric}
    # the kernel variants back to back in one process (same box, same clocks): SVB_KNN_Q / SVB_KNN_MMA are read per call.
    # seconds = wall time of the whole call (pageable host buffers in and out); kernel_ms = CUDA events around the search kernel
    variants = [("default", {}), ("simt_q1", {"SVB_KNN_Q": "1", "SVB_KNN_MMA": "0"}), ("simt_q2", {"SVB_KNN_Q": "2", "SVB_KNN_MMA": "0"}),
                ("dmma", {"SVB_KNN_MMA": "1"}), ("pad8", {"SVB_KNN_WIDTHS": "0"})]
    if os.environ.get("KNN_VARIANTS") == "default":
        variants = variants[:1]
    elif os.environ.get("KNN_VARIANTS") == "widths":     # exact-width instance against the multiple-of-8 one
        variants = [variants[0], variants[4]]
    lib.svb_profile_enable(1)
    for name, env in variants:
        for key in ("SVB_KNN_Q", "SVB_KNN_MMA", "SVB_KNN_WIDTHS"):
            os.environ.pop(key, None)
        os.environ.update(env)
        lib.svb_profile_reset()
        t0 = time.perf_counter()
        idx, dist = sv.ann(X, k, metric=metric)
        dt = time.perf_counter() - t0
        ms = (ctypes.c_double * 6)()
        ln = (ctypes.c_int64 * 6)()
        by = (ctypes.c_double * 6)()
        L.check(lib.svb_profile_get(ms, ln, by))
        kms = ms[4]
        out[name] = {"seconds": round(dt, 4), "kernel_ms": round(kms, 2), "D": padded(d, name == "dmma", name != "pad8"),
                     "kernel_fp64_tflops": round(2.0 * n * n * padded(d, name == "dmma", name != "pad8") / (kms * 1e-3) / 1e12, 2)}
    for key in ("SVB_KNN_Q", "SVB_KNN_MMA", "SVB_KNN_WIDTHS"):
        os.environ.pop(key, None)
    lib.svb_profile_enable(0)
    idx, dist = sv.ann(X, k, metric=metric)
    bad = 0
    for i in rng.integers(0, n, 64):
        if metric == "euclidean":
            dd = np.sqrt(((X - X[i]) ** 2).sum(axis=1))
        else:
            dd = np.maximum(1.0 - (X @ X[i]) / (np.linalg.norm(X, axis=1) * np.linalg.norm(X[i])), 0.0)
        dd[i] = -1.0
        ref = np.argsort(dd, kind="stable")[:k]
        bad += int(set(ref) != set(idx[i]))
    out["spot_check_mismatching_rows_of_64"] = bad
    line = json.dumps(out)
    print(line, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "knn_check_%d_%d.json" % (n, d)), "w") as f:
        f.write(line + "\n")


if __name__ == "__main__":
    main()
