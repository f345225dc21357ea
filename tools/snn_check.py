"""tools/snn_check.py — timing of the device Jaccard index / shared-nearest-neighbour graph (svb_jaccard_index) at scale, 1 GPU.
usage: python tools/snn_check.py n d k
Builds the exact kNN graph of n clustered points on the device (svb_knn), then times svb_jaccard_index on the resident
graph (device-synchronised on both sides, two calls per device path) and the whole host call (upload of the pattern, compute, download of
the CSC result). Spot check: 64 random columns recomputed on the host with Python set intersections. One JSON line."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import severo_jl_b200 as sv  # noqa: E402


def main():
    n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    prune = 1.0 / 15.0
    sv.init(0)
    L = sv._lib
    lib = sv.lib()
    rng = np.random.default_rng(1)
    centres = rng.standard_normal((64, d)) * 3.0
    X = np.asfortranarray(centres[rng.integers(0, 64, n)] + rng.standard_normal((n, d)))
    t0 = time.perf_counter()
    idx, _ = sv.ann(X, k)
    knn_s = time.perf_counter() - t0
    srt = np.sort(idx, axis=1)
    pattern = sp.csc_matrix((np.ones(n * k, dtype=np.int32), srt.ravel(), np.arange(0, n * k + 1, k, dtype=np.int64)), shape=(n, n))
    dN = sv.DeviceMatrix.from_host(pattern)
    times = {}
    nnz = 0
    # both device paths back to back (SVB_SNN_HASH is read per call): per-warp hash tables (default) / general enumeration
    for name, env in (("hash", "1"), ("general", "0")):
        os.environ["SVB_SNN_HASH"] = env
        times[name] = []
        for _ in range(2):
            h = ctypes.c_void_p()
            lib.svb_synchronize()
            t0 = time.perf_counter()
            L.check(lib.svb_jaccard_index(dN._h, k, prune, L.SVB_F64, ctypes.byref(h)))
            lib.svb_synchronize()
            times[name].append(round(time.perf_counter() - t0, 5))
            o = sv.DeviceMatrix(h)
            nnz = o.nnz
            o.free()
    os.environ.pop("SVB_SNN_HASH", None)
    dN.free()
    t0 = time.perf_counter()
    S = sv.jaccard_index(pattern.astype(bool), k, prune)
    host_s = time.perf_counter() - t0
    indeg = np.bincount(srt.ravel(), minlength=n).astype(np.float64)
    sets = {}

    def nb(i):
        if i not in sets:
            sets[i] = set(srt[i].tolist())
        return sets[i]

    rev = sp.csr_matrix(pattern)
    bad = 0
    for j in rng.integers(0, n, 64):
        cand = set()
        for p in srt[j]:
            cand.update(rev.indices[rev.indptr[p]:rev.indptr[p + 1]].tolist())
        ref = {}
        for i in cand:
            x = float(len(nb(i) & nb(j)))
            v = x / (k + (k - x))
            if not abs(v) <= prune:
                ref[i] = v
        a, b = S.indptr[j], S.indptr[j + 1]
        got = dict(zip(S.indices[a:b].tolist(), S.data[a:b].tolist()))
        bad += int(got != ref or list(S.indices[a:b]) != sorted(ref))
    out = {"n": n, "d": d, "k": k, "prune": prune, "knn_s": round(knn_s, 4), "jaccard_device_s": times,
           "jaccard_host_call_s": round(host_s, 4), "snn_nnz": int(nnz), "snn_nnz_per_cell": round(nnz / n, 2),
           "candidate_pairs": float((indeg ** 2).sum()), "max_indegree": int(indeg.max()),
           "spot_check_mismatching_columns_of_64": bad}
    line = json.dumps(out)
    print(line, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "snn_check_%d_%d_%d.json" % (n, d, k)), "w") as f:
        f.write(line + "\n")


if __name__ == "__main__":
    main()
