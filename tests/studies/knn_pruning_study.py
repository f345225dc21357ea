"""tests/studies/knn_pruning_study.py — CPU study (not a test, not product code): would an exact cluster-pruned search beat the
brute-force kNN kernel? PCA coordinates of a planted-program count matrix (oracle pipeline), k-means into C clusters, then the
fraction of the n^2 distance evaluations an exact search must still do when a (query, cluster) pair is skipped only if the
triangle-inequality bound d(q, centre) - radius is not below the query's k-th distance inside its own cluster.
Result (20,000 cells, 50 programs, k = 20): 10 coordinates 0.42-0.77 of n^2 for C = 512..64, 50 coordinates
0.82-0.98 — at most 2.4x / 1.2x before any overhead, so the brute-force design stays.
usage: python tests/studies/knn_pruning_study.py 20000 50"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, scipy.sparse as sp
from oracle import severo_oracle as orc
from conftest import planted_counts

def coords(m, g, progs, nu, seed=3):
    X = planted_counts(m, g, progs, seed=seed, mean_nnz=80.0)
    Y = orc.normalize_cells(X, "lognormalize", 1e4)
    S = orc.scale_features(Y, scale_max=10.0)
    rng = np.random.default_rng(0)
    r = orc.irlba(S, nu, init=rng.standard_normal(S.shape[1]), tol=1e-5)
    return np.ascontiguousarray(r.U[:, :nu] * r.S[None, :nu])

def kmeans(X, C, iters=8, seed=0):
    rng = np.random.default_rng(seed)
    cen = X[rng.choice(X.shape[0], C, replace=False)].copy()
    for _ in range(iters):
        d2 = (X*X).sum(1)[:,None] - 2*X@cen.T + (cen*cen).sum(1)[None,:]
        lab = d2.argmin(1)
        for c in range(C):
            mm = lab == c
            if mm.any(): cen[c] = X[mm].mean(0)
    d2 = (X*X).sum(1)[:,None] - 2*X@cen.T + (cen*cen).sum(1)[None,:]
    return cen, d2.argmin(1)

def study(X, k, C):
    n = X.shape[0]
    cen, lab = kmeans(X, C)
    order = np.argsort(lab, kind="stable"); X = X[order]; lab = lab[order]
    starts = np.searchsorted(lab, np.arange(C+1))
    dc = np.sqrt(np.maximum((X*X).sum(1)[:,None] - 2*X@cen.T + (cen*cen).sum(1)[None,:], 0))   # n x C
    rad = np.array([dc[starts[c]:starts[c+1], c].max() if starts[c+1] > starts[c] else 0.0 for c in range(C)])
    tau = np.full(n, np.inf)
    evals = 0
    for c in range(C):
        a, b = starts[c], starts[c+1]
        if b - a == 0: continue
        P = X[a:b]
        D = np.sqrt(np.maximum((P*P).sum(1)[:,None] - 2*P@P.T + (P*P).sum(1)[None,:], 0))
        evals += (b-a)**2
        if b - a >= k:
            tau[a:b] = np.partition(D, k-1, axis=1)[:, k-1]      # include_self: kth incl. itself
    pairs = 0
    need_total = 0
    for c in range(C):
        a, b = starts[c], starts[c+1]
        if b - a == 0: continue
        lb = dc[a:b] - rad[None, :]                  # lower bound of the distance from q to any point of cluster B
        sizes = (starts[1:] - starts[:-1])
        nq = (lb < tau[a:b, None])
        nq[:, c] = False
        pairs += (nq * sizes[None, :]).sum()
        need_total += nq.any(axis=0).sum()
    return (evals + pairs) / float(n) ** 2, need_total / C, np.isinf(tau).mean()

if __name__ == "__main__":
    m = int(sys.argv[1]); nu = int(sys.argv[2])
    t0 = time.time()
    Z = coords(m, 3000, 50, 50)
    print("coords", Z.shape, round(time.time()-t0,1), "s")
    for d in (10, 50):
        for C in (64, 128, 256, 512):
            f, avg_need, inf = study(Z[:, :d], 20, C)
            print(f"d={d} C={C}: evaluated fraction of n^2 = {f:.4f}  (speed-up bound {1/f:.1f}x), clusters visited per cluster {avg_need:.1f}, tau=inf {inf:.3f}")
