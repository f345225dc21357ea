"""Generates tests/golden/reference_fixed_inputs.json.

Julia and libcell are absent from this image, so the reference cannot be executed. The golden vectors
are therefore the *fixed inputs of the reference's own tests* together with the quantities those tests
assert on, evaluated here by an INDEPENDENT dense numpy computation (not by the oracle):

  test/test_scaling.jl:22-45   10x5 Int64 matrix -> mean, std (ddof=1), mu./std, (X .- mu') ./ std'
  test/test_scaling.jl:72-112  4x3 operator: Q*r, 2Q*r + y0, Q'*y, 2Q'*y + r0
  test/test_input.jl:38-42,62-74  filter_counts survivors; relative-count row sums 1 and 10
plus the order-dependent bit patterns of the sequential Welford (scaling.jl:18-34) computed by a literal
pure-Python transcription, stored as hex floats.

Run:  python tests/golden/make_golden.py
"""
import json
import os

import numpy as np

X = np.array([[0, 0, 0, 3, 0], [0, 1, 0, 0, 6], [5, 0, 0, 0, 0], [3, 0, 0, 0, 0], [0, 0, 6, 2, 0],
              [0, 0, 0, 0, 0], [0, 0, 0, 0, 2], [0, 3, 0, 0, 0], [0, 0, 0, 0, 3], [2, 0, 0, 0, 0]], dtype=np.int64)


def welford(col, n):
    vals = [float(v) for v in col if v != 0]
    count = n - len(vals)
    mu = s = 0.0
    for v in vals:
        count += 1
        delta = v - mu
        mu += delta / count
        s += delta * (v - mu)
    return mu, s / (n - 1)


def main():
    out = {"X": X.tolist()}
    mean = X.mean(axis=0)
    std = X.std(axis=0, ddof=1)
    out["mean"] = mean.tolist()
    out["std"] = std.tolist()
    out["mu_over_std"] = (mean / std).tolist()
    out["scaled_dense"] = ((X - mean[None, :]) / std[None, :]).tolist()
    w = [welford(X[:, j], X.shape[0]) for j in range(X.shape[1])]
    out["welford_mu_hex"] = [float(a).hex() for a, _ in w]
    out["welford_var_hex"] = [float(b).hex() for _, b in w]
    out["welford_mu_over_sd_hex"] = [float(a / np.sqrt(b)).hex() for a, b in w]
    # filter_counts(min_features=1, min_cells=2, min_umi=2) survivors (test_input.jl:38-42), 0-based
    out["filter_cells"] = [0, 1, 2, 3, 4, 7, 8]
    out["filter_genes"] = [0, 1, 3, 4]
    Xf = X[np.ix_(out["filter_cells"], out["filter_genes"])]
    out["relcounts_dense"] = (Xf / Xf.sum(axis=1, keepdims=True)).tolist()
    # operator (test_scaling.jl:72-112)
    A = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1]], dtype=float)
    mu = np.array([0.1, 0.2, 0.3])
    r = np.array([0.991, 0.228, 0.291])
    y0 = np.array([0.820, 0.430, 0.264, 0.789])
    Q = A - mu[None, :]
    y = 2 * Q @ r + y0
    out["op"] = {"A": A.tolist(), "mu": mu.tolist(), "r": r.tolist(), "y0": y0.tolist(), "Qr": (Q @ r).tolist(),
                 "y": y.tolist(), "Qty": (Q.T @ y).tolist(), "r2": (2 * Q.T @ y + r).tolist()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixed_inputs.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
