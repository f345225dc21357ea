"""Host loess (local quadratic, tricube weights) for the :vst mean-variance trend.

The reference calls the third-party package Loess.jl (``variablefeatures.jl:41,44``; un-pinned, kd-tree
vertices + interpolation). Only O(genes) work happens here, between the two device sweeps; the fit is
evaluated exactly at up to 256 vertices (quantiles of x) and interpolated linearly, which is the same
vertex-then-interpolate structure. HVG *selection* parity is un-pinned upstream (no reference test).
"""
from __future__ import annotations

import numpy as np


def _local_fit(x, y, x0, q, degree):
    d = np.abs(x - x0)
    idx = np.argpartition(d, q - 1)[:q]
    h = d[idx].max()
    if h <= 0:
        return float(np.mean(y[idx]))
    u = d[idx] / h
    w = (1.0 - u ** 3) ** 3
    w[u >= 1.0] = 0.0
    sw = np.sqrt(w)
    xc = x[idx] - x0
    cols = [np.ones_like(xc), xc]
    if degree >= 2:
        cols.append(xc * xc)
    A = np.stack(cols, axis=1) * sw[:, None]
    coef, *_ = np.linalg.lstsq(A, y[idx] * sw, rcond=None)
    return float(coef[0])


def loess_fit_predict(x, y, span=0.5, degree=2, max_vertices=256):
    """Fit y ~ loess(x) and return the fitted values at x."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.shape[0]
    if n == 0:
        return np.zeros(0)
    if n <= degree + 1:
        return y.copy()
    q = int(min(n, max(degree + 2, np.ceil(span * n))))
    xs = np.unique(x)
    if xs.shape[0] > max_vertices:
        xs = np.unique(np.quantile(x, np.linspace(0.0, 1.0, max_vertices)))
    fit = np.array([_local_fit(x, y, v, q, degree) for v in xs])
    return np.interp(x, xs, fit)
