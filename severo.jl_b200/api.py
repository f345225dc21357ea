"""Host-side mirror of the reference's Julia interface for the IRLBA-PCA hot path.

Same names, argument meaning, defaults and error behaviour as ExaScience/Severo.jl so the parity
tests read like the reference's own tests; every compute step is a call into the C ABI of
``libsevero_b200.so`` (include/severo_b200.h). Julia is not available in this image, so this Python
layer plays the role the Julia overlay (julia/SeveroB200.jl) plays for a Julia user.

reference                                              here
---------                                              ----
normalize_cells(X; method, scale_factor, dtype)        normalize_cells        (src/normalize.jl:40-79)
find_variable_features(X, n; method=:vst, ...)         find_variable_features (src/variablefeatures.jl:128-161)
scale_features(X; scale_max, dtype, features)          scale_features         (src/scaling.jl:335-357)
CenteredMatrix(A, mu), mul!, adjoint, convert          CenteredMatrix         (src/scaling.jl:219-314)
Severo.irlba(A, nu; init, tol, svtol, maxit), restart  irlba                  (src/irlba.jl:47-99)
_pca / pca / embedding(X, k; method=:pca, ...)         _pca / pca / embedding (src/embedding.jl:46-94,202-212)
svd_flip!                                              svd_flip               (src/utils.jl:215-228)
"""
from __future__ import annotations

import ctypes
import warnings
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import scipy.sparse as sp

from . import _lib as L
from .loess import loess_fit_predict

__all__ = [
    "DeviceMatrix", "NamedArray", "convert_counts", "filter_cells", "filter_features", "filter_counts", "normalize_cells", "mean_var", "mean_std",
    "standardized_var_clipped", "find_variable_features", "scale_features", "CenteredMatrix", "CountsCenteredMatrix",
    "scale_features_counts", "irlba", "mul_sparse_vector", "mul_sparse_dense", "gram", "tssvd", "ann", "nearest_neighbours", "jaccard_index", "shared_nearest_neighbours",
    "SVD", "svd_flip", "pca", "embedding", "LinearEmbedding", "synthetic_counts",
    "init_devices", "shutdown_devices", "devices_info", "irlba_devices", "pca_counts_devices",
]

_DT = {np.dtype(np.float32): L.SVB_F32, np.dtype(np.float64): L.SVB_F64,
       np.dtype(np.int32): L.SVB_I32, np.dtype(np.int64): L.SVB_I64}
_NP = {L.SVB_F32: np.float32, L.SVB_F64: np.float64, L.SVB_I32: np.int32, L.SVB_I64: np.int64}


# ------------------------------------------------------------------------------------------------
# containers
# ------------------------------------------------------------------------------------------------
@dataclass
class NamedArray:
    """Minimal stand-in for NamedArrays.NamedArray: an array plus per-dimension names."""
    array: object
    names: tuple
    dimnames: tuple = ("cells", "features")

    @property
    def shape(self):
        return self.array.shape


def convert_counts(X, barcodes: Optional[Sequence[str]] = None, features: Optional[Sequence[str]] = None):
    """src/input.jl:776-809: label a counts matrix ``(barcodes, features)`` with dims (:cells, :features)."""
    m, n = X.shape
    if barcodes is None:
        barcodes = [f"cell-{i + 1}" for i in range(m)]
    if features is None:
        features = [f"gene-{i + 1}" for i in range(n)]
    X = sp.csc_matrix(X)
    if not np.issubdtype(X.dtype, np.integer):
        if np.all(np.round(X.data) == X.data):
            X = X.astype(np.int64)
        else:
            warnings.warn("non-integer counts")  # input.jl:778
    return NamedArray(X, (list(barcodes), list(features)), ("cells", "features"))


class DeviceMatrix:
    """A sparse matrix resident in HBM (svb_matrix_t): CSC, rows = cells, columns = genes."""

    def __init__(self, handle):
        self._h = ctypes.c_void_p(handle) if not isinstance(handle, ctypes.c_void_p) else handle
        nr, nc, nnz, vt = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
        L.check(L.lib().svb_matrix_info(self._h, nr, nc, nnz, vt))
        self.shape = (nr.value, nc.value)
        self.nnz = nnz.value
        self.vtype = vt.value
        self.dtype = np.dtype(_NP[vt.value])

    @classmethod
    def from_host(cls, X):
        X = sp.csc_matrix(X)
        if not X.has_sorted_indices:
            X = X.copy()
            X.sort_indices()
        if X.dtype not in _DT:
            X = X.astype(np.int64 if np.issubdtype(X.dtype, np.integer) else np.float64)
        colptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
        if X.indices.dtype == np.int32:
            rowval, rtype = np.ascontiguousarray(X.indices), L.SVB_I32
        else:
            rowval, rtype = np.ascontiguousarray(X.indices, dtype=np.int64), L.SVB_I64
        nz = np.ascontiguousarray(X.data)
        h = ctypes.c_void_p()
        L.check(L.lib().svb_csc_upload(X.shape[0], X.shape[1], L.ptr(colptr), L.ptr(rowval), rtype, L.ptr(nz),
                                        _DT[nz.dtype], 0, ctypes.byref(h)))
        return cls(h)

    @classmethod
    def from_julia_arrays(cls, nrow, ncol, colptr, rowval, nzval):
        """1-based Int64 ``colptr`` / ``rowval`` exactly as a Julia SparseMatrixCSC{T,Int64} holds them."""
        colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        rowval = np.ascontiguousarray(rowval, dtype=np.int64)
        nzval = np.ascontiguousarray(nzval)
        h = ctypes.c_void_p()
        L.check(L.lib().svb_csc_upload(nrow, ncol, L.ptr(colptr), L.ptr(rowval), L.SVB_I64, L.ptr(nzval),
                                        _DT[nzval.dtype], 1, ctypes.byref(h)))
        return cls(h)

    def to_host(self, dtype=None):
        m, n = self.shape
        colptr = np.empty(n + 1, dtype=np.int64)
        rowval = np.empty(self.nnz, dtype=np.int64)
        dt = np.dtype(dtype) if dtype is not None else (np.dtype(np.int64) if self.vtype == L.SVB_I32 else self.dtype)
        nz = np.empty(self.nnz, dtype=dt)
        L.check(L.lib().svb_matrix_download(self._h, L.ptr(colptr), L.ptr(rowval), L.ptr(nz), _DT[dt], 0))
        return sp.csc_matrix((nz, rowval, colptr), shape=(m, n))

    def values(self, dtype=None):
        dt = np.dtype(dtype) if dtype is not None else (np.dtype(np.int64) if self.vtype == L.SVB_I32 else self.dtype)
        nz = np.empty(self.nnz, dtype=dt)
        L.check(L.lib().svb_matrix_download(self._h, None, None, L.ptr(nz), _DT[dt], 0))
        return nz

    def colptr(self):
        cp = np.empty(self.shape[1] + 1, dtype=np.int64)
        L.check(L.lib().svb_matrix_download(self._h, L.ptr(cp), None, None, L.SVB_F64, 0))
        return cp

    def columns(self, idx):
        """X[:, idx] (docs/src/pbmc.md:121)."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        h = ctypes.c_void_p()
        L.check(L.lib().svb_column_subset(self._h, L.ptr(idx), idx.shape[0], 0, ctypes.byref(h)))
        return DeviceMatrix(h)

    def rows(self, row0, row1):
        h = ctypes.c_void_p()
        L.check(L.lib().svb_row_slice(self._h, int(row0), int(row1), ctypes.byref(h)))
        return DeviceMatrix(h)

    def transpose(self):
        h = ctypes.c_void_p()
        L.check(L.lib().svb_transpose(self._h, ctypes.byref(h)))
        return DeviceMatrix(h)

    def free(self):
        if self._h is not None and self._h.value:
            L.load().svb_matrix_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _unwrap(X):
    """-> (payload, names, dimnames) for NamedArray / plain inputs."""
    if isinstance(X, NamedArray):
        return X.array, X.names, X.dimnames
    return X, None, None


def _to_device(A):
    return (A, False) if isinstance(A, DeviceMatrix) else (DeviceMatrix.from_host(A), True)


def _rewrap(result, names, dimnames):
    return result if names is None else NamedArray(result, names, dimnames)


# ------------------------------------------------------------------------------------------------
# filtering.jl
# ------------------------------------------------------------------------------------------------
def _filter(X, min_cells, min_features, min_feature_count, min_umi):
    A, names, dimnames = _unwrap(X)
    dA, temp = _to_device(A)
    if dA.vtype != L.SVB_I32:
        raise TypeError("filtering expects a count matrix (integer values)")
    m, n = dA.shape
    ci, fi = np.zeros(m, dtype=np.uint8), np.zeros(n, dtype=np.uint8)
    h = ctypes.c_void_p()
    L.check(L.lib().svb_filter_counts(dA._h, int(min_cells), int(min_features), int(min_feature_count), int(min_umi),
                                       L.ptr(ci), L.ptr(fi), ctypes.byref(h)))
    out = DeviceMatrix(h)
    CI, FI = ci.astype(bool), fi.astype(bool)
    if temp:
        host = out.to_host()
        out.free()
        dA.free()
        out = host
    if names is not None:
        out = NamedArray(out, ([b for b, k in zip(names[0], CI) if k], [f for f, k in zip(names[1], FI) if k]), dimnames)
    return out, CI, FI


def filter_features(A, min_cells=0):
    """filtering.jl:15-20,50-54: keep the features detected (count > 0) in at least ``min_cells`` cells.
    Plain matrix in -> (A[:, FI], FI) like the unlabelled method; NamedArray in -> labelled matrix."""
    out, _, FI = _filter(A, min_cells, 0, 0, 0)
    return out if isinstance(A, NamedArray) else (out, FI)


def filter_cells(A, min_features=0, min_feature_count=0, min_umi=0):
    """filtering.jl:22-35,72-76: keep the cells with at least ``min_features`` features above ``min_feature_count`` and,
    when ``min_umi > 0``, a UMI total strictly above ``min_umi``."""
    out, CI, _ = _filter(A, 0, min_features, min_feature_count, min_umi)
    return out if isinstance(A, NamedArray) else (out, CI)


def filter_counts(A, min_cells=0, min_features=0, min_feature_count=0, min_umi=0):
    """filtering.jl:101-106: filter_cells, then filter_features on the remaining cells."""
    out, CI, FI = _filter(A, min_cells, min_features, min_feature_count, min_umi)
    return out if isinstance(A, NamedArray) else (out, CI, FI)


# ------------------------------------------------------------------------------------------------
# normalize.jl
# ------------------------------------------------------------------------------------------------
def normalize_cells(X, method="lognormalize", scale_factor=1.0, dtype=np.float64):
    """normalize.jl:40-55,76-79. Host (scipy / NamedArray) in -> host out; DeviceMatrix in -> DeviceMatrix out."""
    method = str(method)
    if method == "lognormalize":
        code = L.NORM_LOGNORMALIZE
    elif method == "relativecounts":
        code = L.NORM_RELATIVECOUNTS
    else:
        raise ValueError(f"unknown normalization method: {method}")
    dtype = np.dtype(dtype)
    if dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError("dtype must be a float type")
    A, names, dimnames = _unwrap(X)
    dA, temp = _to_device(A)
    if dA.vtype != L.SVB_I32:
        raise TypeError("normalize_cells expects integer counts")
    h = ctypes.c_void_p()
    L.check(L.lib().svb_normalize(dA._h, code, float(scale_factor), _DT[dtype], ctypes.byref(h)))
    out = DeviceMatrix(h)
    if temp:
        host = out.to_host()
        out.free()
        dA.free()
        return _rewrap(host, names, dimnames)
    return out


# ------------------------------------------------------------------------------------------------
# scaling.jl — moments
# ------------------------------------------------------------------------------------------------
def mean_var(A):
    """scaling.jl:132-147 mean_var(A::SparseMatrixCSC): order-exact Welford per gene (device)."""
    A, _, _ = _unwrap(A)
    dA, temp = _to_device(A)
    n = dA.shape[1]
    mu = np.empty(n)
    var = np.empty(n)
    L.check(L.lib().svb_mean_var(dA._h, L.ptr(mu), L.ptr(var)))
    if temp:
        dA.free()
    if dA.vtype == L.SVB_F32:
        return mu.astype(np.float32), var.astype(np.float32)
    return mu, var


def mean_std(A):
    """scaling.jl:119-130."""
    mu, var = mean_var(A)
    return mu, np.sqrt(var)


def standardized_var_clipped(A, mu, sd, vmax=None):
    """variablefeatures.jl:21-28 on the device."""
    A, _, _ = _unwrap(A)
    dA, temp = _to_device(A)
    n = dA.shape[1]
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    sd = np.ascontiguousarray(sd, dtype=np.float64)
    out = np.zeros(n)
    L.check(L.lib().svb_stdvar_clipped(dA._h, L.ptr(mu), L.ptr(sd), -1.0 if vmax is None else float(vmax), L.ptr(out)))
    if temp:
        dA.free()
    return out


def _call_trend(fn, mu, sd):
    """``expected_std_fn(mu)`` or ``expected_std_fn(mu, sd)``: a deterministic replacement of the loess trend of :vst."""
    import inspect
    try:
        nargs = len([p for p in inspect.signature(fn).parameters.values()
                     if p.default is inspect.Parameter.empty and p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)])
    except (TypeError, ValueError):
        nargs = 1
    return fn(mu, sd) if nargs >= 2 else fn(mu)


def variance_stabilizing_transformation(A, loess_span=0.5, expected_std_fn=None, dtype=np.float64):
    """variablefeatures.jl:34-50. The two data sweeps run on the device; the loess fit between them
    is host code (Loess.jl in the reference: third-party and un-pinned). A deterministic
    ``expected_std_fn(mu[, sd])`` may replace the loess for the large synthetic configurations.
    ``dtype`` (variablefeatures.jl:34,37): Float64 (default) or Float32 — the per-gene moments ``mean_std(dtype, A)`` are then
    the reference's Float32 Welford (device kernel on the Float32 copy of the counts, exact for counts < 2^24) and the
    trend is fitted in Float32; the clipped variance is accumulated in Float64 as in the reference (``zeros(size(A,2))``),
    evaluated from the Float32 moments (the reference rounds each ``(x-mu)/std`` to Float32 first: <= 1e-7 relative apart)."""
    dtype = np.dtype(dtype)
    if dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError("dtype must be Float32 or Float64")
    if dtype == np.float32:
        P, _, _ = _unwrap(A)
        if isinstance(P, DeviceMatrix):
            P = P.to_host()
        mu, sd = mean_std(sp.csc_matrix(P).astype(np.float32))
        mu, sd = np.asarray(mu, dtype=np.float32), np.asarray(sd, dtype=np.float32)
    else:
        mu, sd = mean_std(A)
        mu = np.asarray(mu, dtype=np.float64)
        sd = np.asarray(sd, dtype=np.float64)
    non_const = sd > 0
    expected = sd.copy()
    if expected_std_fn is not None:
        expected[non_const] = _call_trend(expected_std_fn, mu[non_const], sd[non_const])
    else:
        xs, ys = np.log10(mu[non_const]), np.log10(sd[non_const])
        expected[non_const] = (10.0 ** loess_fit_predict(xs, ys, span=float(dtype.type(loess_span)))).astype(dtype)
    expected = np.where(np.isnan(expected), 0.0, expected)  # nan2zero! variablefeatures.jl:30-32
    return standardized_var_clipped(A, mu, expected)


def _nan2zero(v):
    return np.where(np.isnan(v), 0.0, v)                      # variablefeatures.jl:30-32


def _log_vmr(mu, var):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log1p(mu), np.log(var / mu)                 # scaling.jl:190-193


def _cut(v, nbreaks, method="width"):
    """utils.jl:140-156 + `_cut!` :110-128 (right = true): (breaks, 1-based bin labels, 0 = outside)."""
    v = np.asarray(v, dtype=np.float64)
    method = str(method)
    if method == "width":
        from fractions import Fraction
        lo, hi = float(v.min()), float(v.max())
        dx = hi - lo
        # collect(range(lo, hi, length = nbreaks + 1)): Julia's twice-precision range, restated exactly (nbreaks + 1 values)
        breaks = np.array([float(Fraction(lo) + (Fraction(hi) - Fraction(lo)) * i / nbreaks) for i in range(nbreaks + 1)])
        breaks[0] -= dx / 1000
        breaks[-1] += dx / 1000
    elif method == "frequency":
        breaks = np.quantile(v, np.linspace(0.0, 1.0, nbreaks))
    else:
        raise ValueError(f"unknown binning method: {method}")
    idx = np.searchsorted(breaks, v, side="right")            # searchsortedlast
    on_break = breaks[np.maximum(idx, 1) - 1] == v
    labels = np.where((v >= breaks[0]) & (v <= breaks[-1]), idx - (on_break & (idx > 0)), 0)
    return breaks, labels.astype(np.int64)


def _mean_std_labels(x, labels, nlabels):
    """scaling.jl:89-112 on a dense vector: Welford per label in index order, unbiased variance, sqrt."""
    if labels.size and labels.min() < 1:
        # :frequency binning labels the smallest mean 0 (utils.jl:120-122) and the reference then indexes out of bounds
        raise IndexError("meanvarplot: a feature fell outside every bin (label 0) — the reference fails here as well")
    mu, var, n = np.zeros(nlabels), np.zeros(nlabels), np.zeros(nlabels, dtype=np.int64)
    for v, k in zip(x.tolist(), (labels - 1).tolist()):
        n[k] += 1
        delta = v - mu[k]
        mu[k] += delta / n[k]
        var[k] += delta * (v - mu[k])
    with np.errstate(divide="ignore", invalid="ignore"):
        return mu, np.sqrt(var / (n - 1))


def _dispersion_metric(mu, var):
    """select_dispersion, variablefeatures.jl:73-76."""
    return _nan2zero(_log_vmr(mu, var)[1])


def _meanvarplot_metric(mu, var, num_bins=20, binning_method="width"):
    """select_meanvarplot, variablefeatures.jl:78-92: z-score of the dispersion inside equal-width bins of log1p(mean)."""
    lmu, disp = _log_vmr(mu, var)
    lmu, disp = _nan2zero(lmu), _nan2zero(disp)
    _, bins = _cut(lmu, int(num_bins), binning_method)
    bin_mean, bin_std = _mean_std_labels(disp, bins, int(num_bins))
    with np.errstate(divide="ignore", invalid="ignore"):
        return _nan2zero((disp - bin_mean[bins - 1]) / bin_std[bins - 1])


def _saunders_metric(mu, var, trx_per_cell, ncells, alpha_thresh=0.99):
    """select_features_saunders, variablefeatures.jl:54-71."""
    from scipy.stats import norm as _normal
    ngenes = mu.shape[0]
    nolan = float(np.mean(1.0 / np.asarray(trx_per_cell, dtype=np.float64)))
    z = float(_normal.ppf(1.0 - (float(alpha_thresh) / ngenes) / 2.0))      # qnorm, :52
    with np.errstate(divide="ignore", invalid="ignore"):
        upper = mu + z * np.sqrt(mu * nolan / ncells)
        J = (var / nolan) > upper
        metric = np.zeros(ngenes)
        metric[J] = np.log10(var[J]) - np.log10(mu[J] * nolan)
    return metric


def _variable_feature_metric(A, method, kw):
    """The selectors next to :vst (variablefeatures.jl:135-155). Data sweeps on the device: row_norm(counts, 1) =
    ``normalize_cells(:relativecounts)``, ``mean_var`` of it, the UMI totals; the rest is gene-length host arithmetic."""
    kw = dict(kw)
    norm = kw.pop("norm", None)
    if norm is not None:
        norm = norm.array if isinstance(norm, NamedArray) else norm
    dA, temp = _to_device(A)
    dN = None
    try:
        if norm is None:
            dN = normalize_cells(dA, method="relativecounts", scale_factor=1.0)      # row_norm(counts.array, one(dtype)) :140
            norm = dN
        mu, var = mean_var(norm)
        mu, var = np.asarray(mu, dtype=np.float64), np.asarray(var, dtype=np.float64)
        if method == "saunders":
            m = dA.shape[0]
            trx = np.zeros(m, dtype=np.int64)
            L.check(L.lib().svb_row_sums(dA._h, L.ptr(trx)))
            metric = _saunders_metric(mu, var, trx, m, kw.pop("alpha_thresh", 0.1))   # call-site default :144
        elif method == "dispersion":
            kw.pop("num_bins", None), kw.pop("binning_method", None)                  # read and unused upstream too (:146-148)
            metric = _dispersion_metric(mu, var)
        else:
            metric = _meanvarplot_metric(mu, var, kw.pop("num_bins", 20), kw.pop("binning_method", "width"))
    finally:
        if dN is not None:
            dN.free()
        if temp:
            dA.free()
    return metric


def find_variable_features(counts, nfeatures=2000, method="vst", **kw):
    """variablefeatures.jl:128-161: ``:vst`` (default), ``:dispersion``, ``:meanvarplot``, ``:saunders``; ``norm=`` passes a
    precomputed normalised matrix to the last three (:136-141); ``dtype=`` (Float32 | Float64, :128) goes to the :vst moments.
    Returns 0-based gene indices ordered by DEcreasing metric (partialsortperm(..., rev=true), :159)."""
    method = str(method)
    A, names, dimnames = _unwrap(counts)
    if method == "vst":
        metric = variance_stabilizing_transformation(A, **kw)
    elif method in ("dispersion", "meanvarplot", "saunders"):
        metric = _variable_feature_metric(A, method, kw)
    else:
        raise ValueError(f"unknown selection method: {method}")
    nfeatures = min(int(nfeatures), metric.shape[0])
    selected = np.argsort(-metric, kind="stable")[:nfeatures]
    if names is None:
        return selected
    return NamedArray(selected, ([names[1][i] for i in selected],), (dimnames[1],))


# ------------------------------------------------------------------------------------------------
# scaling.jl — scale_features / CenteredMatrix
# ------------------------------------------------------------------------------------------------
def scale_features(X, scale_max=np.inf, dtype=None, features=None):
    """scaling.jl:335-357. Returns CenteredMatrix(B, mu) with mu = mean/std (the reference's stored value)."""
    A, names, dimnames = _unwrap(X)
    if features is not None:
        fidx = np.asarray(features.array if isinstance(features, NamedArray) else features, dtype=np.int64)
        if names is not None:
            names = (names[0], [names[1][i] for i in fidx])
    dA, temp = _to_device(A)
    if features is not None:
        sub = dA.columns(fidx)
        if temp:
            dA.free()
        dA, temp_sub = sub, True
    else:
        temp_sub = False
    if dtype is None:
        dtype = np.float32 if dA.vtype == L.SVB_F32 else np.float64   # dtype=T for float input, Float64 for counts
    dtype = np.dtype(dtype)
    n = dA.shape[1]
    mu = np.empty(n)
    h = ctypes.c_void_p()
    L.check(L.lib().svb_scale(dA._h, float(scale_max), _DT[dtype], ctypes.byref(h), L.ptr(mu)))
    B = DeviceMatrix(h)
    if temp or temp_sub:
        dA.free()
    if temp:  # host in -> host-visible result, device copy kept for the operator
        host = B.to_host()
        Bn = _rewrap(host, names, dimnames)
        mun = mu.astype(dtype) if names is None else NamedArray(mu.astype(dtype), (names[1],), (dimnames[1],))
        C = CenteredMatrix(Bn, mun)
        C._dev = B
        return C
    return CenteredMatrix(B, mu.astype(dtype))


class CenteredMatrix:
    """scaling.jl:219-232: S = A - 1*mu'. ``A`` may be a scipy CSC (cells x genes), a scipy CSR obtained as
    ``X.T`` of a CSC (the lazy ``Adjoint`` of test_irlba.jl:111), a DeviceMatrix, a dense ndarray, or a
    NamedArray of those. Products run on the GPU through svb_mul / svb_irlba."""

    def __init__(self, A, mu, transposed=False, storage=None):
        self.A = A
        self._storage = {None: 0, "f32": L.SVB_F32, "f64": L.SVB_F64}[storage]  # value width of the device layouts
        self.mu = mu
        payload, _, _ = _unwrap(A)
        self._transposed = bool(transposed)
        if sp.issparse(payload) and payload.format == "csr" and not transposed:
            # X' of a CSC: keep the parent, mark the operator as its adjoint
            self._parent = sp.csc_matrix((payload.data, payload.indices, payload.indptr),
                                         shape=(payload.shape[1], payload.shape[0]))
            self._transposed = True
        else:
            self._parent = payload
        pm, pn = self._parent.shape
        self.shape = (pn, pm) if self._transposed else (pm, pn)
        muv = None if mu is None else np.asarray(mu.array if isinstance(mu, NamedArray) else mu, dtype=np.float64)
        if muv is not None and muv.shape[0] != self.shape[1]:
            raise AssertionError("n == length(mu)")  # scaling.jl:226
        self._mu = muv
        self._dev = None
        self._op = None

    # -- names(C) etc. (scaling.jl:316-319)
    @property
    def names(self):
        return self.A.names if isinstance(self.A, NamedArray) else None

    @property
    def dtype(self):
        pd = np.dtype(getattr(self._parent, "dtype", np.float64))
        return np.promote_types(pd if pd.kind == "f" else np.float64, np.float64 if self._mu is None else np.float64)

    def _operator(self):
        if self._op is None:
            lib = L.lib()
            h = ctypes.c_void_p()
            mu = None if self._mu is None else np.ascontiguousarray(self._mu)
            P = self._parent
            if isinstance(P, np.ndarray):
                Pf = np.asfortranarray(P, dtype=np.float64)
                L.check(lib.svb_operator_create_dense(Pf.shape[0], Pf.shape[1], L.ptr(Pf), Pf.shape[0], L.ptr(mu),
                                                      int(self._transposed), ctypes.byref(h)))
            else:
                dev = self._dev if self._dev is not None else (P if isinstance(P, DeviceMatrix) else None)
                temp = dev is None
                if temp:
                    Ph = sp.csc_matrix(P)
                    if Ph.dtype.kind != "f":
                        Ph = Ph.astype(np.float64)
                    dev = DeviceMatrix.from_host(Ph)
                L.check(lib.svb_operator_create_ex(dev._h, L.ptr(mu), int(self._transposed), self._storage, ctypes.byref(h)))
                if temp:
                    dev.free()
            self._op = h
        return self._op

    @property
    def T(self):
        return _AdjointCentered(self)

    adjoint = T

    def mul(self, v, alpha=1.0, beta=0.0, y=None, trans=False):
        """mul!(y, S, v, alpha, beta) (scaling.jl:245-250,259-264) / mul!(y, S', v, ...) (:252-257,266-272)."""
        v = np.asfortranarray(v, dtype=np.float64)
        m, n = self.shape
        inL, outL = (m, n) if trans else (n, m)
        if v.shape[0] != inL:
            raise ValueError("DimensionMismatch")
        k = 1 if v.ndim == 1 else v.shape[1]
        if y is None:
            y = np.zeros((outL,) if v.ndim == 1 else (outL, k), order="F")
            beta = 0.0
        elif not (y.flags.f_contiguous and y.dtype == np.float64):
            raise ValueError("y must be a column-major Float64 array")
        L.check(L.lib().svb_mul(self._operator(), b"T" if trans else b"N", float(alpha), L.ptr(v), float(beta),
                                 L.ptr(y), k))
        return y

    def __matmul__(self, v):
        return self.mul(v)

    def to_dense(self):
        """convert(Matrix, C) (scaling.jl:298-309) — host helper for tests."""
        P = self._parent
        if isinstance(P, DeviceMatrix):
            P = P.to_host()
        X = np.asarray(P.toarray() if sp.issparse(P) else P, dtype=np.float64)
        X = X.T.copy() if self._transposed else X.copy()
        if self._mu is not None:
            X -= self._mu[None, :]
        return X

    def free(self):
        if self._op is not None:
            L.load().svb_operator_free(self._op)
            self._op = None
        if self._dev is not None:
            self._dev.free()
            self._dev = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CountsCenteredMatrix(CenteredMatrix):
    """The operator of ``scale_features(normalize_cells(X, :lognormalize)[:, hvf]; scale_max)`` held over the RAW COUNTS:
    the scaled matrix (scaling.jl:199-217) is never materialised, a nonzero is one 16-bit code in HBM and its value
    ``min(log1p(sf*c/s_i)/sd_j, scale_max + mean_j/sd_j)`` is rebuilt inside the product kernels (csrc/factored.cu).
    Same interface as CenteredMatrix (mul, T, irlba, pca). ``counts``: DeviceMatrix of int32 counts, cells x HVGs."""

    def __init__(self, counts, libsize, scale_factor, scale_max=np.inf, mean=None, var=None, levels=0, names=None):
        if not isinstance(counts, DeviceMatrix) or counts.vtype != L.SVB_I32:
            raise TypeError("CountsCenteredMatrix expects a DeviceMatrix of integer counts")
        m, n = counts.shape
        libsize = np.ascontiguousarray(libsize, dtype=np.int64)
        if libsize.shape[0] != m:
            raise ValueError("libsize must have one entry per cell")
        self.A = counts
        self._names = names
        self._parent = counts
        self._transposed = False
        self.shape = (m, n)
        self._dev = None
        self._storage = 0
        self.libsize, self.scale_factor, self.scale_max = libsize, float(scale_factor), float(scale_max)
        mean_c = None if mean is None else np.ascontiguousarray(mean, dtype=np.float64)
        var_c = None if var is None else np.ascontiguousarray(var, dtype=np.float64)
        mu = np.empty(n)
        h = ctypes.c_void_p()
        L.check(L.lib().svb_operator_create_counts(counts._h, L.ptr(libsize), self.scale_factor, L.ptr(mean_c), L.ptr(var_c),
                                                    self.scale_max, int(levels), L.ptr(mu), ctypes.byref(h)))
        self._op = h
        self._mu = mu
        self.mu = mu

    @property
    def names(self):
        return self._names

    def info(self):
        lv = ctypes.c_int()
        v = [ctypes.c_int64() for _ in range(5)]
        L.check(L.lib().svb_operator_counts_info(self._op, lv, *v))
        fr, ar, al = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        fp, apass = ctypes.c_double(), ctypes.c_double()
        L.check(L.lib().svb_operator_counts_layout(self._op, fr, fp, ar, al, apass))
        return dict(levels=lv.value, tile_cells=v[0].value, nnz_coded=v[1].value, nnz_exception=v[2].value,
                    fwd_chunks=v[3].value, adj_chunks=v[4].value, fwd_replicas=fr.value, fwd_passes_per_set=fp.value,
                    adj_replicas=ar.value, adj_replicated_levels=al.value, adj_passes_per_set=apass.value)

    def _operator(self):
        if self._op is None:
            raise L.SeveroB200Error(L.SVB_EARG, "operator has been freed")
        return self._op

    def to_dense(self):
        raise TypeError("the count-level operator has no stored values; use scale_features for the explicit matrix")


def scale_features_counts(counts, scale_factor=1e4, scale_max=np.inf, features=None, libsize=None, moments="exact", levels=0):
    """Fused ``normalize_cells(X, :lognormalize; scale_factor)`` -> ``Y[:, features]`` -> ``scale_features(; scale_max)``
    (normalize.jl:40-55, docs/src/pbmc.md:121, scaling.jl:335-357) returning the centred operator over the raw counts
    (CountsCenteredMatrix) for ``irlba`` / ``embedding``. ``counts``: the full count matrix (host sparse, NamedArray or
    DeviceMatrix); library sizes are its row sums unless ``libsize`` is given. ``moments``: "exact" = the reference's
    sequential Welford on the log-normalised HVG columns (svb_mean_var; stored mu bit-identical to scale_features),
    "fast" = two parallel passes inside the operator build (a few ulp away; with a communicator they span all ranks)."""
    A, names, dimnames = _unwrap(counts)
    dA, temp = _to_device(A)
    if dA.vtype != L.SVB_I32:
        raise TypeError("scale_features_counts expects integer counts")
    lib = L.lib()
    if libsize is None:
        libsize = np.empty(dA.shape[0], dtype=np.int64)
        L.check(lib.svb_row_sums(dA._h, L.ptr(libsize)))
    sub = dA
    if features is not None:
        fidx = np.asarray(features.array if isinstance(features, NamedArray) else features, dtype=np.int64)
        sub = dA.columns(fidx)
        if names is not None:
            names = (names[0], [names[1][i] for i in fidx])
        if temp:
            dA.free()
    mean = var = None
    if moments == "exact":
        h = ctypes.c_void_p()
        L.check(lib.svb_normalize_libsize(sub._h, L.ptr(np.ascontiguousarray(libsize, dtype=np.int64)), L.NORM_LOGNORMALIZE,
                                          float(scale_factor), L.SVB_F64, ctypes.byref(h)))
        Y = DeviceMatrix(h)
        mean, var = np.empty(sub.shape[1]), np.empty(sub.shape[1])
        L.check(lib.svb_mean_var(Y._h, L.ptr(mean), L.ptr(var)))
        Y.free()
    elif moments != "fast":
        raise ValueError("moments must be 'exact' or 'fast'")
    C = CountsCenteredMatrix(sub, libsize, scale_factor, scale_max, mean, var, levels, names=names)
    C._owns_counts = sub is not counts
    return C


class _AdjointCentered:
    def __init__(self, parent):
        self.parent = parent
        self.shape = parent.shape[::-1]

    def mul(self, v, alpha=1.0, beta=0.0, y=None):
        return self.parent.mul(v, alpha, beta, y, trans=True)

    def __matmul__(self, v):
        return self.mul(v)

    @property
    def T(self):
        return self.parent


# ------------------------------------------------------------------------------------------------
# irlba.jl
# ------------------------------------------------------------------------------------------------
class SVD:
    """LinearAlgebra.SVD(U, S, V') as returned by Severo.irlba (irlba.jl:75)."""

    def __init__(self, U, S, Vt, iters=0, mprod=0):
        self.U, self.S, self.Vt = U, S, Vt
        self.iters, self.mprod = iters, mprod

    @property
    def V(self):
        return self.Vt.T

    def __iter__(self):
        return iter((self.U, self.S, self.V))


def _as_operator(A):
    if isinstance(A, _AdjointCentered):
        P = A.parent
        if P._mu is not None:
            raise TypeError("irlba of the adjoint of a centred matrix is not part of the reference path")
        return CenteredMatrix(P._parent, None, transposed=not P._transposed)
    if isinstance(A, CenteredMatrix):
        return A
    return CenteredMatrix(A, None)


def irlba(A, nu, S: Optional[SVD] = None, init=None, tol=1e-5, svtol=None, maxit=1000, rng=None, work=None):
    """Severo.irlba (irlba.jl:47-99). ``A``: CenteredMatrix, scipy sparse, DeviceMatrix or dense ndarray.
    Raises RuntimeError("convergence failed") like irlba.jl:73 when the solver reports non-zero."""
    C = _as_operator(A)
    m, n = C.shape
    nu = int(nu)
    m_b = nu + 7 if work is None else int(work)   # irlba.jl:50
    if m_b < nu:
        m_b = nu + 1
    if svtol is None:
        svtol = tol                               # irlba.jl:60 computes min(sqrt(eps), svtol) but never passes it on (T5)
    if init is None:
        rng = np.random.default_rng() if rng is None else rng
        init = rng.standard_normal(n)             # irlba.jl:62-64
    init = np.ascontiguousarray(init, dtype=np.float64)
    if init.shape[0] != n:
        raise ValueError("init must have length n")
    U = np.zeros((m, nu), order="F")
    s = np.zeros(nu)
    V = np.zeros((n, nu), order="F")
    restart = 0
    if S is not None:                             # warm restart irlba.jl:87-99
        d = len(S.S)
        U[:, :d] = S.U
        s[:d] = S.S
        V[:, :d] = S.V
        restart = d
    it, mp = ctypes.c_int64(), ctypes.c_int64()
    rc = L.lib().svb_irlba(C._operator(), nu, m_b, int(maxit), restart, float(tol), float(svtol), L.ptr(init),
                           L.ptr(s), L.ptr(U), L.ptr(V), ctypes.byref(it), ctypes.byref(mp))
    if rc in (L.SVB_ENOCONV, L.SVB_ENULLSPACE):
        raise RuntimeError("convergence failed")  # irlba.jl:73
    L.check(rc)
    return SVD(U, s, V.T, it.value, mp.value)


def mul_sparse_vector(y, A, x, alpha=1.0, beta=0.0):
    """mul.jl:50-77 ``mul!(y, A::SparseMatrixCSC, x::SparseVector, alpha, beta)``: y = beta*y + alpha*A*x in place (and returned).
    ``A``: scipy sparse / DeviceMatrix (m x n); ``x``: a scipy sparse column or row vector of length n (its STORED entries take
    part, zeros included); ``y``: Float64[m]. Raises ValueError on a size mismatch (the reference's DimensionMismatch)."""
    dA, temp = _to_device(A)
    m, n = dA.shape
    xs = sp.csc_matrix(x) if sp.issparse(x) else sp.csc_matrix(np.asarray(x, dtype=np.float64).reshape(-1, 1))
    if xs.shape[1] != 1:
        xs = sp.csc_matrix(xs.T)
    if xs.shape != (n, 1) or y.shape != (m,):
        raise ValueError("DimensionMismatch")
    if not (y.dtype == np.float64 and y.flags.c_contiguous):
        raise ValueError("y must be a contiguous Float64 vector")
    xs.sort_indices()
    idx = np.ascontiguousarray(xs.indices, dtype=np.int64)
    val = np.ascontiguousarray(xs.data, dtype=np.float64)
    L.check(L.lib().svb_spmspv(dA._h, L.ptr(idx), L.ptr(val), idx.shape[0], 0, float(alpha), float(beta), L.ptr(y)))
    if temp:
        dA.free()
    return y


def mul_sparse_dense(C, A, B, alpha=1.0, beta=0.0):
    """mul.jl:82-114 ``mul!(C::StridedMatrix, A::SparseMatrixCSC, B::SparseMatrixCSC, alpha, beta)``: C = beta*C + alpha*A*B in
    place (and returned); ``A`` may be given as ``X.T`` of a CSC (a scipy CSR) for the Transpose / Adjoint methods of
    mul.jl:79-80. ``C``: column-major Float64 (size(A,1) x size(B,2))."""
    trans = sp.issparse(A) and A.format == "csr"
    if trans:                                            # X' of a CSC: keep the parent, ask for A' * B
        A = sp.csc_matrix((A.data, A.indices, A.indptr), shape=(A.shape[1], A.shape[0]))
    dA, ta = _to_device(A)
    dB, tb = _to_device(B)
    m = dA.shape[1] if trans else dA.shape[0]
    inner = dA.shape[0] if trans else dA.shape[1]
    if inner != dB.shape[0] or C.shape != (m, dB.shape[1]):
        raise ValueError("DimensionMismatch")
    if not (C.dtype == np.float64 and C.flags.f_contiguous):
        raise ValueError("C must be a column-major Float64 matrix")
    L.check(L.lib().svb_spgemm_dense(dA._h, int(trans), dB._h, float(alpha), float(beta), L.ptr(C), max(m, 1)))
    if ta:
        dA.free()
    if tb:
        dB.free()
    return C


# ------------------------------------------------------------------------------------------------
# one process, one calling thread, N GPUs (csrc/multi.cu): the call shape of src/irlba.jl:66-71 on a whole node
# ------------------------------------------------------------------------------------------------
def init_devices(ndev=0, devices=None):
    """svb_init_devices: one worker thread per GPU inside the library (ndev <= 0: every visible GPU)."""
    d = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
    L.check(L.load().svb_init_devices(int(ndev if devices is None else len(d)), L.ptr(d)))


def shutdown_devices():
    L.check(L.load().svb_shutdown_devices())


def devices_info():
    n, pm = ctypes.c_int(), ctypes.c_int()
    dev = np.zeros(16, dtype=np.int32)
    L.check(L.load().svb_devices_info(ctypes.byref(n), L.ptr(dev), ctypes.byref(pm)))
    return dict(ndev=n.value, devices=dev[:n.value].tolist(), peer_mailboxes=bool(pm.value))


def _host_csc(A, want_float):
    A = sp.csc_matrix(A)
    if not A.has_sorted_indices:
        A = A.copy()
        A.sort_indices()
    if want_float and A.dtype not in (np.float32, np.float64):
        A = A.astype(np.float64)
    if not want_float and A.dtype not in (np.int32, np.int64):
        raise TypeError("integer counts required")
    return A


def irlba_devices(A, nu, mu=None, init=None, tol=1e-5, svtol=None, maxit=1000, rng=None, work=None):
    """``irlba(CenteredMatrix(A, mu), nu)`` (irlba.jl:47-85) on the device group of ``init_devices``: the WHOLE host matrix
    goes in one call from one thread; the library shards it by cells over its GPUs. ``A``: scipy sparse, cells x genes."""
    A = _host_csc(A, True)
    m, n = A.shape
    nu = int(nu)
    m_b = nu + 7 if work is None else int(work)
    if svtol is None:
        svtol = tol
    if init is None:
        rng = np.random.default_rng() if rng is None else rng
        init = rng.standard_normal(n)
    init = np.ascontiguousarray(init, dtype=np.float64)
    muv = None if mu is None else np.ascontiguousarray(mu, dtype=np.float64)
    colptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
    rowval = np.ascontiguousarray(A.indices)
    rt = L.SVB_I64 if rowval.dtype == np.int64 else L.SVB_I32
    if rowval.dtype not in (np.int32, np.int64):
        rowval, rt = rowval.astype(np.int64), L.SVB_I64
    nz = np.ascontiguousarray(A.data)
    U = np.zeros((m, nu), order="F")
    s = np.zeros(nu)
    V = np.zeros((n, nu), order="F")
    it, mp = ctypes.c_int64(), ctypes.c_int64()
    rc = L.load().svb_irlba_csc_devices(m, n, L.ptr(colptr), L.ptr(rowval), rt, L.ptr(nz), _DT[nz.dtype], 0, L.ptr(muv), nu, m_b,
                                        int(maxit), float(tol), float(svtol), L.ptr(init), L.ptr(s), L.ptr(U), L.ptr(V),
                                        ctypes.byref(it), ctypes.byref(mp))
    if rc in (L.SVB_ENOCONV, L.SVB_ENULLSPACE):
        raise RuntimeError("convergence failed")  # irlba.jl:73
    L.check(rc)
    return SVD(U, s, V.T, it.value, mp.value)


def pca_counts_devices(counts_hvg, libsize, nu, scale_factor=1e4, scale_max=np.inf, init=None, tol=1e-5, svtol=None, maxit=1000,
                       rng=None, work=None):
    """The fused PCA call on the device group: raw counts of the HVG columns (host scipy sparse, cells x HVGs) + the library
    sizes of the full matrix -> (SVD, stored centre mu). The scaled matrix is never materialised on any GPU."""
    A = _host_csc(counts_hvg, False)
    m, n = A.shape
    nu = int(nu)
    m_b = nu + 7 if work is None else int(work)
    if svtol is None:
        svtol = tol
    if init is None:
        rng = np.random.default_rng() if rng is None else rng
        init = rng.standard_normal(n)
    init = np.ascontiguousarray(init, dtype=np.float64)
    lib = np.ascontiguousarray(libsize, dtype=np.int64)
    if lib.shape[0] != m:
        raise ValueError("libsize must have one entry per cell")
    colptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
    rowval = np.ascontiguousarray(A.indices)
    if rowval.dtype not in (np.int32, np.int64):
        rowval = rowval.astype(np.int64)
    rt = L.SVB_I64 if rowval.dtype == np.int64 else L.SVB_I32
    nz = np.ascontiguousarray(A.data)
    U = np.zeros((m, nu), order="F")
    s = np.zeros(nu)
    V = np.zeros((n, nu), order="F")
    mu = np.zeros(n)
    it, mp = ctypes.c_int64(), ctypes.c_int64()
    rc = L.load().svb_pca_counts_devices(m, n, L.ptr(colptr), L.ptr(rowval), rt, L.ptr(nz), _DT[nz.dtype], 0, L.ptr(lib),
                                         float(scale_factor), float(scale_max), nu, m_b, int(maxit), float(tol), float(svtol),
                                         L.ptr(init), L.ptr(mu), L.ptr(s), L.ptr(U), L.ptr(V), ctypes.byref(it), ctypes.byref(mp))
    if rc in (L.SVB_ENOCONV, L.SVB_ENULLSPACE):
        raise RuntimeError("convergence failed")
    L.check(rc)
    return SVD(U, s, V.T, it.value, mp.value), mu


def gram(A):
    """``C'C`` for a CenteredMatrix (scaling.jl:274-296; plain ``A'A`` through mul.jl:82-114 when there is no centre):
    the n x n Gram matrix, computed on the device (``svb_gram``)."""
    C = _as_operator(A)
    n = C.shape[1]
    G = np.zeros((n, n), order="F")
    L.check(L.lib().svb_gram(C._operator(), L.ptr(G)))
    return G


def tssvd(A, nsv=6, ritzvec=True, tol=0.0, maxiter=1000, ncv=None, init=None, rng=None):
    """Severo.tssvd (embedding.jl:30-44): ``SVD(U, Sigma, phi')`` from the ``nsv`` largest eigenpairs of
    ``Hermitian(A'A)``. Keywords as upstream (``ncv`` defaults to ``2*nsv``, ``tol = 0.0`` = machine precision);
    ``init`` is Arpack's ``v0``. ``ritzvec=False`` returns an m x 0 ``U`` like embedding.jl:40-42."""
    C = _as_operator(A)
    m, n = C.shape
    nsv = int(nsv)
    ncv = 2 * nsv if ncv is None else int(ncv)
    ncv = max(min(ncv, n), min(nsv + 1, n))
    if init is None:
        rng = np.random.default_rng() if rng is None else rng
        init = rng.standard_normal(n)
    init = np.ascontiguousarray(init, dtype=np.float64)
    if init.shape[0] != n:
        raise ValueError("init must have length n")
    h = ctypes.c_void_p()
    rc = L.lib().svb_tssvd(C._operator(), nsv, ncv, int(maxiter), float(tol), L.ptr(init), ctypes.byref(h))
    L.check(rc)
    try:
        it, mp, info = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
        L.check(L.lib().svb_result_info(h, None, None, None, ctypes.byref(it), ctypes.byref(mp), ctypes.byref(info)))
        if info.value != 0:
            raise RuntimeError("convergence failed")
        s = np.zeros(nsv)
        V = np.zeros((n, nsv), order="F")
        U = np.zeros((m, nsv if ritzvec else 0), order="F")
        L.check(L.lib().svb_result_download(h, L.ptr(s), L.ptr(U) if ritzvec else None, L.ptr(V), 0))
    finally:
        L.lib().svb_result_free(h)
    return SVD(U, s, V.T, it.value, mp.value)


def svd_flip(S: SVD, u_based_decision=True):
    """utils.jl:215-228 svd_flip!: make the max-|.| entry of every U (or V) column positive (in place)."""
    M = S.U if u_based_decision else S.V
    idx = np.argmax(np.abs(M), axis=0)
    signs = np.sign(M[idx, np.arange(M.shape[1])])
    S.U *= signs[None, :]
    S.Vt *= signs[:, None]
    return S


# ------------------------------------------------------------------------------------------------
# embedding.jl
# ------------------------------------------------------------------------------------------------
@dataclass
class LinearEmbedding:
    """embedding.jl:20-25."""
    parent: object
    coordinates: object
    stdev: object
    basis: object


def _pca(X, npcs, algorithm="arpack", **kw):
    """embedding.jl:46-76. ``algorithm``:

    * ``:irlba``  — the path of this build (irlba.jl -> svb_irlba).
    * ``:arpack`` — the reference's DEFAULT (embedding.jl:46, ``svds`` of the external Arpack.jl). Converged singular
      triplets are unique up to sign, so the request is served by the same device solver run to Arpack-like accuracy
      (``tol`` default 1e-10; Arpack's own default is machine eps); ``maxiter`` / ``ncv`` map to ``maxit`` / ``work``.
      Iterates differ from Arpack's, results agree to the tolerance.
    * ``:tssvd`` — embedding.jl:30-44: the Gram matrix ``C'C`` (scaling.jl:274-296) is built on the device, its
      eigenpairs come from the device solver (``svb_tssvd``); keywords ``tol``, ``maxiter``, ``ncv`` as upstream.
    * dense ``svd`` — LAPACK on the host: outside this path.
    """
    algorithm = str(algorithm)
    if algorithm == "arpack":
        kw = dict(kw)
        kw.setdefault("tol", 1e-10)
        if "maxiter" in kw:
            kw["maxit"] = kw.pop("maxiter")
        if "ncv" in kw:
            kw["work"] = kw.pop("ncv")
        kw.pop("nsv", None)
    elif algorithm not in ("irlba", "tssvd"):
        raise ValueError(f"algorithm {algorithm} is outside the B200 hot path (use algorithm=:irlba, :arpack or :tssvd)")
    C = _as_operator(X)
    m, n = C.shape
    npcs = min(min(m, n), int(npcs))
    if npcs > 0.5 * min(m, n) and algorithm != "tssvd":
        # the reference switches to a dense LAPACK svd here (embedding.jl:50-53); there is no CPU fallback
        # in this build, IRLBA with work = min(m, n) spans the whole space and is exact in that regime.
        warnings.warn("Computing too large a percentage of principal components")
    if algorithm == "tssvd":
        kw = dict(kw)
        kw.pop("nsv", None)
        S = tssvd(C, nsv=npcs, **kw)                       # embedding.jl:58-59
    else:
        S = irlba(C, npcs, **kw)
    Z = S.U[:, :npcs] * S.S[None, :npcs]                     # embedding.jl:67
    stdev = S.S[:npcs] / np.sqrt(max(1, m - 1))             # embedding.jl:68
    loadings = S.V[:, :npcs]
    return Z, stdev, loadings


def pca(X, npcs, **kw):
    """embedding.jl:81-94."""
    Z, stdev, loadings = _pca(X, npcs, **kw)
    k = stdev.shape[0]
    latent = [f"PC-{i + 1}" for i in range(k)]
    names = X.names if isinstance(X, CenteredMatrix) else (X.names if isinstance(X, NamedArray) else None)
    if names is None:
        return LinearEmbedding(X, Z, stdev, loadings)
    rowdim = (getattr(X.A, "dimnames", ("cells", "features")) if isinstance(X, CenteredMatrix) else X.dimnames)[0]
    coordinates = NamedArray(Z, (names[0], latent), (rowdim, "latent"))
    stdevn = NamedArray(stdev, (latent,), ("latent",))
    basis = NamedArray(loadings, (names[1], latent), (rowdim, "latent"))  # sic: (rowdim, :latent) embedding.jl:92
    return LinearEmbedding(X, coordinates, stdevn, basis)


def embedding(X, ncomponents=50, method="pca", **kw):
    """embedding.jl:202-212."""
    method = str(method)
    if method == "pca":
        return pca(X, int(ncomponents), **kw)
    raise ValueError(f"unknown reduction method: {method}")


# ------------------------------------------------------------------------------------------------
# neighbours.jl (the step after the path)
# ------------------------------------------------------------------------------------------------
_METRICS = {"euclidean": L.METRIC_EUCLIDEAN, "euclidian": L.METRIC_EUCLIDEAN, "cosine": L.METRIC_COSINE}


def ann(X, k, metric="euclidean", include_self=True, ntables=None, rng=None):
    """Severo.ann (neighbours.jl:19-31): ``nn_index`` (n x k Int32) and ``distances`` (n x k, element type of ``X``) of the
    k nearest neighbours of every row of ``X``. The search is exact on the device (``svb_knn``); ``ntables`` / ``rng``, the
    knobs of the reference's randomised approximation, are accepted and ignored. Indices are 0-based here (the Julia
    overlay asks for 1-based)."""
    X = X.array if isinstance(X, NamedArray) else X
    X = np.asarray(X)
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)
    X = np.asfortranarray(X)
    n, d = X.shape
    nn_index = np.zeros((n, int(k)), dtype=np.int32, order="F")
    distances = np.zeros((n, int(k)), dtype=X.dtype, order="F")
    L.check(L.lib().svb_knn(L.ptr(X), _DT[X.dtype], n, d, X.strides[1] // X.itemsize if d > 1 else n, int(k),
                            _METRICS[str(metric).lower()], int(bool(include_self)), 0, L.ptr(nn_index), L.ptr(distances)))
    return nn_index, distances


def nearest_neighbours(X, k, dims=None, metric="euclidean", include_self=True, ntables=None, rng=None):
    """Severo.nearest_neighbours (neighbours.jl:76-86,176-180,224): the k-nearest-neighbour graph as an n x n boolean
    sparse matrix, entry (j, i) set when cell j is among the k nearest neighbours of cell i. ``X``: coordinates (n x d),
    a NamedArray of them, or a LinearEmbedding (its ``coordinates``); ``dims``: which coordinates to use (``:`` = all)."""
    names = None
    if isinstance(X, LinearEmbedding):
        X = X.coordinates
    rowdim = "cells"
    if isinstance(X, NamedArray):
        names, rowdim, X = X.names[0], X.dimnames[0], X.array
    X = np.asarray(X)
    if dims is not None:
        X = X[:, dims]
    idx, _ = ann(X, k, metric=metric, include_self=include_self)
    n = X.shape[0]
    nn = sp.csc_matrix((np.ones(n * int(k), dtype=bool), idx.ravel(order="C"), np.arange(0, n * int(k) + 1, int(k))), shape=(n, n))
    nn.sort_indices()
    if names is None:
        return nn
    return NamedArray(nn, (names, names), (rowdim, rowdim))           # neighbours.jl:179


def _jaccard(nn, k, prune, dtype):
    nn = sp.csc_matrix(nn)
    if nn.shape[0] != nn.shape[1]:
        raise ValueError("jaccard_index: the neighbour graph must be square (cells x cells)")
    if nn.dtype != np.bool_ or nn.nnz != int(np.count_nonzero(nn.data)):
        nn = nn.copy()
        nn.eliminate_zeros()                          # only `true` entries are neighbours
    if not nn.has_sorted_indices:
        nn = nn.copy()
        nn.sort_indices()
    dt = np.dtype(dtype)
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError("jaccard_index: dtype must be Float32 or Float64")
    pattern = sp.csc_matrix((np.ones(nn.nnz, dtype=np.int32), nn.indices, nn.indptr), shape=nn.shape)
    dN = DeviceMatrix.from_host(pattern)
    h = ctypes.c_void_p()
    try:
        L.check(L.lib().svb_jaccard_index(dN._h, 0 if k is None else int(k), float(prune), _DT[dt], ctypes.byref(h)))
    finally:
        dN.free()
    out = DeviceMatrix(h)
    snn = out.to_host(dt)
    out.free()
    return snn


def jaccard_index(nn, k=None, prune=1.0 / 15.0, dtype=np.float64):
    """Severo.jaccard_index (neighbours.jl:112-131,133-152 over ``_jaccard_index`` :88-110): the shared-nearest-neighbour
    graph of a neighbour graph ``nn`` (n x n boolean sparse, column i = the neighbours of cell i, as ``nearest_neighbours``
    returns it). Entry (i, j) = |N(i) ∩ N(j)| / |N(i) ∪ N(j)| computed as ``x / (k + (k - x))`` with the given ``k`` or,
    without it, the size of N(j); entries with a value <= ``prune`` are removed. Computed on the device (``svb_jaccard_index``)
    without forming ``nn' * nn``. Returns CSC of ``dtype`` (labelled like the input when that is a NamedArray)."""
    A, names, dimnames = _unwrap(nn)
    if k is not None and int(k) <= 0:
        raise ValueError("jaccard_index: k must be positive")
    snn = _jaccard(A, k, prune, dtype)
    if names is None:
        return snn
    return NamedArray(snn, (names[0], names[0]), (dimnames[0], dimnames[0]))      # neighbours.jl:130


def shared_nearest_neighbours(X, k, dims=None, metric="euclidean", include_self=True, ntables=None, prune=1.0 / 15.0, rng=None):
    """Severo.shared_nearest_neighbours (neighbours.jl:263-270,289,308,327): ``nearest_neighbours`` followed by the Jaccard
    index with the fixed ``k``, in the element type of the coordinates (``prune = convert(T, prune)``, :267)."""
    nn = nearest_neighbours(X, k, dims=dims, metric=metric, include_self=include_self)
    if isinstance(X, LinearEmbedding):
        X = X.coordinates
    arr = X.array if isinstance(X, NamedArray) else np.asarray(X)
    dt = arr.dtype if arr.dtype in (np.float32, np.float64) else np.dtype(np.float64)
    return jaccard_index(nn, int(k), prune, dt)


# ------------------------------------------------------------------------------------------------
# synthetic inputs (benchmark / tests)
# ------------------------------------------------------------------------------------------------
def synthetic_counts(m_total, genes, mean_nnz_per_cell, programs=64, fold=6.0, seed=20260101, rows=None):
    """Device-generated Poisson count matrix (cells x genes, int32), see csrc/synth.cu."""
    row0, row1 = (0, m_total) if rows is None else rows
    h = ctypes.c_void_p()
    L.check(L.lib().svb_synth_counts(int(m_total), int(genes), int(row0), int(row1), float(mean_nnz_per_cell),
                                      int(programs), float(fold), int(seed), ctypes.byref(h)))
    return DeviceMatrix(h)
