"""CPU ORACLE for the Severo.jl IRLBA-PCA hot path — TEST INFRASTRUCTURE, NOT THE PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module. The product package (``severo.jl_b200``)
never imports, links or executes anything under ``oracle/``.

It restates, on the CPU and in Float64, the algorithms of the reference's path (paths relative
to the reference repo root):

* ``src/normalize.jl:17-55``      row_norm / log_norm / normalize_cells
* ``src/scaling.jl:18-44,119-147`` mean_var / mean_std (sequential Welford, implicit zeros pre-counted)
* ``src/variablefeatures.jl:19-28`` standardized_var_clipped
* ``src/scaling.jl:199-217``      scale_data
* ``src/scaling.jl:219-272,298-314`` CenteredMatrix and its mul! family / convert
* ``src/irlba.jl:47-99``          irlba!/irlba wrapper defaults (work = nu+7, tol = 1e-5, maxit = 1000)
* ``src/embedding.jl:46-76``      _pca post-processing
* ``src/utils.jl:215-228``        svd_flip!

Order-exact loops (Welford, ``sf*x/s``, ``x/std``) run in ``csrc/severo_oracle.c`` compiled with
``-ffp-contract=off`` (Julia never contracts to FMA); small inputs can also be run through the
pure-Python twins below (``*_py``), which the CPU tests use to validate the C build.

PARITY STATUS
-------------
* Pre-processing + operator: pinned against the reference's own fixed-input tests
  (``test/test_scaling.jl:22-45,72-112``, ``test/test_input.jl:47-75``) — see tests/golden/.
* IRLBA: the arithmetic lives in the external, un-vendored and un-pinned binary ``libcell``
  (package ``Severo_jll``, no version in Project.toml; call site ``src/irlba.jl:66-71``). Neither
  Julia nor libcell exists in this image, so the *iterates* are **parity unpinned**. The loop
  below restates the published algorithm libcell derives from (Baglama & Reichel's implicitly
  restarted Lanczos bidiagonalisation as implemented in B. W. Lewis' ``irlb.c`` of the R package
  *irlba*: same signature, same defaults). What IS pinned is the converged result, by the
  reference's own criteria in ``test/test_irlba.jl`` (singular values vs dense LAPACK ``svd`` at
  rtol sqrt(eps); ``||X'U - V S|| / ||X|| < tol``; rank-k reconstruction error).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRCS = [os.path.join(_HERE, "csrc", "severo_oracle.c"), os.path.join(_HERE, "csrc", "synth_twin.c")]
_BUILD = os.path.join(_HERE, "_build")
_LIB = os.path.join(_BUILD, "libsevero_oracle.so")
_STAMP = os.path.join(_BUILD, "cpu_flags.txt")

_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)
_f32p = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def _cpu_flags() -> str:
    """The instruction-set flags of this host: the library is built -march=native, so a copy built on another machine
    (the container builds, the GPU box runs) is rebuilt when the flags differ."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " ".join(sorted(line.split(":", 1)[1].split()))
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    """Compile csrc/*.c -> _build/libsevero_oracle.so: gcc -O3 -march=native, no FMA contraction, no fast-math
    (BASELINE.md section 4's flags for the CPU baseline)."""
    flags = _cpu_flags()
    fresh = os.path.exists(_LIB) and all(os.path.getmtime(_LIB) >= os.path.getmtime(src) for src in _SRCS)
    if fresh and os.path.exists(_STAMP):
        with open(_STAMP) as f:
            fresh = f.read() == flags
    else:
        fresh = False
    if fresh and not force:
        return _LIB
    os.makedirs(_BUILD, exist_ok=True)
    cmd = ["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-fvisibility=hidden", "-o", _LIB] + _SRCS + ["-lm"]
    subprocess.run(cmd, check=True)
    with open(_STAMP, "w") as f:
        f.write(flags)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def _p(a, t):
    return a.ctypes.data_as(t)


def _csc(A):
    """Canonical CSC with int64 indices, sorted rows (what Julia's SparseMatrixCSC guarantees)."""
    A = sp.csc_matrix(A)
    if not A.has_sorted_indices:
        A = A.copy()
        A.sort_indices()
    return A


def _idx64(A):
    return (np.ascontiguousarray(A.indptr, dtype=np.int64),
            np.ascontiguousarray(A.indices, dtype=np.int64))


# --------------------------------------------------------------------------------------
# normalize.jl
# --------------------------------------------------------------------------------------
def filter_features(A, min_cells=0):
    """filtering.jl:15-20: cells_per_feature = vec(sum(>(0), A, dims=1)); FI = cells_per_feature .>= min_cells."""
    A = _csc(A)
    cells_per_feature = np.asarray((A > 0).sum(axis=0)).ravel()
    FI = cells_per_feature >= min_cells
    return A[:, FI], FI


def filter_cells(A, min_features=0, min_feature_count=0, min_umi=0):
    """filtering.jl:22-35: features_per_cell = vec(sum(>(min_feature_count), A, dims=2)); CI = .>= min_features;
    if min_umi > 0: CI .&= vec(sum(A, dims=2)) .> min_umi."""
    A = _csc(A)
    features_per_cell = np.asarray((A > min_feature_count).sum(axis=1)).ravel()
    CI = features_per_cell >= min_features
    if min_umi > 0:
        CI &= np.asarray(A.sum(axis=1)).ravel() > min_umi
    return sp.csc_matrix(A[CI, :]), CI


def filter_counts(A, min_cells=0, min_features=0, min_feature_count=0, min_umi=0):
    """filtering.jl:101-106: cells first, then features on what remains."""
    counts, CI = filter_cells(A, min_features, min_feature_count, min_umi)
    counts, FI = filter_features(counts, min_cells)
    return sp.csc_matrix(counts), CI, FI


def row_norm(A, scale_factor=1.0, dtype=np.float64):
    """normalize.jl:17-32. Integer counts in, ``dtype`` values out, same sparsity pattern."""
    A = _csc(A)
    colptr, rowval = _idx64(A)
    if not np.issubdtype(A.dtype, np.integer):
        raise TypeError("oracle row_norm restates the integer-count path (normalize.jl:24 exact Int sum)")
    nz = np.ascontiguousarray(A.data, dtype=np.int64)
    m, n = A.shape
    s = np.zeros(m, dtype=np.int64)
    L = lib()
    L.orc_row_sums_i64(ctypes.c_int64(n), _p(colptr, _i64p), _p(rowval, _i64p), _p(nz, _i64p),
                       ctypes.c_int64(m), _p(s, _i64p))
    return _row_norm_apply(A, rowval, nz, s, scale_factor, dtype, do_log=0)


def _row_norm_apply(A, rowval, nz, s, scale_factor, dtype, do_log):
    L = lib()
    out = np.empty(nz.shape[0], dtype=dtype)
    if np.dtype(dtype) == np.float64:
        L.orc_row_norm_f64(ctypes.c_int64(nz.shape[0]), _p(rowval, _i64p), _p(nz, _i64p), _p(s, _i64p),
                           ctypes.c_double(float(scale_factor)), ctypes.c_int(do_log), _p(out, _f64p))
    elif np.dtype(dtype) == np.float32:
        L.orc_row_norm_f32(ctypes.c_int64(nz.shape[0]), _p(rowval, _i64p), _p(nz, _i64p), _p(s, _i64p),
                           ctypes.c_float(float(np.float32(scale_factor))), ctypes.c_int(do_log),
                           _p(out, _f32p))
    else:
        raise TypeError("dtype must be float32 or float64")
    return sp.csc_matrix((out, A.indices.copy(), A.indptr.copy()), shape=A.shape)


def log_norm(A, scale_factor=1.0, dtype=np.float64):
    """normalize.jl:34-38 (row_norm then log1p on the stored values)."""
    A = _csc(A)
    colptr, rowval = _idx64(A)
    nz = np.ascontiguousarray(A.data, dtype=np.int64)
    m, n = A.shape
    s = np.zeros(m, dtype=np.int64)
    lib().orc_row_sums_i64(ctypes.c_int64(n), _p(colptr, _i64p), _p(rowval, _i64p), _p(nz, _i64p),
                           ctypes.c_int64(m), _p(s, _i64p))
    return _row_norm_apply(A, rowval, nz, s, scale_factor, dtype, do_log=1)


def normalize_cells(X, method="lognormalize", scale_factor=1.0, dtype=np.float64):
    """normalize.jl:40-55 — method Symbol-or-String dispatch, scale_factor converted to dtype."""
    method = str(method)
    if method == "lognormalize":
        f = log_norm
    elif method == "relativecounts":
        f = row_norm
    else:
        raise ValueError(f"unknown normalization method: {method}")
    return f(X, np.dtype(dtype).type(scale_factor), dtype)


def row_norm_py(A, scale_factor=1.0, do_log=False):
    """Pure-Python twin of row_norm/log_norm (tiny inputs only) — validates the C build."""
    import math
    A = _csc(A)
    m, n = A.shape
    s = [0] * m
    for c in range(n):
        for j in range(A.indptr[c], A.indptr[c + 1]):
            s[A.indices[j]] += int(A.data[j])
    out = np.empty(A.nnz, dtype=np.float64)
    for j in range(A.nnz):
        v = (float(scale_factor) * float(A.data[j])) / float(s[A.indices[j]])
        out[j] = math.log1p(v) if do_log else v
    return sp.csc_matrix((out, A.indices.copy(), A.indptr.copy()), shape=A.shape)


# --------------------------------------------------------------------------------------
# scaling.jl — moments
# --------------------------------------------------------------------------------------
def mean_var_py(values, n, dtype=np.float64):
    """scaling.jl:18-34 literally, in Python floats (Float64) or numpy float32 scalars."""
    T = np.dtype(dtype).type
    count = n - len(values)
    mu = T(0)
    s = T(0)
    for v in values:
        v = T(v)
        count += 1
        delta = v - mu
        mu = T(mu + T(delta / T(count)))
        s = T(s + T(delta * T(v - mu)))
    var = T(s / T(n - 1))
    return mu, var


def mean_var(A, dtype=None):
    """scaling.jl:132-147 mean_var(A::SparseMatrixCSC): integer data -> Float64; float data -> own type."""
    A = _csc(A)
    colptr, _ = _idx64(A)
    m, n = A.shape
    L = lib()
    if np.issubdtype(A.dtype, np.integer):
        if dtype is not None and np.dtype(dtype) != np.float64:
            # mean_var(Float32, A) on integer data: run the generic twin
            return _mean_var_generic(A, dtype)
        nz = np.ascontiguousarray(A.data, dtype=np.int64)
        mu = np.empty(n)
        var = np.empty(n)
        L.orc_mean_var_csc_i64(ctypes.c_int64(m), ctypes.c_int64(n), _p(colptr, _i64p), _p(nz, _i64p),
                               _p(mu, _f64p), _p(var, _f64p))
        return mu, var
    if A.dtype == np.float32 and (dtype is None or np.dtype(dtype) == np.float32):
        nz = np.ascontiguousarray(A.data, dtype=np.float32)
        mu = np.empty(n, dtype=np.float32)
        var = np.empty(n, dtype=np.float32)
        L.orc_mean_var_csc_f32(ctypes.c_int64(m), ctypes.c_int64(n), _p(colptr, _i64p), _p(nz, _f32p),
                               _p(mu, _f32p), _p(var, _f32p))
        return mu, var
    nz = np.ascontiguousarray(A.data, dtype=np.float64)
    mu = np.empty(n)
    var = np.empty(n)
    L.orc_mean_var_csc_f64(ctypes.c_int64(m), ctypes.c_int64(n), _p(colptr, _i64p), _p(nz, _f64p),
                           _p(mu, _f64p), _p(var, _f64p))
    return mu, var


def _mean_var_generic(A, dtype):
    m, n = A.shape
    mu = np.empty(n, dtype=dtype)
    var = np.empty(n, dtype=dtype)
    for c in range(n):
        mu[c], var[c] = mean_var_py(A.data[A.indptr[c]:A.indptr[c + 1]], m, dtype)
    return mu, var


def mean_std(A, dtype=None):
    """scaling.jl:119-130."""
    mu, var = mean_var(A, dtype)
    return mu, np.sqrt(var)


# --------------------------------------------------------------------------------------
# variablefeatures.jl — the data sweep after the (host, unpinned) loess fit
# --------------------------------------------------------------------------------------
def standardized_var_clipped(A, mu, sd, vmax=None):
    """variablefeatures.jl:21-28. Returns the per-gene clipped standardised variance."""
    A = _csc(A)
    if not np.issubdtype(A.dtype, np.integer):
        raise TypeError("standardized_var_clipped is defined on integer counts (variablefeatures.jl:21)")
    colptr, _ = _idx64(A)
    nz = np.ascontiguousarray(A.data, dtype=np.int64)
    m, n = A.shape
    if vmax is None:
        vmax = np.sqrt(float(m))
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    sd = np.ascontiguousarray(sd, dtype=np.float64)
    out = np.zeros(n)
    lib().orc_stdvar_clipped_i64(ctypes.c_int64(m), ctypes.c_int64(n), _p(colptr, _i64p), _p(nz, _i64p),
                                 _p(mu, _f64p), _p(sd, _f64p), ctypes.c_double(float(vmax)), _p(out, _f64p))
    return out


# --------------------------------------------------------------------------------------
# variablefeatures.jl:52-103 — the selectors next to :vst (:saunders, :dispersion, :meanvarplot). Data sweeps = row sums,
# row_norm and mean_var (above); everything else is arithmetic on gene-length vectors, restated literally (loops).
# --------------------------------------------------------------------------------------
def log_VMR(norm):
    """scaling.jl:190-197: (log1p.(mu), log.(var ./ mu)) of the per-gene moments."""
    mu, var = mean_var(norm)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log1p(mu), np.log(var / mu)


def _nan2zero(v):
    """variablefeatures.jl:30-32."""
    return np.where(np.isnan(v), 0.0, v)


def cut_width(v, nbreaks):
    """utils.jl:140-148,155 with utils.jl:110-128 (`_cut!`, right = true): equal-width breaks over [min, max], the two outer
    breaks moved out by dx/1000; label = searchsortedlast(breaks, x), minus one when x sits exactly on that break; 0 outside."""
    v = np.asarray(v, dtype=np.float64)
    lo, hi = float(v.min()), float(v.max())
    dx = hi - lo
    from fractions import Fraction
    # range(min, max, length = n+1): Julia evaluates it in twice-precision arithmetic, i.e. (almost always) the correctly
    # rounded min + i*(max-min)/n — restated with exact rationals
    breaks = [float(Fraction(lo) + (Fraction(hi) - Fraction(lo)) * i / nbreaks) for i in range(nbreaks + 1)]
    breaks[0] -= dx / 1000
    breaks[-1] += dx / 1000
    labels = np.zeros(v.shape[0], dtype=np.int64)
    for i, x in enumerate(v):
        if breaks[0] <= x <= breaks[-1]:
            idx = 0                                   # searchsortedlast: number of breaks <= x (1-based index of the last one)
            for b in breaks:
                if b <= x:
                    idx += 1
            if x == breaks[idx - 1]:
                idx -= 1
            labels[i] = idx
    return np.array(breaks), labels


def mean_std_labels(x, lbls, nlabels):
    """scaling.jl:89-112: Welford per label over a dense vector, unbiased, std = sqrt(var). Labels are 1-based."""
    mu = np.zeros(nlabels)
    var = np.zeros(nlabels)
    n = np.zeros(nlabels, dtype=np.int64)
    for v, k in zip(x, lbls):
        k = int(k) - 1
        if k < 0:
            raise IndexError("label 0: the reference indexes out of bounds here (utils.jl:122, scaling.jl:96)")
        n[k] += 1
        delta = v - mu[k]
        mu[k] += delta / n[k]
        var[k] += delta * (v - mu[k])
    with np.errstate(divide="ignore", invalid="ignore"):
        var = var / (n - 1)
        return mu, np.sqrt(var)


def select_dispersion(norm):
    """variablefeatures.jl:73-76."""
    _, disp = log_VMR(norm)
    return _nan2zero(disp)


def select_meanvarplot(norm, num_bins=20):
    """variablefeatures.jl:78-92 (binning_method = :width)."""
    mu, disp = log_VMR(norm)
    mu = _nan2zero(mu)
    disp = _nan2zero(disp)
    _, bins = cut_width(mu, num_bins)
    bin_mean, bin_std = mean_std_labels(disp, bins, num_bins)
    out = np.empty_like(disp)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(disp.shape[0]):
            k = bins[i] - 1
            out[i] = (disp[i] - bin_mean[k]) / bin_std[k]
    return _nan2zero(out)


def select_features_saunders(counts, norm, alpha_thresh=0.99):
    """variablefeatures.jl:54-71."""
    from scipy.stats import norm as _normal
    counts = _csc(counts)
    ncells, ngenes = counts.shape
    trx_per_cell = np.asarray(counts.sum(axis=1)).ravel().astype(np.float64)
    mu, var = mean_var(norm)
    nolan = float(np.mean(1.0 / trx_per_cell))
    corrected = float(alpha_thresh) / ngenes
    z = float(_normal.ppf(1.0 - corrected / 2.0))
    metric = np.zeros(ngenes)
    for j in range(ngenes):
        upper = mu[j] + z * np.sqrt(mu[j] * nolan / ncells)
        if var[j] / nolan > upper:
            metric[j] = np.log10(var[j]) - np.log10(mu[j] * nolan)
    return metric


def variable_feature_metric(counts, method, norm=None, **kw):
    """variablefeatures.jl:128-157 for the three selectors above: ``norm`` defaults to row_norm(counts, 1) (:136-141)."""
    if norm is None:
        norm = row_norm(counts, 1.0)
    if method == "saunders":
        return select_features_saunders(counts, norm, kw.get("alpha_thresh", 0.1))          # default of the call site, :144
    if method == "dispersion":
        return select_dispersion(norm)
    if method == "meanvarplot":
        return select_meanvarplot(norm, kw.get("num_bins", 20))
    raise ValueError(f"unknown selection method: {method}")


# --------------------------------------------------------------------------------------
# scaling.jl — scale_data / CenteredMatrix
# --------------------------------------------------------------------------------------
def scale_data(A, scale_max=np.inf):
    """scaling.jl:199-217. Returns (B, mu) with mu = mean/std (trap T3)."""
    A = _csc(A)
    colptr, _ = _idx64(A)
    m, n = A.shape
    out = np.empty(A.nnz)
    mu = np.empty(n)
    L = lib()
    if np.issubdtype(A.dtype, np.integer):
        nz = np.ascontiguousarray(A.data, dtype=np.int64)
        L.orc_scale_data_i64(ctypes.c_int64(m), ctypes.c_int64(n), _p(colptr, _i64p), _p(nz, _i64p),
                             ctypes.c_double(float(scale_max)), _p(out, _f64p), _p(mu, _f64p))
    else:
        nz = np.ascontiguousarray(A.data, dtype=np.float64)
        L.orc_scale_data_f64(ctypes.c_int64(m), ctypes.c_int64(n), _p(colptr, _i64p), _p(nz, _f64p),
                             ctypes.c_double(float(scale_max)), _p(out, _f64p), _p(mu, _f64p))
    return sp.csc_matrix((out, A.indices.copy(), A.indptr.copy()), shape=A.shape), mu


def scale_features(X, scale_max=np.inf, features=None):
    """scaling.jl:335-357 (without the NamedArray labels). Returns a CenteredMatrix."""
    X = _csc(X)
    if features is not None:
        X = _csc(X[:, np.asarray(features)])
    B, mu = scale_data(X, scale_max)
    return CenteredMatrix(B, mu)


class CenteredMatrix:
    """scaling.jl:219-272: the implicit operator S = A - 1*mu' ; A is CSC, a lazy adjoint of a CSC
    (``transposed=True`` keeps the parent, as ``CenteredMatrix(X', mu)`` in test_irlba.jl:111), or dense."""

    def __init__(self, A, mu, transposed=False):
        self.dense = not sp.issparse(A)
        self.transposed = bool(transposed)
        if self.dense:
            P = np.asarray(A, dtype=np.float64)
            self.P = P.T if transposed else P
            self.transposed = False
        else:
            P = _csc(A).astype(np.float64)
            self.P = P
            self.colptr, self.rowval = _idx64(P)
            self.nz = np.ascontiguousarray(P.data, dtype=np.float64)
            self._csr = None
        pm, pn = self.P.shape
        self.shape = (pn, pm) if self.transposed else (pm, pn)
        self.mu = None if mu is None else np.ascontiguousarray(mu, dtype=np.float64)
        if self.mu is not None:
            assert self.mu.shape[0] == self.shape[1]  # scaling.jl:226

    # -- raw products with the *parent* CSC P (pm x pn) --------------------------------
    def _P_mul(self, v, alpha, beta, y, parallel=False):
        pm, pn = self.P.shape
        if parallel:
            if self._csr is None:
                rowptr = np.empty(pm + 1, dtype=np.int64)
                colidx = np.empty(self.nz.shape[0], dtype=np.int32)
                val = np.empty(self.nz.shape[0])
                lib().orc_csc_to_csr(ctypes.c_int64(pm), ctypes.c_int64(pn), _p(self.colptr, _i64p),
                                     _p(self.rowval, _i64p), _p(self.nz, _f64p), _p(rowptr, _i64p),
                                     _p(colidx, _i32p), _p(val, _f64p))
                self._csr = (rowptr, colidx, val)
            rowptr, colidx, val = self._csr
            lib().orc_csr_mul(ctypes.c_int64(pm), _p(rowptr, _i64p), _p(colidx, _i32p), _p(val, _f64p),
                              _p(v, _f64p), ctypes.c_double(alpha), ctypes.c_double(beta), _p(y, _f64p))
        else:
            lib().orc_csc_mul(ctypes.c_int64(pm), ctypes.c_int64(pn), _p(self.colptr, _i64p),
                              _p(self.rowval, _i64p), _p(self.nz, _f64p), _p(v, _f64p),
                              ctypes.c_double(alpha), ctypes.c_double(beta), _p(y, _f64p))

    def _Pt_mul(self, v, alpha, beta, y):
        pm, pn = self.P.shape
        lib().orc_csc_mul_t(ctypes.c_int64(pm), ctypes.c_int64(pn), _p(self.colptr, _i64p),
                            _p(self.rowval, _i64p), _p(self.nz, _f64p), _p(v, _f64p),
                            ctypes.c_double(alpha), ctypes.c_double(beta), _p(y, _f64p))

    # -- mul!(C, S, v, a, b)  scaling.jl:245-250 ;  mul!(C, S', v, a, b)  scaling.jl:252-257 --
    def mul(self, v, alpha=1.0, beta=0.0, y=None, trans=False, parallel=False):
        v = np.ascontiguousarray(v, dtype=np.float64)
        m, n = self.shape
        if v.ndim == 2:
            return self._mul_mat(v, alpha, beta, y, trans, parallel)
        out_len = n if trans else m
        assert v.shape[0] == (m if trans else n)
        if y is None:
            y = np.zeros(out_len)
            beta = 0.0
        if self.dense:
            Av = (self.P.T @ v) if trans else (self.P @ v)
            y[:] = alpha * Av + (beta * y if beta != 0.0 else 0.0)
        else:
            use_parent_adjoint = (trans != self.transposed)
            if use_parent_adjoint:
                self._Pt_mul(v, alpha, beta, y)
            else:
                self._P_mul(v, alpha, beta, y, parallel)
        if self.mu is not None:
            if trans:
                y += (-alpha * np.sum(v)) * self.mu        # axpy!(-a*sum(v), mu, C)  scaling.jl:256
            else:
                y -= alpha * np.dot(self.mu, v)             # C .-= a*dot(mu, v)       scaling.jl:248
        return y

    def _mul_mat(self, V, alpha, beta, Y, trans, parallel):
        """scaling.jl:259-272 matrix forms, column by column. The adjoint form SUBTRACTS the
        rank-1 term (mathematically correct); the reference's :271 adds it — untested upstream
        sign slip, SURVEY trap T4."""
        k = V.shape[1]
        m, n = self.shape
        out_len = n if trans else m
        if Y is None:
            Y = np.zeros((out_len, k), order="F")
            beta = 0.0
        for c in range(k):
            yc = np.ascontiguousarray(Y[:, c])
            Y[:, c] = self.mul(np.ascontiguousarray(V[:, c]), alpha, beta, yc, trans, parallel)
        return Y

    def to_dense(self):
        """convert(Matrix, C) scaling.jl:298-303."""
        X = self.P if self.dense else self.P.toarray()
        X = np.array(X.T if self.transposed else X, dtype=np.float64)
        if self.mu is not None:
            X = X - self.mu[None, :]
        return X


# --------------------------------------------------------------------------------------
# IRLBA (libcell.irlba restated; see module docstring for parity status)
# --------------------------------------------------------------------------------------
@dataclass
class IrlbaResult:
    U: np.ndarray
    S: np.ndarray
    V: np.ndarray
    iters: int
    mprod: int
    info: int


_EPS23 = np.finfo(np.float64).eps ** (2.0 / 3.0)


def _orthog(X, y, j):
    """Classical Gram-Schmidt, one pass: y -= X[:, :j] (X[:, :j]' y)  (irlb.c `orthog`)."""
    if j > 0:
        t = X[:, :j].T @ y
        y -= X[:, :j] @ t
    return y


def irlba(A, nu, init=None, tol=1e-5, svtol=None, maxit=1000, work=None, rng=None, parallel=False,
          restart_from=None):
    """Restated ``libcell.irlba`` as wrapped by src/irlba.jl:47-85.

    ``A`` needs ``.shape`` and ``.mul(v, trans=...)`` (CenteredMatrix) or is a numpy / scipy matrix.
    Defaults follow irlba.jl: work = min(nu+7, min(m,n)) (:50-58), tol = 1e-5, maxit = 1000,
    init = randn(n) (:62-64). ``svtol`` defaults to ``tol`` (the wrapper computes a svtol at :60 but
    never passes it to C, trap T5).
    Returns IrlbaResult with info = 0 (converged) or -2 (maxit) / -4 (init in null space).
    """
    if not isinstance(A, CenteredMatrix):
        A = CenteredMatrix(A, None)
    m, n = A.shape
    nu = int(nu)
    if work is None:
        work = nu + 7                       # irlba.jl:50
    if work < nu:
        work = nu + 1
    work = min(work, min(m, n))             # irlba.jl:56-58
    if svtol is None:
        svtol = tol
    if rng is None:
        rng = np.random.default_rng(0)
    if init is None:
        init = rng.standard_normal(n)
    w = work
    V = np.zeros((n, w), order="F")
    W = np.zeros((m, w), order="F")
    B = np.zeros((w, w))
    F = np.zeros(n)
    k = 0
    if restart_from is not None:            # warm restart irlba.jl:87-99 (broken upstream, test_irlba.jl:61)
        U0, s0, V0 = restart_from
        k = len(s0)
        V[:, :k] = V0
        W[:, :k] = U0
        B[np.arange(k), np.arange(k)] = s0
        f0 = _orthog(V, np.array(init, dtype=np.float64), k)   # start vector orthogonal to the supplied V (fixes upstream)
        V[:, k] = f0 / np.linalg.norm(f0)
    else:
        V[:, 0] = init / np.linalg.norm(init)
    sv_prev = np.zeros(w)
    smax = 0.0
    it = 0
    mprod = 0
    info = -2
    BU = BS = BVt = None
    while it < maxit:
        j = k if (it > 0 or restart_from is not None) else 0
        W[:, j] = A.mul(np.ascontiguousarray(V[:, j]), trans=False, parallel=parallel)
        mprod += 1
        if j > 0:
            _orthog(W, W[:, j], j)
        s = np.linalg.norm(W[:, j])
        if s < _EPS23 and j == 0:
            info = -4
            break
        W[:, j] /= s
        while j < w:
            F = A.mul(np.ascontiguousarray(W[:, j]), trans=True, parallel=parallel)
            mprod += 1
            F -= s * V[:, j]
            _orthog(V, F, j + 1)
            if j + 1 < w:
                r = np.linalg.norm(F)
                if r < _EPS23:              # Lanczos breakdown: fresh random direction, B entry 0
                    F = rng.standard_normal(n)
                    _orthog(V, F, j + 1)
                    V[:, j + 1] = F / np.linalg.norm(F)
                    r = 0.0
                else:
                    V[:, j + 1] = F / r
                B[j, j] = s
                B[j, j + 1] = r
                Wn = A.mul(np.ascontiguousarray(V[:, j + 1]), trans=False, parallel=parallel)
                mprod += 1
                Wn -= r * W[:, j]
                _orthog(W, Wn, j + 1)
                s = np.linalg.norm(Wn)
                if s < _EPS23:
                    Wn = rng.standard_normal(m)
                    _orthog(W, Wn, j + 1)
                    W[:, j + 1] = Wn / np.linalg.norm(Wn)
                    s = 0.0
                else:
                    W[:, j + 1] = Wn / s
            else:
                B[j, j] = s
            j += 1
        BU, BS, BVt = np.linalg.svd(B)
        rF = np.linalg.norm(F)
        F = F / rF if rF > 0 else F
        res = rF * BU[w - 1, :]
        smax = max(smax, BS[0])
        ratio = np.abs(sv_prev - BS) / BS
        conv = (np.abs(res) < tol * smax) & (ratio < svtol)
        # count over the nu wanted Ritz values only: counting all `w` (irlb.c's convtests) can stop with the
        # nu-th value unconverged and then misses the reference's own `S.S ≈ svd(X).S` test (1.8e-7 vs sqrt(eps))
        nconv = int(np.count_nonzero(conv[:nu]))
        it += 1
        # invariant subspace (|F| at rounding level): every Ritz value is exact; restarting would seed on noise
        if nconv >= nu or s == 0.0 or rF <= 1000.0 * np.finfo(np.float64).eps * smax:
            info = 0
            break
        if it >= maxit:
            break
        sv_prev = BS.copy()
        k = max(k, nu + nconv)
        k = min(k, w - 3)
        k = max(k, 1)
        V[:, :k] = V @ BVt[:k, :].T
        V[:, k] = F
        W[:, :k] = W @ BU[:, :k]
        B[:] = 0.0
        B[np.arange(k), np.arange(k)] = BS[:k]
        B[:k, k] = res[:k]
    if BU is None:
        return IrlbaResult(np.zeros((m, nu)), np.zeros(nu), np.zeros((n, nu)), it, mprod, info)
    U = W @ BU[:, :nu]
    Vout = V @ BVt[:nu, :].T
    return IrlbaResult(np.asfortranarray(U), BS[:nu].copy(), np.asfortranarray(Vout), it, mprod, info)


# --------------------------------------------------------------------------------------
# C'C (scaling.jl:274-296 over mul.jl:82-114) and tssvd (embedding.jl:30-44)
# --------------------------------------------------------------------------------------
def gram(C: "CenteredMatrix"):
    """scaling.jl:274-296 ``mul!(C, S', S, 1, 0)``, statement by statement: ``A'A`` (the CSC x CSC -> dense triple loop of
    mul.jl:82-114, here scipy's product — same sums, unspecified order), ``q = sum(A, dims=1)``, ``C -= mu*q``,
    ``q -= m*mu``, ``C -= q'*mu'``. Unpinned upstream (no reference test calls it): the tests pin it on the dense
    ``convert(Matrix, S)`` (scaling.jl:298-303) instead."""
    m, n = C.shape
    if C.dense:
        A = np.asarray(C.P, dtype=np.float64)
        G = A.T @ A
        q = A.sum(axis=0)
    else:
        A = C.P.T.tocsc() if C.transposed else C.P
        G = np.asarray((A.T @ A).todense(), dtype=np.float64)
        q = np.asarray(A.sum(axis=0), dtype=np.float64).ravel()
    if C.mu is not None:
        G = G - np.outer(C.mu, q)                 # scaling.jl:284  - M'A
        q = q - m * C.mu                          # scaling.jl:292  axpy!(-size(Al,1), mul, q)
        G = G - np.outer(q, C.mu)                 # scaling.jl:293  + (M'M - A'M)
    return np.asfortranarray(G)


def tssvd(C: "CenteredMatrix", nsv=6):
    """embedding.jl:30-44: eigenpairs of Hermitian(A'A) (Arpack ``eigs`` upstream; LAPACK ``eigh`` here — converged
    eigenpairs are the same up to sign), ``Sigma = sqrt(lambda)``, ``U = A*phi*inv(Diagonal(Sigma))``."""
    G = gram(C)
    lam, phi = np.linalg.eigh((G + G.T) * 0.5)
    order = np.argsort(lam)[::-1][:nsv]
    lam, phi = lam[order], phi[:, order]
    sigma = np.sqrt(lam)
    U = np.column_stack([C.mul(np.ascontiguousarray(phi[:, i])) for i in range(nsv)]) / sigma[None, :]
    return IrlbaResult(np.asfortranarray(U), sigma, np.asfortranarray(phi), 0, 0, 0)


# --------------------------------------------------------------------------------------
# neighbours.jl:19-86 : k-nearest neighbours (the exact answer the reference's test holds its approximate
# libcell search against, test/test_nn.jl:31-38)
# --------------------------------------------------------------------------------------
def knn(X, k, metric="euclidean", include_self=True):
    """Exact k nearest neighbours of every row of ``X`` (n x d): ``partialsortperm`` of the pairwise distances
    (Distances.jl ``Euclidean`` = sqrt(sum (x-y)^2), ``CosineDist`` = max(1 - x.y/(|x||y|), 0)), ties by lower index.
    ``include_self``: the row itself comes first at distance 0 (neighbours.jl `include_self`), otherwise it is excluded.
    Returns 0-based ``nn_index`` (n x k int32) and ``distances`` (n x k, increasing)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    idx = np.empty((n, k), dtype=np.int32)
    dist = np.empty((n, k))
    nrm = np.sqrt((X * X).sum(axis=1))
    for i0 in range(0, n, 512):
        Q = X[i0:i0 + 512]
        if metric == "cosine":
            den = nrm[i0:i0 + 512, None] * nrm[None, :]
            with np.errstate(divide="ignore", invalid="ignore"):
                D = np.maximum(1.0 - np.where(den > 0, (Q @ X.T) / den, 0.0), 0.0)
        elif metric == "euclidean":
            D = np.zeros((Q.shape[0], n))
            for c in range(X.shape[1]):                  # sum_c (q_c - x_c)^2, coordinate by coordinate
                D += (Q[:, c, None] - X[None, :, c]) ** 2
            D = np.sqrt(D)
        else:
            raise ValueError(metric)
        rows = np.arange(Q.shape[0])
        D[rows, i0 + rows] = -1.0 if include_self else np.inf
        order = np.argsort(D, axis=1, kind="stable")[:, :k]
        idx[i0:i0 + 512] = order
        d = np.take_along_axis(D, order, axis=1)
        if include_self:
            d[:, 0] = 0.0
        dist[i0:i0 + 512] = d
    return idx, dist


def nearest_neighbours(X, k, dims=None, metric="euclidean", include_self=True):
    """neighbours.jl:76-80: ``sparse(vec(nn_index'), repeat(1:n, inner=k), trues)`` — entry (j, i) is set when cell j is one
    of the k nearest neighbours of cell i ("k-neighbours are stored as rows for each cell (cols)")."""
    X = np.asarray(X, dtype=np.float64)
    if dims is not None:
        X = X[:, dims]
    idx, _ = knn(X, k, metric, include_self)
    n = X.shape[0]
    nn = sp.csc_matrix((np.ones(n * k, dtype=bool), idx.ravel(), np.arange(0, n * k + 1, k)), shape=(n, n))
    nn.sort_indices()
    return nn


# --------------------------------------------------------------------------------------
# neighbours.jl:88-110 : Jaccard index of the neighbour sets (shared-nearest-neighbour graph), the consumer of the kNN graph
# --------------------------------------------------------------------------------------
def jaccard_index(nn, k=None, prune=1.0 / 15.0, dtype=np.float64):
    """``_jaccard_index`` (neighbours.jl:88-94 with a fixed ``k``, :96-110 without): ``snn = nn' * nn`` — entry (i, j) =
    number of neighbours cells i and j share — converted to ``dtype``; every stored x becomes ``x / (k + (k - x))`` with
    ``k`` the given neighbourhood size or, without it, ``diag(snn)[j]`` of the entry's COLUMN j (:103-106); then
    ``droptol!(snn, prune)`` removes the entries with ``abs(x) <= prune`` (:92,108). ``nn``: n x n sparse pattern, column i =
    the neighbours of cell i (stored entries count as ``true``). Returns CSC with ascending row indices."""
    T = np.dtype(dtype).type
    nn = sp.csc_matrix(nn)
    nn.sort_indices()
    n = nn.shape[1]
    P = sp.csc_matrix((np.ones(nn.nnz, dtype=np.int64), nn.indices, nn.indptr), shape=nn.shape)
    snn = sp.csc_matrix(P.T @ P)
    snn.sort_indices()
    x = snn.data.astype(T)
    col = np.repeat(np.arange(n), np.diff(snn.indptr))
    kk = np.full(x.shape, T(k), dtype=T) if k is not None else snn.diagonal().astype(T)[col]
    with np.errstate(divide="ignore", invalid="ignore"):
        v = x / (kk + (kk - x))
    keep = ~(np.abs(v) <= T(prune))
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(col[keep], minlength=n), out=indptr[1:])
    return sp.csc_matrix((v[keep], snn.indices[keep], indptr), shape=(n, n))


def shared_nearest_neighbours(X, k, dims=None, metric="euclidean", include_self=True, prune=1.0 / 15.0):
    """neighbours.jl:263-270: the kNN graph of ``X`` and its Jaccard index with the fixed ``k``, in the element type of X."""
    X = np.asarray(X)
    T = X.dtype if X.dtype in (np.float32, np.float64) else np.float64
    return jaccard_index(nearest_neighbours(X, k, dims, metric, include_self), k, prune, T)


# --------------------------------------------------------------------------------------
# embedding.jl:46-76 _pca post-processing ; utils.jl:215-228 svd_flip!
# --------------------------------------------------------------------------------------
def pca_post(U, S, V, npcs, m):
    """Z = U[:, :k] * Diagonal(S[:k]); stdev = S[:k] / sqrt(max(1, m-1)); loadings = V[:, :k]."""
    Z = U[:, :npcs] * S[None, :npcs]
    stdev = S[:npcs] / np.sqrt(max(1, m - 1))
    return Z, stdev, V[:, :npcs]


def svd_flip(U, V, u_based_decision=True):
    """utils.jl:215-228: make the max-|.| entry of each U (or V) column positive."""
    M = U if u_based_decision else V
    idx = np.argmax(np.abs(M), axis=0)
    signs = np.sign(M[idx, np.arange(M.shape[1])])
    return U * signs[None, :], V * signs[None, :]


def principal_angle(X, Y):
    """Largest principal angle (radians) between span(X) and span(Y) (both n x k)."""
    Qx, _ = np.linalg.qr(X)
    Qy, _ = np.linalg.qr(Y)
    # sin of the largest angle = ||(I - Qx Qx') Qy||_2 : accurate for small angles
    R = Qy - Qx @ (Qx.T @ Qy)
    return float(np.arcsin(min(1.0, np.linalg.norm(R, 2))))


# --------------------------------------------------------------------------------------
# Host twin of the benchmark's synthetic generator (csrc/synth_twin.c) — the input of bench.py's reference arm
# --------------------------------------------------------------------------------------
@dataclass
class SynthTables:
    """Gene / cell tables of one synthetic configuration (the host part of svb_synth_counts)."""
    m_total: int
    genes: int
    K: int
    seed: int
    row0: int
    rows: int
    lamtab: np.ndarray   # [genes*K] Float64
    lib: np.ndarray      # [rows] Float64 library factors L_i
    prog: np.ndarray     # [rows] uint8 program of every cell
    scale: float


def synth_tables(m_total, genes, mean_nnz_per_cell, programs=64, fold=6.0, seed=20260101, rows=None) -> SynthTables:
    L = lib()
    L.orc_synth_gene_tables.restype = ctypes.c_double
    lo, hi = (0, m_total) if rows is None else rows
    assert lo % 4 == 0 and 1 <= programs <= 255
    K = int(programs)
    lamtab = np.empty(genes * K)
    scale = L.orc_synth_gene_tables(ctypes.c_int64(genes), K, ctypes.c_double(mean_nnz_per_cell), ctypes.c_double(fold),
                                    ctypes.c_uint64(seed), _p(lamtab, _f64p))
    n = hi - lo
    libf = np.empty(max(n, 1))
    prog = np.empty(max(n, 1), dtype=np.uint8)
    L.orc_synth_cell_params(ctypes.c_int64(lo), ctypes.c_int64(n), K, ctypes.c_uint64(seed), _p(libf, _f64p), _p(prog, _u8p))
    return SynthTables(m_total, genes, K, seed, lo, n, lamtab, libf, prog, scale)


def synth_stats(t: SynthTables, hist_bins=1024):
    """One streaming pass over ALL genes: (libsize, gene_nnz, mean, var, hist) — see orc_synth_pass_stats."""
    libsize = np.zeros(t.rows, dtype=np.int64)
    gene_nnz = np.zeros(t.genes, dtype=np.int64)
    mean, var = np.zeros(t.genes), np.zeros(t.genes)
    hist = np.zeros(t.genes * hist_bins, dtype=np.int64)
    over = ctypes.c_int64()
    rc = lib().orc_synth_pass_stats(ctypes.c_int64(t.row0), ctypes.c_int64(t.rows), ctypes.c_int64(t.genes), t.K,
                                    ctypes.c_uint64(t.seed), _p(t.lamtab, _f64p), _p(t.lib, _f64p), _p(t.prog, _u8p), hist_bins,
                                    _p(libsize, _i64p), _p(gene_nnz, _i64p), _p(mean, _f64p), _p(var, _f64p), _p(hist, _i64p),
                                    ctypes.byref(over))
    if rc != 0:
        raise MemoryError("orc_synth_pass_stats")
    if over.value:
        raise ValueError(f"{over.value} counts >= {hist_bins}: raise hist_bins")
    return libsize, gene_nnz, mean, var, hist.reshape(t.genes, hist_bins)


def stdvar_clipped_hist(nrow, hist, gene_nnz, mu, sd, vmax=None):
    """variablefeatures.jl:19-28 from per-gene count histograms (orc_stdvar_clipped_hist)."""
    genes, HB = hist.shape
    if vmax is None:
        vmax = np.sqrt(float(nrow))
    out = np.zeros(genes)
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    sd = np.ascontiguousarray(sd, dtype=np.float64)
    h = np.ascontiguousarray(hist)
    lib().orc_stdvar_clipped_hist(ctypes.c_int64(nrow), ctypes.c_int64(genes), HB, _p(h, _i64p), _p(gene_nnz, _i64p), _p(mu, _f64p),
                                  _p(sd, _f64p), ctypes.c_double(float(vmax)), _p(out, _f64p))
    return out


def synth_columns(t: SynthTables, sel):
    """The gene columns ``sel`` (0-based, any order) as raw CSC arrays (colptr int64, rowval int64, counts int32)."""
    sel = np.ascontiguousarray(sel, dtype=np.int64)
    n = sel.shape[0]
    L = lib()
    colnnz = np.zeros(n, dtype=np.int64)
    args = (ctypes.c_int64(t.row0), ctypes.c_int64(t.rows), t.K, ctypes.c_uint64(t.seed), _p(t.lamtab, _f64p), _p(t.lib, _f64p),
            _p(t.prog, _u8p), _p(sel, _i64p), ctypes.c_int64(n))
    L.orc_synth_columns_count(*args, _p(colnnz, _i64p))
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(colnnz, out=colptr[1:])
    nnz = int(colptr[-1])
    rowval = np.empty(max(nnz, 1), dtype=np.int64)
    val = np.empty(max(nnz, 1), dtype=np.int32)
    L.orc_synth_columns_fill(*args, _p(colptr, _i64p), _p(rowval, _i64p), _p(val, _i32p))
    return colptr, rowval[:nnz], val[:nnz]


def synthetic_counts(m_total, genes, mean_nnz_per_cell, programs=64, fold=6.0, seed=20260101, rows=None):
    """Host twin of severo_jl_b200.synthetic_counts: the same matrix, bit for bit (scipy CSC, int64 counts)."""
    t = synth_tables(m_total, genes, mean_nnz_per_cell, programs, fold, seed, rows)
    colptr, rowval, val = synth_columns(t, np.arange(genes))
    return sp.csc_matrix((val.astype(np.int64), rowval, colptr), shape=(t.rows, genes))


def lognorm_columns(rowval, counts_i32, libsize, scale_factor=1e4):
    """normalize.jl:25-29,36 on HVG columns with the library sizes of the full matrix (Y[:, hvf] without forming Y)."""
    out = np.empty(counts_i32.shape[0])
    lib().orc_lognorm_i32(ctypes.c_int64(counts_i32.shape[0]), _p(rowval, _i64p), _p(counts_i32, _i32p), _p(libsize, _i64p),
                          ctypes.c_double(scale_factor), _p(out, _f64p))
    return out


def scale_data_arrays(nrow, colptr, nzval, scale_max=np.inf):
    """scaling.jl:199-217 on raw CSC arrays (no scipy copies: the full-size reference arm holds 7.5e8 nonzeros)."""
    ncol = colptr.shape[0] - 1
    out = np.empty(nzval.shape[0])
    mu = np.empty(ncol)
    lib().orc_scale_data_f64(ctypes.c_int64(nrow), ctypes.c_int64(ncol), _p(colptr, _i64p), _p(nzval, _f64p),
                             ctypes.c_double(float(scale_max)), _p(out, _f64p), _p(mu, _f64p))
    return out, mu


def centered_from_arrays(nrow, colptr, rowval, nzval, mu):
    """CenteredMatrix over caller-owned CSC arrays (int64 indices, Float64 values) without copying them."""
    C = CenteredMatrix.__new__(CenteredMatrix)
    ncol = colptr.shape[0] - 1
    C.dense = False
    C.transposed = False
    C.P = sp.csc_matrix((nzval, rowval, colptr), shape=(nrow, ncol), copy=False)
    C.colptr, C.rowval, C.nz = colptr, rowval, nzval
    C._csr = None
    C.shape = (nrow, ncol)
    C.mu = np.ascontiguousarray(mu, dtype=np.float64)
    return C


# --------------------------------------------------------------------------------------
# mul.jl:50-77 (SpMSpV) and mul.jl:79-114 (CSC x CSC -> dense): the reference loops, literally (csrc/severo_oracle.c)
# --------------------------------------------------------------------------------------
def spmspv(y, A, x, alpha=1.0, beta=0.0):
    """mul!(y, A::SparseMatrixCSC, x::SparseVector, alpha, beta) — in place. ``x``: scipy sparse n x 1 (stored entries count)."""
    A = _csc(A).astype(np.float64)
    xs = sp.csc_matrix(x)
    xs.sort_indices()
    m, n = A.shape
    assert xs.shape == (n, 1) and y.shape == (m,)          # DimensionMismatch
    colptr, rowval = _idx64(A)
    nz = np.ascontiguousarray(A.data, dtype=np.float64)
    xi = np.ascontiguousarray(xs.indices, dtype=np.int64)
    xv = np.ascontiguousarray(xs.data, dtype=np.float64)
    lib().orc_spmspv(ctypes.c_int64(m), ctypes.c_int64(n), _p(colptr, _i64p), _p(rowval, _i64p), _p(nz, _f64p), _p(xi, _i64p),
                     _p(xv, _f64p), ctypes.c_int64(xi.shape[0]), ctypes.c_double(alpha), ctypes.c_double(beta), _p(y, _f64p))
    return y


def spgemm_dense(C, A, B, alpha=1.0, beta=0.0, transpose_a=False):
    """mul!(C, A, B, alpha, beta) for CSC A, B and dense column-major C; ``transpose_a``: mul.jl:79-80, mul!(C, copy(A'), B, ...)."""
    A = _csc(sp.csc_matrix(A).T) if transpose_a else _csc(A)
    A = A.astype(np.float64)
    B = _csc(B).astype(np.float64)
    m, p = A.shape[0], B.shape[1]
    assert A.shape[1] == B.shape[0] and C.shape == (m, p) and C.flags.f_contiguous
    acp, arv = _idx64(A)
    bcp, brv = _idx64(B)
    lib().orc_spgemm_dense(ctypes.c_int64(m), ctypes.c_int64(p), _p(acp, _i64p), _p(arv, _i64p),
                           _p(np.ascontiguousarray(A.data), _f64p), _p(bcp, _i64p), _p(brv, _i64p),
                           _p(np.ascontiguousarray(B.data), _f64p), ctypes.c_double(alpha), ctypes.c_double(beta), _p(C, _f64p))
    return C
