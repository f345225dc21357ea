"""Cell sharding for 2/4/8 GPUs: one process per GPU, contiguous cell ranges.

``torch.distributed`` is plumbing only: it carries the 128-byte NCCL unique id from rank 0 to the
other ranks; the data-path allreduce is issued by the library itself on its own stream (csrc/comm.cpp).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib as L


def shard_bounds(m_total: int, nranks: int, row_nnz=None, align: int = 4):
    """Contiguous [lo, hi) cell ranges, one per rank; nnz-balanced when per-cell counts are given.
    Boundaries are multiples of ``align`` (the generator draws cells in quads)."""
    if nranks < 1:
        raise ValueError("nranks must be >= 1")
    if row_nnz is None:
        cuts = [(m_total * r) // nranks for r in range(nranks + 1)]
    else:
        csum = np.concatenate([[0], np.cumsum(np.asarray(row_nnz, dtype=np.int64))])
        total = csum[-1]
        cuts = [0]
        for r in range(1, nranks):
            cuts.append(int(np.searchsorted(csum, total * r / nranks)))
        cuts.append(m_total)
    cuts = [min(m_total, (c // align) * align) if 0 < i < nranks else c for i, c in enumerate(cuts)]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(nranks)]


def init_comm_from_torch():
    """Join the library's NCCL communicator using the already-initialised torch.distributed group."""
    import torch
    import torch.distributed as dist

    lib = L.lib()
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return 1, 0
    idbuf = (ctypes.c_ubyte * 128)()
    if rank == 0:
        L.check(lib.svb_comm_unique_id(ctypes.cast(idbuf, ctypes.c_void_p)))
    t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    buf = (ctypes.c_ubyte * 128).from_buffer_copy(raw)
    L.check(lib.svb_comm_init(world, rank, ctypes.cast(buf, ctypes.c_void_p)))
    return world, rank


def gather_rows(local: np.ndarray, bounds, rank: int, group=None):
    """All-gather the row blocks of a cell-sharded dense array (e.g. U) through torch.distributed."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    if world == 1:
        return local
    k = local.shape[1]
    m_total = bounds[-1][1]
    out = np.zeros((m_total, k), order="F")
    for r, (lo, hi) in enumerate(bounds):
        t = torch.from_numpy(np.ascontiguousarray(local if r == rank else np.zeros((hi - lo, k))))
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=r, group=group)
        out[lo:hi] = t.cpu().numpy()
    return out


# ------------------------------------------------------------------------------------------------
# cell-sharded pre-processing: per-rank sweeps + a merge of per-gene statistics.
# "fast mode": the merged moments are mathematically the moments of the whole matrix but are not the
# bits of one sequential Welford over all cells (that needs a rank-ordered carry; single-GPU runs use
# the order-exact kernel). Every rank performs the same merge in rank order -> identical results.
# ------------------------------------------------------------------------------------------------
def allgather_f64(local: np.ndarray) -> np.ndarray:
    """[nranks, len(local)] via the library's own allreduce (zero-padded slots)."""
    lib = L.lib()
    nr, rk = ctypes.c_int(), ctypes.c_int()
    lib.svb_comm_info(ctypes.byref(nr), ctypes.byref(rk))
    local = np.ascontiguousarray(local, dtype=np.float64).ravel()
    buf = np.zeros((nr.value, local.shape[0]))
    buf[rk.value] = local
    L.check(lib.svb_comm_allreduce_f64(L.ptr(buf), buf.size))
    return buf


def merged_mean_var(dA):
    """Per-gene mean / unbiased variance of the whole (row-sharded) matrix from per-rank Welford moments,
    merged with Chan's pairwise formula in rank order."""
    from .api import mean_var
    mu_r, var_r = mean_var(dA)
    n_r = float(dA.shape[0])
    g = mu_r.shape[0]
    packed = allgather_f64(np.concatenate([[n_r], np.asarray(mu_r, dtype=np.float64), np.asarray(var_r, dtype=np.float64)]))
    n = 0.0
    mean = np.zeros(g)
    m2 = np.zeros(g)
    for row in packed:
        nb, mb, vb = row[0], row[1:1 + g], row[1 + g:]
        if nb == 0:
            continue
        m2b = vb * (nb - 1.0) if nb > 1 else np.zeros(g)
        delta = mb - mean
        tot = n + nb
        mean = mean + delta * (nb / tot)
        m2 = m2 + m2b + delta * delta * (n * nb / tot)
        n = tot
    return mean, m2 / (n - 1.0), int(n)


def sharded_vst_metric(counts, loess_span=0.5, expected_std_fn=None):
    """variance_stabilizing_transformation (variablefeatures.jl:34-50) over a row-sharded count matrix."""
    from .api import standardized_var_clipped, _call_trend
    from .loess import loess_fit_predict
    mu, var, m_total = merged_mean_var(counts)
    sd = np.sqrt(var)
    non_const = sd > 0
    expected = sd.copy()
    if expected_std_fn is not None:
        expected[non_const] = _call_trend(expected_std_fn, mu[non_const], sd[non_const])
    else:
        expected[non_const] = 10.0 ** loess_fit_predict(np.log10(mu[non_const]), np.log10(sd[non_const]), span=loess_span)
    expected = np.where(np.isnan(expected), 0.0, expected)
    # per-rank numerator sum_nz clip(.)^2 + (#zeros) clip(0)^2 is additive over cells; vmax uses the global m
    part = standardized_var_clipped(counts, mu, expected, vmax=np.sqrt(float(m_total))) * (counts.shape[0] - 1.0)
    total = allgather_f64(part).sum(axis=0)
    return total / (m_total - 1.0)


def sharded_scale_features(Y, scale_max=np.inf, features=None):
    """scale_features (scaling.jl:335-357) over a row-sharded matrix: returns (DeviceMatrix B_local, mu)."""
    from .api import DeviceMatrix, _DT
    if features is not None:
        Y = Y.columns(np.asarray(features, dtype=np.int64))
    mean, var, _ = merged_mean_var(Y)
    mu = np.empty(Y.shape[1])
    h = ctypes.c_void_p()
    dtype = np.float32 if Y.vtype == L.SVB_F32 else np.float64
    L.check(L.lib().svb_scale_with_moments(Y._h, L.ptr(np.ascontiguousarray(mean)), L.ptr(np.ascontiguousarray(var)),
                                            float(scale_max), _DT[np.dtype(dtype)], ctypes.byref(h), L.ptr(mu)))
    return DeviceMatrix(h), mu


def exact_mean_var(dA):
    """Order-exact per-gene mean / variance of a row-sharded matrix: the sequential Welford state of every gene
    is carried from rank to rank in rank order, so the result has the bits of one pass over all cells
    (scaling.jl:18-34). Serial over ranks by construction — use ``merged_mean_var`` when speed matters more."""
    lib = L.lib()
    nr, rk = ctypes.c_int(), ctypes.c_int()
    lib.svb_comm_info(ctypes.byref(nr), ctypes.byref(rk))
    nranks, rank = nr.value, rk.value
    g = dA.shape[1]
    nnz_local = np.diff(dA.colptr()).astype(np.float64)
    tot = allgather_f64(np.concatenate([[float(dA.shape[0])], nnz_local])).sum(axis=0)
    m_total, nnz_total = int(tot[0]), tot[1:]
    count = (m_total - nnz_total).astype(np.int64)   # implicit zeros first (scaling.jl:21)
    mu = np.zeros(g)
    s = np.zeros(g)
    for r in range(nranks):
        if rank == r:
            L.check(lib.svb_welford_carry(dA._h, L.ptr(count), L.ptr(mu), L.ptr(s)))
            packed = np.concatenate([count.astype(np.float64), mu, s])
        else:
            packed = np.zeros(3 * g)
        state = allgather_f64(packed)[r] if nranks > 1 else packed
        count, mu, s = state[:g].astype(np.int64), state[g:2 * g].copy(), state[2 * g:].copy()
    return mu, s / (m_total - 1.0), m_total
