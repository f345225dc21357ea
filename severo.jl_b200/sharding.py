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
