"""Host-side partitioning / exchange logic of the multi-GPU sweep. Device-agnostic so it can be
exercised with the gloo backend on CPU tensors; sampler.py uses it with NCCL on the GPU.

The reference splits both factors into contiguous index ranges, one per node (Sys::from()/to(), c++/bpmf.h:171-172,
c++/assign.cpp:52-209) and ships every freshly sampled column to the nodes that need it (Sys::bcast,
c++/bpmf.cpp:263-278; send_item in the MPI/GASPI back ends). Here every rank keeps a full replica of U and V, so the
exchange is one all-gather of the fresh slices per sweep.
"""
import numpy as np


def split_range(n, world, rank):
    """Equal-count contiguous slices, ceil(n / world) items each (trailing ones may be short or empty).
    Returns (lo, hi, chunk). Equal chunks are what an in-place all_gather_into_tensor needs."""
    chunk = (n + world - 1) // world
    lo = min(n, rank * chunk)
    return lo, min(n, lo + chunk), chunk


def balanced_ranges(colptr, world, fixed_cost=64, align=1):
    """Contiguous ranges balanced on work = fixed_cost + nnz per item, the reference's heuristic (c++/assign.cpp:111 uses
    10 + nnz; on the B200 the per-item tail — normals, LDL^T, solves: 3.0 ms per million items — costs as much as the gather
    and Gram of ~65 ratings, 4.7 ms per 100 M, at K=32: profiles/r01_probes.log, r02_tune_chain_probes.log).
    Returns world+1 boundaries. Used by the push exchange, which does not need equal-sized slices.
    align > 1 rounds the inner boundaries to multiples of it (the statistics-block size: every rank then reduces whole
    blocks of the fixed decomposition, bpmf_gpu_stats_block_items)."""
    colptr = np.asarray(colptr, np.int64)
    n = len(colptr) - 1
    work = colptr[1:] - colptr[:-1] + fixed_cost
    cum = np.concatenate([[0], np.cumsum(work)])
    targets = cum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(cum, targets, side="left")
    if align > 1:
        cuts = ((cuts + align // 2) // align) * align
    b = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    return np.maximum.accumulate(np.minimum(b, n))


def padded_items(n, world):
    """Number of items the replica buffer must hold for an equal-chunk all-gather."""
    return ((n + world - 1) // world) * world


def allgather_slices(dist, buf, rank, world):
    """In-place all-gather of rank-major equal chunks of `buf` (items x K): every rank contributes
    buf[rank*chunk:(rank+1)*chunk] and ends up with all of them."""
    chunk = buf.shape[0] // world
    assert chunk * world == buf.shape[0]
    mine = buf[rank * chunk:(rank + 1) * chunk]
    if dist.get_backend() == "gloo":   # gloo has no all_gather_into_tensor for in-place views
        parts = [buf[r * chunk:(r + 1) * chunk].clone() for r in range(world)]
        dist.all_gather(parts, mine.clone())
        for r in range(world):
            buf[r * chunk:(r + 1) * chunk].copy_(parts[r])
    else:
        dist.all_gather_into_tensor(buf, mine)
