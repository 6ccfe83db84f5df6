"""Host-side logic of the multi-GPU sweep on CPU: world_size 2 with the gloo backend.

Each rank runs the ORACLE's per-item draws for its own range (the checker stands in for the CUDA kernel here, as the
container has no GPU), exchanges the fresh slices exactly the way bpmf_b200.sampler does, and the result must equal a
single-process oracle sweep. This pins: split_range / balanced_ranges, the padded in-place all-gather, and that the
protocol (hyper draw replicated on every rank from the replicated cov, per-range draws, exchange, global reduction)
is partition-independent."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from bpmf_b200 import partition

K = 16


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    train, test = util.synth_ratings(91, 67, 1500, 5, empty_rows=7, heavy_col=40)
    orc = util.make_oracle(K, train, test, nthreads=1)
    out = {}
    for it in range(3):
        for side in (util.MOVIES, util.USERS):
            n = orc.num(side)
            # --- what GibbsSampler.sample does, with the oracle standing in for the item kernel
            orc.set_iter(side, it)
            cov = torch.from_numpy(orc.stats(side)[2].copy())
            ref = cov.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(cov, ref)                       # cov is replicated bit for bit -> same hyper draw everywhere
            from oracle import oracle as o
            mu, LU, LF = o.hyper(K, n, it, cov.numpy())
            orc.set_hyper(side, mu, LF)
            lo, hi, chunk = partition.split_range(n, world, rank)
            before = orc.items(side).copy()
            orc.sample_range(side, lo, hi)
            buf = torch.zeros(partition.padded_items(n, world), K, dtype=torch.float64)
            buf[:n] = torch.from_numpy(orc.items(side))
            buf[:lo] = torch.from_numpy(before[:lo])           # other ranks' slices are stale until the exchange
            partition.allgather_slices(dist, buf, rank, world)
            x = buf[:n].numpy().copy()
            orc.set_items(side, x)
            # global reduction over ALL items -> cov (identical on every rank)
            s, p = x.sum(0), x.T @ x
            orc.set_cov(side, ((p - np.outer(s, s) / n) / (n - 1)).T.copy())
            out[(it, side)] = x
    q.put((rank, {k: v.tobytes() for k, v in out.items()}))
    dist.destroy_process_group()


def test_two_ranks_equal_one_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1]                                    # replicas identical on both ranks
    # single process: same protocol with world = 1
    train, test = util.synth_ratings(91, 67, 1500, 5, empty_rows=7, heavy_col=40)
    orc = util.make_oracle(K, train, test, nthreads=1)
    from oracle import oracle as o
    for it in range(3):
        for side in (util.MOVIES, util.USERS):
            n = orc.num(side)
            orc.set_iter(side, it)
            mu, LU, LF = o.hyper(K, n, it, orc.stats(side)[2])
            orc.set_hyper(side, mu, LF)
            orc.sample_range(side, 0, n)
            x = orc.items(side)
            s, p = x.sum(0), x.T @ x
            orc.set_cov(side, ((p - np.outer(s, s) / n) / (n - 1)).T.copy())
            assert x.tobytes() == res[0][(it, side)], (it, side)


def test_split_and_balanced_ranges():
    for n, w in [(10, 3), (1000000, 8), (7, 8), (0, 2), (1682, 2)]:
        spans = [partition.split_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and max(s[1] for s in spans) == n
        assert all(a[1] == b[0] or b[0] == n for a, b in zip(spans, spans[1:]))
        assert partition.padded_items(n, w) >= n and partition.padded_items(n, w) % w == 0
    rng = np.random.default_rng(0)
    nnz = rng.poisson(3, 5000)
    nnz[17] = 110000                                            # a ChEMBL-style hot column
    colptr = np.concatenate([[0], np.cumsum(nnz)])
    b = partition.balanced_ranges(colptr, 8)
    assert b[0] == 0 and b[-1] == 5000 and np.all(np.diff(b) >= 0)
    work = np.diff(colptr) + 64
    loads = [work[b[r]:b[r + 1]].sum() for r in range(8)]
    assert max(loads) <= work[17] + work.sum() / 8              # no rank holds more than the hot item plus a fair share


def test_balanced_ranges_on_statistics_block_boundaries():
    """push exchange: inner range boundaries are multiples of the statistics-block size (bpmf_gpu_stats_block_items), so
    every rank reduces whole blocks of the fixed decomposition"""
    rng = np.random.default_rng(3)
    n = 100003
    colptr = np.concatenate([[0], np.cumsum(rng.poisson(40, n))])
    for world in (2, 3, 8):
        for align in (1, 344, 3392):
            b = partition.balanced_ranges(colptr, world, align=align)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
            assert all(int(x) % align == 0 for x in b[1:-1])
            work = np.diff(colptr[b]) + 64 * np.diff(b)
            assert work.max() <= work.mean() * (1.0 + 1.5 * world * align * 104.0 / work.sum()) + 1   # within ~one block (40 + 64 per item)
