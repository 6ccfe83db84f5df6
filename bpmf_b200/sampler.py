"""Host-side mirror of the reference's two `Sys` objects and their main-loop order (c++/bpmf.cpp:180-190), driving
libbpmf_b200.so through its C ABI. One process drives one GPU; with torch.distributed initialised, items of both
factors are split into contiguous ranges (Sys::from()/to(), c++/bpmf.h:171-172) and the freshly sampled columns are
exchanged after every sweep, which replaces Sys::bcast / send_item of the MPI back ends (c++/bpmf.cpp:263-278,
c++/mpi_isendirecv.h).  torch is used for device buffers, streams and the NCCL plumbing only.

Nothing here computes on the CPU; without the CUDA library or a GPU the constructor raises.
"""
import numpy as np

from . import capi
from .partition import allgather_slices, balanced_ranges, split_range

MOVIES, USERS = capi.MOVIES, capi.USERS


class GibbsSampler:
    """movies / users pair on one GPU (one rank).

    exchange: "allgather" (NCCL all-gather of the fresh slice into every replica) or "push" (the item kernel stores
    every fresh K-vector straight into all peer replicas over NVLink while it runs; a barrier follows)."""

    def __init__(self, ratings, K, device=0, alpha=2.0, variant=capi.KERNEL_AUTO, exchange="allgather", with_test=True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.K, self.alpha, self.variant = K, alpha, variant
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.exchange = exchange if self.world > 1 else "none"
        self.device = device
        torch.cuda.set_device(device)
        self.ctx = capi.Context(K, device)
        self.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        self.num = [0, 0]
        self.range = [None, None]
        self.items = [None, None]     # allgather mode: torch tensors the latent matrices are bound to (padded to chunk * world)
        self._tiny = torch.zeros(1, device="cuda:%d" % device)
        for side in (MOVIES, USERS):
            n, n_other, ptr, idx, val = ratings.side(side)
            self.num[side] = n
            if self.exchange == "push":
                # ragged, work-balanced ranges on statistics-block boundaries; the library's own (cudaMalloc) storage is what
                # CUDA IPC can export. Only the ratings of the rank's own items become resident (c++/bpmf.h:161-176).
                b = balanced_ranges(ptr, self.world, align=capi.stats_block_items_for(K, n))
                lo, hi, chunk = int(b[self.rank]), int(b[self.rank + 1]), 0
                self.ctx.load_side_slice(side, n, n_other, lo, hi, ptr, idx, val, ratings.mean_rating)
            else:
                self.ctx.load_side(side, n, n_other, ptr, idx, val, ratings.mean_rating)
                lo, hi, chunk = split_range(n, self.world, self.rank)
                buf = torch.zeros(chunk * self.world, K, dtype=torch.float64, device="cuda:%d" % device)
                self.ctx.bind_items(side, buf.data_ptr())
                self.items[side] = buf
            self.range[side] = (lo, hi, chunk)
            self.ctx.set_range(side, lo, hi)
        if self.exchange == "push":
            for side in (MOVIES, USERS):
                handles = [None] * self.world
                dist.all_gather_object(handles, self.ctx.ipc_export(side))
                ptrs = [self.ctx.items_device_ptr(side) if r == self.rank else self.ctx.ipc_open(handles[r])
                        for r in range(self.world)]
                self.ctx.set_peers(side, ptrs)
                # the sweep statistics: every rank reduces the blocks of its own range and stores them into all ranks' buffers
                dist.all_gather_object(handles, self.ctx.ipc_export_stats(side))
                sptrs = [self.ctx.stats_device_ptr(side) if r == self.rank else self.ctx.ipc_open(handles[r])
                         for r in range(self.world)]
                self.ctx.set_stats_peers(side, sptrs)
        self._readers = False     # a host-initiated read of a replica is in flight that the next remote stores must not overtake
        if with_test:
            for side in (MOVIES, USERS):
                self.ctx.load_test(side, *ratings.test_side(side))
        self.nnz = ratings.nnz

    # ---- one Sys::sample(other) (sample.cpp:341-385), multi-GPU aware -------------------------------------------
    def sample(self, side):
        ctx = self.ctx
        self._fence_readers()
        if self.exchange == "push":
            # one C call: hyper draw (every rank draws the same (mu, Lambda): cov is replicated), the item kernel storing each
            # fresh column into every replica, own statistics blocks into every rank's buffer, the device-side cross-GPU
            # barrier (bpmf_gpu_peer_barrier), the fixed-order sum — identical on every rank, no NCCL call in the sweep
            ctx.sample(side, self.alpha, self.variant)
            return
        it = ctx.get_iter(side) + 1
        ctx.set_iter(side, it)
        ctx.sample_hyper(side, it)
        ctx.sample_items(side, it, self.alpha, self.variant)
        self._exchange(side)
        ctx.reduce_stats(side)                           # over ALL items, fixed order: identical on every rank

    def _fence_readers(self):
        """Push mode: a peer's next sweep stores into THIS rank's replicas while it runs. Reads that this rank enqueued after
        its last sweep (predict, items_host, downloads) are ordered before those stores by one more cross-rank barrier, issued
        only when such a read happened since the last sweep. SPMD: every rank reads at the same points of the program."""
        if self._readers and self.exchange == "push":
            self.ctx.peer_barrier(MOVIES)
        self._readers = False

    def _exchange(self, side):
        if self.world == 1:
            return
        if self.exchange == "allgather":
            allgather_slices(self.dist, self.items[side], self.rank, self.world)
        elif self.exchange == "push":
            self.ctx.peer_barrier(side)
        else:
            raise ValueError(self.exchange)

    def upload_slice(self, side, host_ptr):
        """Multi-GPU hosts hold one slice of a latent matrix per rank: upload this rank's slice of `side` from the host
        matrix at `host_ptr` and distribute it to every replica (all-gather, or NVLink copies into the peer replicas)."""
        lo, hi, _ = self.range[side]
        self.ctx.set_items_range_ptr(side, lo, hi, host_ptr)
        if self.world == 1:
            return
        if self.exchange == "allgather":
            allgather_slices(self.dist, self.items[side], self.rank, self.world)
        else:
            self.ctx.push_range(side, lo, hi)
            self.ctx.peer_barrier(side)           # every rank's copies are ordered before the next kernel of any rank

    def sample_host(self, side, host_other_ptr, host_items_ptr):
        """Sys::sample(Sys &other) with HOST-resident latent matrices (pinned): other.items() is read from host memory, this
        side's fresh items() end up in host memory. One GPU: bpmf_gpu_sample_host. Several: every rank's host memory holds
        ITS slice of each matrix (host_*_ptr are the bases of full-size matrices of which only the slice is touched)."""
        if self.world == 1:
            self.ctx.sample_host(side, host_other_ptr, host_items_ptr, self.alpha, self.variant)
            return
        if self.exchange != "push":
            self.upload_slice(1 - side, host_other_ptr)
            self.sample(side)
            lo, hi, _ = self.range[side]
            self.ctx.get_items_range_ptr(side, lo, hi, host_items_ptr)    # synchronises this rank's stream
            return
        # push exchange: one C call per rank (bpmf_gpu_sample_host runs the multi-rank protocol by itself): chunked upload of
        # the slice of the other side, each chunk forwarded to the peers over NVLink while the next one is uploaded; device-side
        # barrier; the sweep in parts, each part downloaded while the next is sampled; barrier; statistics; synchronises
        self._fence_readers()
        self.ctx.sample_host(side, host_other_ptr, host_items_ptr, self.alpha, self.variant)

    def step(self):
        """movies.sample(users); users.sample(movies)  (bpmf.cpp:184-185)"""
        self.sample(MOVIES)
        self.sample(USERS)

    def predict(self, burnin):
        self._readers = True
        return self.ctx.predict(MOVIES, burnin), self.ctx.predict(USERS, burnin)

    def items_host(self, side):
        self._readers = True
        return self.ctx.get_items(side)

    def items_view(self, side):
        """device tensor of the bound latent matrix (allgather mode) or None (push mode: library-owned storage)"""
        return None if self.items[side] is None else self.items[side][: self.num[side]]

    def close(self):
        self.ctx.close()
