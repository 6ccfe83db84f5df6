"""Host-side mirror of the reference's two `Sys` objects and their main-loop order (c++/bpmf.cpp:180-190), driving
libbpmf_b200.so through its C ABI. One process drives one GPU; with torch.distributed initialised, items of both
factors are split into contiguous ranges (Sys::from()/to(), c++/bpmf.h:171-172) and the freshly sampled columns are
exchanged after every sweep, which replaces Sys::bcast / send_item of the MPI back ends (c++/bpmf.cpp:263-278,
c++/mpi_isendirecv.h).  torch is used for device buffers, streams and the NCCL plumbing only.

Nothing here computes on the CPU; without the CUDA library or a GPU the constructor raises.
"""
import numpy as np

from . import capi

MOVIES, USERS = capi.MOVIES, capi.USERS


def split_range(n, world, rank):
    """Equal-count contiguous slices, ceil(n / world) each (the last ones may be short or empty)."""
    chunk = (n + world - 1) // world
    lo = min(n, rank * chunk)
    return lo, min(n, lo + chunk), chunk


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
        self.items = [None, None]     # torch tensors the library's latent matrices are bound to (padded to chunk * world)
        for side in (MOVIES, USERS):
            n, n_other, ptr, idx, val = ratings.side(side)
            self.num[side] = n
            self.ctx.load_side(side, n, n_other, ptr, idx, val, ratings.mean_rating)
            lo, hi, chunk = split_range(n, self.world, self.rank)
            self.range[side] = (lo, hi, chunk)
            self.ctx.set_range(side, lo, hi)
            buf = torch.zeros(chunk * self.world, K, dtype=torch.float64, device="cuda:%d" % device)
            self.ctx.bind_items(side, buf.data_ptr())
            self.items[side] = buf
        if with_test:
            for side in (MOVIES, USERS):
                self.ctx.load_test(side, *ratings.test_side(side))
        self.nnz = ratings.nnz

    # ---- one Sys::sample(other) (sample.cpp:341-385), multi-GPU aware -------------------------------------------
    def sample(self, side):
        ctx = self.ctx
        it = ctx.get_iter(side) + 1
        ctx.set_iter(side, it)
        ctx.sample_hyper(side, it)                       # every rank draws the same (mu, Lambda): cov is replicated
        ctx.sample_items(side, it, self.alpha, self.variant)
        self._exchange(side)
        ctx.reduce_stats(side)                           # over ALL items, fixed order: identical on every rank

    def _exchange(self, side):
        if self.world == 1:
            return
        lo, hi, chunk = self.range[side]
        buf = self.items[side]
        if self.exchange == "allgather":
            self.dist.all_gather_into_tensor(buf, buf[self.rank * chunk:(self.rank + 1) * chunk])
        elif self.exchange == "push":
            # the kernel already wrote into the peers; a barrier makes every rank's stores visible everywhere
            self.dist.barrier(device_ids=[self.device])
        else:
            raise ValueError(self.exchange)

    def step(self):
        """movies.sample(users); users.sample(movies)  (bpmf.cpp:184-185)"""
        self.sample(MOVIES)
        self.sample(USERS)

    def predict(self, burnin):
        return self.ctx.predict(MOVIES, burnin), self.ctx.predict(USERS, burnin)

    def items_host(self, side):
        return self.items[side][: self.num[side]].cpu().numpy()

    def close(self):
        self.ctx.close()
