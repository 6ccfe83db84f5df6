"""torchrun worker: N ranks run a few Gibbs iterations with the partitioned sweep and must reproduce, bit for bit, the
latent matrices a single GPU produces (results are partition-independent: every draw is keyed by item index and
iteration, and the sweep reductions run over all items in a fixed order).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_worker.py [allgather|push]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "allgather"
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from bpmf_b200 import synthetic
    from bpmf_b200.sampler import GibbsSampler, MOVIES, USERS
    ratings = synthetic.generate(3001, 2500, 40.0, 99)     # sizes not divisible by the world size on purpose
    K = 32
    multi = GibbsSampler(ratings, K, device=local, exchange=mode)
    # the single-GPU reference run, on this rank's own GPU, outside the process group's partitioning
    import bpmf_b200
    single = bpmf_b200.Context(K, local)
    for side in (MOVIES, USERS):
        n, n_other, ptr, idx, val = ratings.side(side)
        single.load_side(side, n, n_other, ptr, idx, val, ratings.mean_rating)
    for side in (MOVIES, USERS):
        single.load_test(side, *ratings.test_side(side))
    worst = 0.0
    for it in range(4):
        multi.step()
        single.sample(MOVIES)
        single.sample(USERS)
        for side in (MOVIES, USERS):
            a, b = multi.items_host(side), single.get_items(side)
            worst = max(worst, float(np.abs(a - b).max()))
            assert a.tobytes() == b.tobytes(), "rank %d iteration %d side %d differs by %g" % (rank, it, side, np.abs(a - b).max())
    # readers of a replica after a sweep (predict, downloads) against a peer that is already in its next sweep: rank 1 lags
    # behind before it reads; the sampler's reader fence must keep the peers' next remote stores behind that read
    import time
    for it in range(2):
        multi.step()
        single.sample(MOVIES); single.sample(USERS)
        if rank == world - 1:
            time.sleep(0.3)
        got = multi.predict(burnin=1)
        ref = (single.predict(MOVIES, 1), single.predict(USERS, 1))
        assert got == ref, "rank %d: predict after a lagging read differs: %r vs %r" % (rank, got, ref)
        for side in (MOVIES, USERS):
            assert multi.items_host(side).tobytes() == single.get_items(side).tobytes()
    # the end-to-end path of bench.py: every rank keeps ITS slice of each latent matrix in (pinned) host memory
    host = [torch.from_numpy(multi.items_host(s)).pin_memory() for s in (MOVIES, USERS)]
    for it in range(2):
        for side in (MOVIES, USERS):
            lo, hi, _ = multi.range[1 - side]
            other = host[1 - side].clone()
            other[:lo] = float("nan"); other[hi:] = float("nan")      # only the own slice may be read
            other = other.pin_memory()
            mine = torch.full_like(host[side], float("nan")).pin_memory()
            multi.sample_host(side, other.data_ptr(), mine.data_ptr())
            lo, hi, _ = multi.range[side]
            single.sample(side)
            ref = single.get_items(side)
            assert mine.numpy()[lo:hi].tobytes() == ref[lo:hi].tobytes(), "e2e slice path differs (rank %d)" % rank
            assert np.isnan(mine.numpy()[:lo]).all() and np.isnan(mine.numpy()[hi:]).all()      # only the own slice is written
            assert multi.items_host(side).tobytes() == ref.tobytes()
            host[side] = torch.from_numpy(ref.copy()).pin_memory()
    multi.ctx.sync()
    t = torch.tensor([worst], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("multi-gpu %s x%d: identical to single GPU over 4 iterations (max diff %g)" % (mode, world, t.item()))
    multi.close(); single.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
