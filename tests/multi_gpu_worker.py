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
    worst = 0.0
    for it in range(4):
        multi.step()
        single.sample(MOVIES)
        single.sample(USERS)
        for side in (MOVIES, USERS):
            a, b = multi.items_host(side), single.get_items(side)
            worst = max(worst, float(np.abs(a - b).max()))
            assert a.tobytes() == b.tobytes(), "rank %d iteration %d side %d differs by %g" % (rank, it, side, np.abs(a - b).max())
    multi.ctx.sync()
    t = torch.tensor([worst], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("multi-gpu %s x%d: identical to single GPU over 4 iterations (max diff %g)" % (mode, world, t.item()))
    multi.close(); single.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
