"""Device-resident U+V step of a multi-GPU run (push exchange), CUDA events, max over ranks, and a SHA-1 of rank 0's latent
matrices after the run (the chain must not depend on where the reductions run):

    [BPMF_STATS_MAIN=1 | BPMF_RESERVE_SMS=n] python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29547 bench_micro/multi_step_timing.py
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bpmf_b200  # noqa: E402
from bpmf_b200 import synthetic  # noqa: E402
from bpmf_b200.sampler import GibbsSampler, MOVIES, USERS  # noqa: E402


def main():
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = os.environ.get("TUNE_WORKLOAD", "synthA-1Mx1M-100Mnnz-K32")
    if rank == 0:
        ratings, K = synthetic.workload(wl, cache_dir="/dev/shm", verbose=False)
    dist.barrier(device_ids=[local_rank])
    if rank != 0:
        ratings, K = synthetic.workload(wl, cache_dir="/dev/shm")
    gs = GibbsSampler(ratings, K, device=local_rank, alpha=2.0, variant=bpmf_b200.KERNEL_AUTO, exchange="push", with_test=True)
    for _ in range(3):
        gs.step()
    gs.ctx.sync()
    gs.ctx.items_kernel_time()
    steps = int(os.environ.get("STEPS", "20"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(device_ids=[local_rank]); torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        gs.step()
    e1.record()
    dist.barrier(device_ids=[local_rank]); torch.cuda.synchronize()
    gs.ctx.sync()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    kms, cnt = gs.ctx.items_kernel_time()
    k = torch.tensor([kms / max(1, cnt)], dtype=torch.float64, device="cuda")
    dist.all_reduce(k, op=dist.ReduceOp.MAX)
    rm = gs.predict(burnin=3)[0]
    h = hashlib.sha1()
    for side in (MOVIES, USERS):
        h.update(np.ascontiguousarray(gs.items_host(side)).tobytes())
    if rank == 0:
        mode = "reductions on the main stream" if os.environ.get("BPMF_STATS_MAIN") else "reductions on the auxiliary stream, %s SMs reserved" % os.environ.get("BPMF_RESERVE_SMS", "default")
        print("%d GPUs, %s: %.3f ms per step, slowest rank's item kernel %.3f ms per sweep -> %.1f us per sweep besides it; rmse %.10f sha1 %s"
              % (world, mode, t.item(), k.item(), 1e3 * (t.item() - 2 * k.item()) / 2, rm[0], h.hexdigest()[:16]), flush=True)
    gs.close()
    dist.barrier(device_ids=[local_rank])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
