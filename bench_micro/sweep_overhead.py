"""What a sweep costs besides its item kernel (one GPU): movies.sample + users.sample pairs through bpmf_gpu_sample on the
whole range and on the first 1/DIV of each side (a rank's share at DIV GPUs, without peers), CUDA-event time per step
against the sum of the item kernels' own event times.   python bench_micro/sweep_overhead.py [DIV ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bpmf_b200  # noqa: E402
from bpmf_b200 import capi, synthetic  # noqa: E402

MOVIES, USERS = 0, 1


def main():
    divs = [int(a) for a in sys.argv[1:]] or [1, 8]
    ratings, K = synthetic.workload(os.environ.get("TUNE_WORKLOAD", "synthA-1Mx1M-100Mnnz-K32"), cache_dir="/dev/shm", verbose=True)
    ctx = bpmf_b200.Context(K, 0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for side in (MOVIES, USERS):
        n, n_other, ptr, idx, val = ratings.side(side)
        ctx.load_side(side, n, n_other, ptr, idx, val, ratings.mean_rating)
    for div in divs:
        for side in (MOVIES, USERS):
            n = ratings.side(side)[0]
            bi = capi.stats_block_items_for(K, n)
            ctx.set_range(side, 0, n if div == 1 else max(bi, (n // div // bi) * bi))
        for _ in range(3):
            ctx.sample(MOVIES); ctx.sample(USERS)
        ctx.sync(); ctx.items_kernel_time()
        steps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ctx.sample(MOVIES); ctx.sample(USERS)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        kms, cnt = ctx.items_kernel_time()
        print("range 1/%d: %.3f ms per step, item kernels %.3f ms (%d launches) -> %.1f us per sweep besides the kernel"
              % (div, ms, 2 * kms / cnt, cnt, 1e3 * (ms - 2 * kms / cnt) / 2), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
