"""A/B timing of the K=32 stream-kernel configurations on the bench workload (one GPU):
    python bench_micro/tune_stream.py [cfg ...]          e.g. 2216 3216 3220
For every configuration: max |difference| of one sweep's output against the first configuration from the same state,
then the mean CUDA-event time of the item kernel over a few launches of each side."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bpmf_b200  # noqa: E402
from bpmf_b200 import synthetic  # noqa: E402

MOVIES, USERS = 0, 1


def main():
    cfgs = [int(a) for a in sys.argv[1:] if a.isdigit()] or [2216, 3216, 3220]
    wl = os.environ.get("TUNE_WORKLOAD", "synthA-1Mx1M-100Mnnz-K32")
    ratings, K = synthetic.workload(wl, cache_dir="/dev/shm", verbose=True)
    ctx = bpmf_b200.Context(K, 0)
    for side in (MOVIES, USERS):
        n, n_other, ptr, idx, val = ratings.side(side)
        ctx.load_side(side, n, n_other, ptr, idx, val, ratings.mean_rating)
    # a realistic state: two full iterations with the default kernel
    for _ in range(2):
        ctx.sample(MOVIES, 2.0, bpmf_b200.KERNEL_AUTO)
        ctx.sample(USERS, 2.0, bpmf_b200.KERNEL_AUTO)
    ctx.sync()
    ctx.items_kernel_time()
    ref = None
    nnz = ratings.nnz
    div = int(os.environ.get("TUNE_RANGE_DIV", "1"))      # time only the first 1/div of each side's items (a multi-GPU rank's share)
    if div > 1:
        for side in (MOVIES, USERS):
            ctx.set_range(side, 0, ratings.side(side)[0] // div)
        nnz = nnz // div
    state = [ctx.get_items(MOVIES), ctx.get_items(USERS)]
    for cfg in cfgs:
        ctx.set_items(MOVIES, state[MOVIES])
        ctx.set_items(USERS, state[USERS])
        ctx.set_tuning(cfg)
        ctx.sample_items(MOVIES, 7, 2.0, bpmf_b200.KERNEL_STREAM)

        def sync():          # timing probes compute garbage and may report "Cholesky failed": only their time matters
            try:
                ctx.sync()
            except bpmf_b200.BpmfGpuError as e:
                print("   (cfg %d: %s)" % (cfg, e), flush=True)
        sync()
        out = ctx.get_items(MOVIES)
        if ref is None:
            ref = out
        diff = float(np.abs(out - ref).max())
        ctx.items_kernel_time()
        t0 = time.time()
        for rep in range(4):
            ctx.sample_items(MOVIES, 8 + rep, 2.0, bpmf_b200.KERNEL_STREAM)
            ctx.sample_items(USERS, 8 + rep, 2.0, bpmf_b200.KERNEL_STREAM)
        sync()
        ms, cnt = ctx.items_kernel_time()
        per = ms / cnt
        print("cfg %d: %.3f ms/launch  %.1f GB/s  (%.1f%% of 6459 GB/s)  max|diff vs first| %.3e  wall %.2fs"
              % (cfg, per, nnz * K * 8 / per / 1e6, 100 * nnz * K * 8 / per / 1e6 / 6459.3, diff, time.time() - t0), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
