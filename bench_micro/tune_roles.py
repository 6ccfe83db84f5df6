"""A/B timing of the warp-role configurations of the K=32 stream kernel (items_stream32v5_kernel) against the default
kernel on the bench workload (one GPU):   python bench_micro/tune_roles.py [nw:mask:stages:slots ...]
Every configuration must reproduce the default kernel's output bit for bit (same Gram order, same tail function)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bpmf_b200  # noqa: E402
from bpmf_b200 import synthetic  # noqa: E402

MOVIES, USERS = 0, 1
DEFAULT = ["20:0x0000f:4:8", "20:0x0000f:6:6", "20:0x000ff:2:8", "20:0x000ff:3:6", "20:0x00fff:2:6", "20:0x77777:2:6", "20:0x33333:3:6",
           "20:0x0ffff:2:6", "20:0x11111:4:8", "24:0x00000f:4:8", "24:0x0000ff:3:6", "24:0x777777:2:6", "24:0x000fff:2:6", "16:0x000f:6:8",
           "16:0x00ff:3:8"]


def main():
    cfgs = [a for a in sys.argv[1:] if ":" in a] or DEFAULT
    wl = os.environ.get("TUNE_WORKLOAD", "synthA-1Mx1M-100Mnnz-K32")
    ratings, K = synthetic.workload(wl, cache_dir="/dev/shm", verbose=True)
    ctx = bpmf_b200.Context(K, 0)
    for side in (MOVIES, USERS):
        n, n_other, ptr, idx, val = ratings.side(side)
        ctx.load_side(side, n, n_other, ptr, idx, val, ratings.mean_rating)
    for _ in range(2):
        ctx.sample(MOVIES, 2.0, bpmf_b200.KERNEL_AUTO)
        ctx.sample(USERS, 2.0, bpmf_b200.KERNEL_AUTO)
    ctx.sync()
    ctx.items_kernel_time()
    nnz = ratings.nnz
    state = [ctx.get_items(MOVIES), ctx.get_items(USERS)]
    ref = None
    for cfg in ["default"] + cfgs:
        ctx.set_items(MOVIES, state[MOVIES])
        ctx.set_items(USERS, state[USERS])
        if cfg == "default":
            ctx.set_roles(0, 2, 6, 20)
        else:
            nw, mask, ns, nslot = cfg.split(":")
            ctx.set_roles(int(mask, 16), int(ns), int(nslot), int(nw))
        try:
            ctx.sample_items(MOVIES, 7, 2.0, bpmf_b200.KERNEL_STREAM)
            ctx.sync()
        except bpmf_b200.BpmfGpuError as e:
            print("cfg %s: %s" % (cfg, e), flush=True)
            continue
        out = ctx.get_items(MOVIES)
        if ref is None:
            ref = out
        diff = float(np.abs(out - ref).max())
        ctx.items_kernel_time()
        t0 = time.time()
        for rep in range(3):
            ctx.sample_items(MOVIES, 8 + rep, 2.0, bpmf_b200.KERNEL_STREAM)
            ctx.sample_items(USERS, 8 + rep, 2.0, bpmf_b200.KERNEL_STREAM)
        ctx.sync()
        ms, cnt = ctx.items_kernel_time()
        per = ms / cnt
        print("cfg %-18s %.3f ms/launch  %.1f GB/s  (%.1f%% of 6459 GB/s)  max|diff vs default| %.3e  wall %.2fs"
              % (cfg, per, nnz * K * 8 / per / 1e6, 100 * nnz * K * 8 / per / 1e6 / 6459.3, diff, time.time() - t0), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
