"""Time per Gibbs iteration on the reference's own datasets (tests/golden/data/*.sdm.gz: ML-100K and ChEMBL-20, K = 32), one
GPU, through the C ABI: movies sweep + users sweep (+ both predicts), CUDA events; ChEMBL also with the chunked heavy-item
path switched off (its hottest target has 110 118 ratings).   python bench_micro/real_data_timing.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bpmf_b200  # noqa: E402
import test_real_data as rd  # noqa: E402

MOVIES, USERS = 0, 1


def run(name, heavy_threshold=None, iters=20):
    gold = rd.fixture(name)
    (shape, rows, cols, vals), (_, trows, tcols, tvals) = rd.inputs(gold)
    ctx = bpmf_b200.Context(32, 0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    if heavy_threshold:
        ctx.set_heavy_threshold(heavy_threshold)
    ctx.load_coo(shape[0], shape[1], rows, cols, vals)
    ctx.load_test_coo(trows, tcols, tvals)
    for _ in range(3):
        ctx.sample(MOVIES); ctx.sample(USERS)
    ctx.sync(); ctx.items_kernel_time()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for _ in range(iters):
        ctx.sample(MOVIES)
    ev[1].record()
    for _ in range(iters):
        ctx.sample(USERS)
    ev[2].record()
    for _ in range(iters):
        ctx.sample(MOVIES); ctx.sample(USERS); ctx.predict(MOVIES, 1); ctx.predict(USERS, 1)
    ev[3].record()
    torch.cuda.synchronize()
    m, u, full = ev[0].elapsed_time(ev[1]) / iters, ev[1].elapsed_time(ev[2]) / iters, ev[2].elapsed_time(ev[3]) / iters
    n = shape[0] + shape[1]
    print("%-9s %s: movies sweep %.3f ms, users sweep %.3f ms -> %.2f M samples/s; whole iteration incl. predict %.3f ms"
          % (name, "heavy items chunked (default)" if not heavy_threshold else "chunking off", m, u, n / (m + u) / 1e3, full), flush=True)
    ctx.close()


if __name__ == "__main__":
    run("ml100k")
    run("chembl20")
    run("chembl20", heavy_threshold=1 << 40, iters=5)
