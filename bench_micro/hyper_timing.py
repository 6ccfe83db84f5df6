"""Duration of the single-block hyper-parameter draw (hyper_kernel) for K = 16 ... 128, CUDA events around
bpmf_gpu_sample_hyper, and its results' digest (the A/B must agree bit for bit):

    python bench_micro/hyper_timing.py                              # work arrays in shared memory where they fit
    BPMF_HYPER_GLOBAL_SCRATCH=1 python bench_micro/hyper_timing.py  # everything in global scratch (rounds 1-2a)
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bpmf_b200  # noqa: E402
import util  # noqa: E402


def main():
    mode = "global scratch" if os.environ.get("BPMF_HYPER_GLOBAL_SCRATCH") else "shared memory where it fits"
    for K in (10, 16, 32, 48, 64, 128):
        N = 50000
        cov = util.random_spd(K, 100 + K, scale=0.3)
        ctx = bpmf_b200.Context(K, 0)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.load_side(0, N, 8, np.zeros(N + 1, np.int64), np.zeros(0, np.int32), np.zeros(0), 0.0)
        c = cov.T.copy().reshape(-1)
        for it in range(3):
            ctx.sample_hyper(0, it, None, c)
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(reps):
            ctx.sample_hyper(0, 7, None, c)
        e1.record()
        torch.cuda.synchronize()
        mu, LU, LF = ctx.get_hyper(0)
        h = hashlib.sha1(np.ascontiguousarray(mu).tobytes() + np.ascontiguousarray(LU).tobytes() + np.ascontiguousarray(LF).tobytes()).hexdigest()[:12]
        # (each call also stages cov through pinned memory: ~10 us of copy in front of the kernel)
        print("K = %3d: %.1f us per draw incl. the upload of cov (%s), sha1 %s" % (K, 1e3 * e0.elapsed_time(e1) / reps, mode, h), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
