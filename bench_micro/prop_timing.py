"""Propagated-posterior sweeps (-m / -l, per-item prior precisions) at K = 32: the PROP instantiation of the stream kernel
against the any-K reference-order kernel, on the small synthetic workload (20K x 20K, 1M ratings).

    python bench_micro/prop_timing.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bpmf_b200
    from bpmf_b200 import synthetic
    r, K = synthetic.workload("small-20Kx20K-1Mnnz-K32")
    ctx = bpmf_b200.Context(K)
    for side in (0, 1):
        n, n_other, ptr, idx, val = r.side(side)
        ctx.load_side(side, n, n_other, ptr, idx, val, r.mean_rating)
    for it in range(2):                       # a couple of ordinary sweeps so that the latents are not zero
        ctx.sample(0); ctx.sample(1)
    rng = np.random.default_rng(3)
    for side in (0, 1):
        n = ctx.num[side]
        a = rng.normal(size=(n, K, 4)) * 0.3
        lam = np.einsum("nik,njk->nij", a, a) + 2.0 * np.eye(K)[None]      # SPD per item
        ctx.set_prop_posterior(side, np.zeros((n, K)), lam.reshape(n, K * K))
    out = {}
    for name, variant in (("stream PROP", bpmf_b200.KERNEL_STREAM), ("any-K exact", bpmf_b200.KERNEL_EXACT)):
        saved = [ctx.get_items(s) for s in (0, 1)]
        ctx.sample_items(0, 5, 2.0, variant); ctx.sync()
        t0 = time.time()
        for rep in range(5):
            ctx.sample_items(0, 6 + rep, 2.0, variant)
        ctx.sync()
        out[name] = (time.time() - t0) / 5 * 1e3
        res = ctx.get_items(0)
        for s in (0, 1):
            ctx.set_items(s, saved[s])
        print("%-12s movies sweep with per-item priors: %.3f ms" % (name, out[name]), flush=True)
        out[name + " items"] = res
    d = np.abs(out["stream PROP items"] - out["any-K exact items"]).max()
    print("max |difference| of the two kernels after the same 6 sweeps: %.3g" % d)
    ctx.close()


if __name__ == "__main__":
    main()
