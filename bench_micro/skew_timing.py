"""Times one movies sweep and one users sweep of a ChEMBL-20-shaped problem (483 500 x 5 775, ~0.8M ratings, one target
with 110 000 ratings, 85 000 empty compound rows) with and without the chunked heavy-item path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bpmf_b200  # noqa: E402
import util  # noqa: E402
import scipy.sparse as sp  # noqa: E402


def main():
    K = 32
    train, _ = util.synth_ratings(483500, 5775, 720000, 2020, rank=6, skew=1.0, empty_rows=85000, heavy_col=110000)
    (shape, rows, cols, vals) = train
    R = sp.coo_matrix((vals, (rows, cols)), shape=shape)
    mean = float(vals.mean())
    print("ratings %d, heaviest movie %d, heaviest user %d" % (len(vals), np.bincount(cols).max(), np.bincount(rows).max()))
    for label, thr in (("chunked (threshold 4096)", 4096), ("chunked (threshold 16384)", 16384), ("plain (threshold off)", 1 << 40)):
        ctx = bpmf_b200.Context(K, 0)
        ctx.set_heavy_threshold(thr)
        for side, M in ((0, R.tocsc()), (1, R.T.tocsc())):
            M.sort_indices()
            ctx.load_side(side, M.shape[1], M.shape[0], M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data, mean)
        for _ in range(3):
            ctx.sample(0, 2.0); ctx.sample(1, 2.0)
        ctx.sync(); ctx.items_kernel_time()
        per = {}
        for side in (0, 1):
            for rep in range(5):
                ctx.sample(side, 2.0)
            ctx.sync()
            ms, n = ctx.items_kernel_time()
            per[side] = ms / n
        t0 = time.time()
        for _ in range(10):
            ctx.sample(0, 2.0); ctx.sample(1, 2.0)
        ctx.sync()
        print("%-28s item kernels: movies %.3f ms, users %.3f ms; whole iteration %.3f ms" % (label, per[0], per[1], (time.time() - t0) * 100))
        ctx.close()


if __name__ == "__main__":
    main()
