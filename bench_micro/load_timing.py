"""Load path: both factors' compressed matrices from one coordinate list, built on the device
(bpmf_gpu_load_coo: upload + radix sort + emit, both orientations) against the host build (scipy COO -> CSC and CSR ->
bpmf_gpu_load_side twice). The coordinate list is shuffled like a file in arbitrary order.

    python bench_micro/load_timing.py [workload]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import scipy.sparse as sp
    import bpmf_b200
    from bpmf_b200 import synthetic
    name = sys.argv[1] if len(sys.argv) > 1 else "synthA-1Mx1M-100Mnnz-K32"
    r, K = synthetic.workload(name)
    rows = np.repeat(np.arange(r.nrows, dtype=np.int32), np.diff(r.u_ptr))
    cols, vals = np.asarray(r.u_idx), np.asarray(r.u_val)
    perm = np.random.default_rng(1).permutation(len(vals))
    rows, cols, vals = rows[perm], cols[perm], vals[perm]
    print("%s: %d x %d, %d entries in shuffled order" % (name, r.nrows, r.ncols, len(vals)), flush=True)
    ctx = bpmf_b200.Context(K)
    for rep in range(2):
        t0 = time.time()
        ctx.load_coo(r.nrows, r.ncols, rows, cols, vals)
        ctx.sync()
        t_dev = time.time() - t0
        print("device build (bpmf_gpu_load_coo, both sides, incl. upload and the two host-side mean sums): %.2f s" % t_dev, flush=True)
    t0 = time.time()
    coo = sp.coo_matrix((vals, (rows, cols)), shape=(r.nrows, r.ncols))
    csc = coo.tocsc(); csc.sort_indices()
    csr = coo.tocsr(); csr.sort_indices()
    t_host_build = time.time() - t0
    host = bpmf_b200.Context(K)
    t0 = time.time()
    host.load_side(0, r.ncols, r.nrows, csc.indptr.astype(np.int64), csc.indices, csc.data, float(np.cumsum(csc.data)[-1] / csc.nnz))
    host.load_side(1, r.nrows, r.ncols, csr.indptr.astype(np.int64), csr.indices, csr.data, float(np.cumsum(csr.data)[-1] / csr.nnz))
    host.sync()
    t_host_load = time.time() - t0
    print("host build (scipy tocsc + tocsr, sorted): %.2f s; bpmf_gpu_load_side x 2: %.2f s" % (t_host_build, t_host_load), flush=True)
    for side in (0, 1):
        a, b = ctx.get_side(side), host.get_side(side)
        same = a[0] == b[0] and a[1] == b[1] and all(np.array_equal(x, y) for x, y in zip(a[2:], b[2:]))
        print("side %d: identical arrays and mean: %s" % (side, same), flush=True)
    ctx.close(); host.close()


if __name__ == "__main__":
    main()
