"""Writes a synthetic workload of bpmf_b200/synthetic.py as .sdm files (c++/io.cpp:256-288 layout) and runs the `bpmf`
executable on it: a scale test of the C++ host (loaders, CSC build, Sys / CUDA_Sys) and of the whole reference-style loop.

    python bench_micro/run_cli_synth.py [workload] [iterations] [ngpus]
"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bpmf_b200 import synthetic  # noqa: E402


def write_sdm(path, nrows, ncols, rows, cols, vals):
    with open(path, "wb") as f:
        np.array([nrows, ncols, len(vals)], "<u8").tofile(f)
        (np.asarray(rows, "<u4") + 1).tofile(f)
        (np.asarray(cols, "<u4") + 1).tofile(f)
        np.asarray(vals, "<f8").tofile(f)


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "synthA-1Mx1M-100Mnnz-K32"
    iters = sys.argv[2] if len(sys.argv) > 2 else "6"
    ngpus = sys.argv[3] if len(sys.argv) > 3 else "1"
    r, K = synthetic.workload(wl, cache_dir="/dev/shm", verbose=True)
    d = "/dev/shm/bpmf_cli_" + wl
    os.makedirs(d, exist_ok=True)
    rows = np.repeat(np.arange(r.nrows, dtype=np.uint32), np.diff(r.u_ptr))
    write_sdm(os.path.join(d, "train.sdm"), r.nrows, r.ncols, rows, r.u_idx, r.u_val)
    write_sdm(os.path.join(d, "test.sdm"), r.nrows, r.ncols, r.t_rows, r.t_cols, r.t_vals)
    del rows
    exe = os.path.join(ROOT, "bpmf_b200", "host", "bpmf")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "-s", "bpmf"])
    cmd = [exe, "-n", os.path.join(d, "train.sdm"), "-p", os.path.join(d, "test.sdm"), "-i", iters, "-b", "2", "-d", str(K), "-g", ngpus]
    print(" ".join(cmd), flush=True)
    t0 = time.time()
    rc = subprocess.call(cmd)
    print("bpmf exit code %d, wall %.1f s (load + %s iterations)" % (rc, time.time() - t0, iters))
    return rc


if __name__ == "__main__":
    sys.exit(main())
