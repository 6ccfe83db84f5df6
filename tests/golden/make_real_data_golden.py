"""Fixtures on the REFERENCE'S OWN DATASETS, produced by the REFERENCE EXECUTABLE (oracle/_ref/bpmf_ref_k32: every
translation unit of its CMake target compiled unmodified against the stand-in Eigen / Random123 headers, `make -C oracle
ref-core`):

  * data/movielens/ml-train.mtx + ml-test.mtx (ML-100K, 1682 x 943 after the reference's transpose convention: 943 rows =
    users, 1682 columns = movies; what its only ctests run, CMakeLists.txt:174-182), -i 20 -b 5 -a 2.0 -v -o
  * data/chembl_20/train.mtx + test.mtx (483 500 x 5 775, 818 931 ratings: one column with 110 118 ratings, 84 815 empty
    rows, test columns that have no training rating; BASELINE.json configs[2]), -i 12 -b 4 -a 2.0 -v -o

/root/reference does not exist on the GPU box, so the INPUTS travel as tests/golden/data/*.sdm.gz (the reference's binary
sparse format, c++/io.cpp:256-288, entries in FILE order so that every loader sees the same triplet sequence) and the
reference's results as tests/golden/refexe_real_*.json:
  per iteration: column means of U and V, norms, rows of interest (first/last, the hottest column, empty rows), a fixed
  pseudo-random projection of every latent matrix (sum_i w_i x_i: any single item off by 1e-9 shows), the log fields as
  printed; after the run: "Final Avg RMSE" as printed, digests of Pavg.sdm / Pm2.sdm, and U-mu / V-mu column means.
Run from the repository root (minutes of CPU: ChEMBL is ~13 s per iteration on one core):
    python tests/golden/make_real_data_golden.py [ml100k] [chembl20]
"""
import gzip
import json
import os
import re
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DATA = os.path.join(HERE, "data")
REF_DATA = "/root/reference/data"

CASES = {
    "ml100k": dict(K=32, train=REF_DATA + "/movielens/ml-train.mtx", test=REF_DATA + "/movielens/ml-test.mtx", nsims=20, burnin=5),
    "chembl20": dict(K=32, train=REF_DATA + "/chembl_20/train.mtx", test=REF_DATA + "/chembl_20/test.mtx", nsims=12, burnin=4),
}


def read_mtx_coordinate(path):
    """(nrow, ncol), rows, cols, vals in FILE order, 0-based; values parsed with correct rounding like strtod"""
    import pandas as pd
    with open(path) as f:
        skip = 0
        line = f.readline()
        assert line.lower().startswith("%%matrixmarket matrix coordinate real general"), line
        skip += 1
        line = f.readline()
        while line.startswith("%") or not line.strip():
            skip += 1
            line = f.readline()
        nr, nc, nnz = (int(x) for x in line.split())
        skip += 1
    df = pd.read_csv(path, sep=r"\s+", skiprows=skip, header=None, comment="%", float_precision="round_trip",
                     names=["r", "c", "v"], dtype={"r": np.int64, "c": np.int64, "v": np.float64})
    assert len(df) == nnz, (len(df), nnz)
    return (nr, nc), (df.r.values - 1).astype(np.int32), (df.c.values - 1).astype(np.int32), df.v.values.astype(np.float64)


def write_sdm_gz(path, shape, rows, cols, vals):
    """c++/io.cpp:256-288: u64 nrow, ncol, nnz; u32 rows[nnz] (1-based); u32 cols[nnz] (1-based); f64 vals[nnz]"""
    raw = struct.pack("<QQQ", shape[0], shape[1], len(vals)) + (rows.astype("<u4") + 1).tobytes() + (cols.astype("<u4") + 1).tobytes() \
        + vals.astype("<f8").tobytes()
    with open(path, "wb") as f:
        with gzip.GzipFile(fileobj=f, mode="wb", compresslevel=9, mtime=0, filename="") as g:   # reproducible bytes
            g.write(raw)


def read_sdm(path):
    raw = (gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")).read()
    nr, nc, n = struct.unpack_from("<QQQ", raw, 0)
    rows = np.frombuffer(raw, "<u4", n, 24).astype(np.int32) - 1
    cols = np.frombuffer(raw, "<u4", n, 24 + 4 * n).astype(np.int32) - 1
    vals = np.frombuffer(raw, "<f8", n, 24 + 8 * n).copy()
    return (nr, nc), rows, cols, vals


def read_ddm(path):
    raw = open(path, "rb").read()
    nr, nc = struct.unpack_from("<QQ", raw, 0)
    return np.frombuffer(raw, "<f8", nr * nc, 16).reshape(nc, nr)   # [item, k]


def proj_weights(n):
    """fixed, seed-free weights of the latent-matrix projection"""
    i = np.arange(n, dtype=np.float64)
    return np.cos(0.37 * i + 0.11) + 0.5 * np.sin(0.013 * i)


def rows_of_interest(train):
    """item indices worth pinning entry by entry: first, last, the heaviest, the lightest non-empty and an empty one per side"""
    (nr, nc), rows, cols, _ = train
    out = {}
    for name, idx, n in (("V", cols, nc), ("U", rows, nr)):
        cnt = np.bincount(idx, minlength=n)
        pick = [0, n - 1, int(cnt.argmax())]
        nz = np.flatnonzero(cnt > 0)
        pick.append(int(nz[cnt[nz].argmin()]))
        empty = np.flatnonzero(cnt == 0)
        if len(empty):
            pick += [int(empty[0]), int(empty[len(empty) // 2])]
        out[name] = sorted(set(pick))
    return out


def summarise(name, spec, train, log, out):
    lines = [l for l in log.splitlines() if " iteration " in l]
    assert len(lines) == spec["nsims"], log[-2000:]
    roi = rows_of_interest(train)
    (nr, nc) = train[0]
    wV, wU = proj_weights(nc), proj_weights(nr)
    its = []
    for it, line in enumerate(lines):
        V, U = read_ddm(os.path.join(out, "V-%d.ddm" % it)), read_ddm(os.path.join(out, "U-%d.ddm" % it))
        assert V.shape == (nc, spec["K"]) and U.shape == (nr, spec["K"])
        m = re.match(r"0: (Burnin|Sampling) iteration (\d+):\t RMSE: ([-\d.]+)\tavg RMSE: ([-\d.]+)\tFU\(\s*([\d.]+)\)\tFM\(\s*([\d.]+)\)", line)
        its.append({"V_mean": [float(x) for x in V.mean(0)], "U_mean": [float(x) for x in U.mean(0)],
                    "V_norm": float(np.sqrt((V * V).sum())), "U_norm": float(np.sqrt((U * U).sum())),
                    "V_absmax": float(np.abs(V).max()), "U_absmax": float(np.abs(U).max()),
                    "V_proj": [float(x) for x in wV @ V], "U_proj": [float(x) for x in wU @ U],
                    "V_rows": {str(i): [float(x) for x in V[i]] for i in roi["V"]},
                    "U_rows": {str(i): [float(x) for x in U[i]] for i in roi["U"]},
                    "log": {"phase": m.group(1), "iter": int(m.group(2)), "rmse": m.group(3), "rmse_avg": m.group(4),
                            "FU": m.group(5), "FM": m.group(6)}})
    final = re.search(r"Final Avg RMSE: ([-\d.e+]+)", log).group(1)
    res = {"spec": {k: (os.path.relpath(v, "/root/reference") if isinstance(v, str) else v) for k, v in spec.items()},
           "inputs": {"train": "data/%s_train.sdm.gz" % name, "test": "data/%s_test.sdm.gz" % name},
           "produced_by": "oracle/_ref/bpmf_ref_k%d (reference sources + stand-in Eigen / Random123 headers) on the reference's own files" % spec["K"],
           "iterations": its, "final_avg_rmse_printed": final}
    # -o outputs (c++/bpmf.cpp:221-240): running means of the predictions, posterior means of the latent vectors
    for tag in ("Pavg", "Pm2"):
        (_, _), r, c, v = read_sdm(os.path.join(out, tag + ".sdm"))
        w = proj_weights(len(v))
        step = max(1, len(v) // 64)
        res[tag] = {"n": int(len(v)), "sum": float(v.sum()), "proj": float(w @ v), "absmax": float(np.abs(v).max()),
                    "sample_every": step, "sample": [float(x) for x in v[::step]],
                    "sample_rows": [int(x) for x in r[::step]], "sample_cols": [int(x) for x in c[::step]]}
    for tag, w in (("U-mu", wU), ("V-mu", wV)):
        M = read_ddm(os.path.join(out, tag + ".ddm"))
        res[tag] = {"mean": [float(x) for x in M.mean(0)], "proj": [float(x) for x in w @ M], "absmax": float(np.abs(M).max())}
    # structure facts the tests assert about the inputs (SURVEY §8 sizes table)
    cnt_c, cnt_r = np.bincount(train[2], minlength=nc), np.bincount(train[1], minlength=nr)
    res["structure"] = {"shape": [int(nr), int(nc)], "nnz": int(len(train[3])), "max_col_nnz": int(cnt_c.max()),
                        "max_row_nnz": int(cnt_r.max()), "empty_rows": int((cnt_r == 0).sum()), "empty_cols": int((cnt_c == 0).sum())}
    return res


def main(which):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref-core"])
    os.makedirs(DATA, exist_ok=True)
    for name in which:
        spec = CASES[name]
        train, test = read_mtx_coordinate(spec["train"]), read_mtx_coordinate(spec["test"])
        write_sdm_gz(os.path.join(DATA, name + "_train.sdm.gz"), *train)
        write_sdm_gz(os.path.join(DATA, name + "_test.sdm.gz"), *test)
        exe = os.path.join(ROOT, "oracle", "_ref", "bpmf_ref_k%d" % spec["K"])
        with tempfile.TemporaryDirectory() as d:
            out = os.path.join(d, "out")
            os.makedirs(out)
            # the reference reads ITS OWN text files here; the .sdm.gz copies are what the GPU box gets
            res = subprocess.run([exe, "-n", spec["train"], "-p", spec["test"], "-i", str(spec["nsims"]), "-b", str(spec["burnin"]),
                                  "-a", "2.0", "-v", "-o", out + "/"], capture_output=True, text=True, cwd=d, check=True)
            summary = summarise(name, spec, train, res.stdout, out)
        json.dump(summary, open(os.path.join(HERE, "refexe_real_%s_k%d.json" % (name, spec["K"])), "w"), indent=1)
        print(name, "Final Avg RMSE (as printed by the reference):", summary["final_avg_rmse_printed"], summary["structure"])


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
