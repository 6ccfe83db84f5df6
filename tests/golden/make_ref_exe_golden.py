"""Regenerates tests/golden/refexe_chain_*.json FROM THE REFERENCE EXECUTABLE ITSELF: oracle/_ref/bpmf_ref_k<K> is the
reference's `bpmf` (every translation unit of its CMake target: bpmf.cpp, sample.cpp, mvnormal.cpp, assign.cpp,
counters.cpp, io.cpp, gzstream.cpp, NO_COMM back end) compiled unmodified from /root/reference against the stand-in
Eigen / Random123 headers of oracle/shim/ (`make -C oracle ref`). It is run on
  * the reference's own data/tiny with the arguments of data/tiny/run_test.sh (-i 9 -b 0 -v -o; K = 10), and
  * the seeded synthetic problem of chain_synth_k32.json (-i 8 -b 3 -v -o; K = 32),
and its -v dumps (U-i.ddm, V-i.ddm), its log lines and its "Final Avg RMSE" are summarised in the schema of
chain_*.json (which make_chain_golden.py produces from the ORACLE). tests/test_oracle_vs_reference.py requires the two
families of fixtures to agree. Needs /root/reference; run from the repository root:
    python tests/golden/make_ref_exe_golden.py
"""
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
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import util  # noqa: E402
from make_chain_golden import CASES, problem  # noqa: E402


def write_mtx(path, shape, rows, cols, vals):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n" + "%d %d %d\n" % (shape[0], shape[1], len(vals)))
        for r, c, v in zip(rows, cols, vals):
            f.write("%d %d %.17g\n" % (r + 1, c + 1, v))


def read_ddm(path):
    raw = open(path, "rb").read()
    nr, nc = struct.unpack_from("<QQ", raw, 0)
    return np.frombuffer(raw, "<f8", nr * nc, 16).reshape(nc, nr)   # [item, k]


def run_reference_exe(spec, workdir):
    """-> (stdout, output directory) of the reference executable on the problem of `spec`"""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref-core"])
    exe = os.path.join(ROOT, "oracle", "_ref", "bpmf_ref_k%d" % spec["K"])
    if spec["data"] == "tiny" and os.path.exists("/root/reference/data/tiny/train.mtx"):
        train_path, test_path = "/root/reference/data/tiny/train.mtx", "/root/reference/data/tiny/test.mtx"   # the files themselves
    else:
        train, test = problem(spec)
        train_path, test_path = os.path.join(workdir, "train.mtx"), os.path.join(workdir, "test.mtx")
        write_mtx(train_path, *train)
        write_mtx(test_path, *test)
    out = os.path.join(workdir, "out")
    os.makedirs(out, exist_ok=True)
    res = subprocess.run([exe, "-n", train_path, "-p", test_path, "-i", str(spec["nsims"]), "-b", str(spec["burnin"]), "-a", "2.0",
                          "-v", "-o", out + "/"], capture_output=True, text=True, cwd=workdir, check=True)
    return res.stdout, out


def summarise(spec, log, out):
    lines = [l for l in log.splitlines() if " iteration " in l]
    assert len(lines) == spec["nsims"]
    its = []
    for it, line in enumerate(lines):
        V, U = read_ddm(os.path.join(out, "V-%d.ddm" % it)), read_ddm(os.path.join(out, "U-%d.ddm" % it))
        m = re.match(r"0: (Burnin|Sampling) iteration (\d+):\t RMSE: ([-\d.]+)\tavg RMSE: ([-\d.]+)\tFU\(\s*([\d.]+)\)\tFM\(\s*([\d.]+)\)", line)
        its.append({"V_mean": [float(x) for x in V.mean(0)], "U_mean": [float(x) for x in U.mean(0)],
                    "V_norm": float(np.sqrt((V * V).sum())), "U_norm": float(np.sqrt((U * U).sum())),
                    "V_first": [float(x) for x in V[0]], "U_last": [float(x) for x in U[-1]],
                    "log": {"phase": m.group(1), "iter": int(m.group(2)), "rmse": m.group(3), "rmse_avg": m.group(4),
                            "FU": m.group(5), "FM": m.group(6)}})
    final = re.search(r"Final Avg RMSE: ([-\d.e+]+)", log).group(1)
    return {"spec": spec, "produced_by": "oracle/_ref/bpmf_ref_k%d (reference sources + stand-in Eigen / Random123 headers)" % spec["K"],
            "iterations": its, "final_avg_rmse_printed": final}


if __name__ == "__main__":
    for name, spec in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            log, out = run_reference_exe(spec, d)
            res = summarise(spec, log, out)
        json.dump(res, open(os.path.join(HERE, "refexe_chain_%s.json" % name), "w"), indent=1)
        print(name, "Final Avg RMSE (as printed by the reference):", res["final_avg_rmse_printed"])
