"""THE DROP-IN, EXECUTED: the reference's own executable — its main loop (c++/bpmf.cpp), its host predict and reductions
(c++/sample.cpp), its assignment and file formats, all compiled unmodified — with oracle/cuda_comm/cuda_comm.h (the
binding of INTEGRATION.md §1) as its communication back end, linked against libbpmf_b200.so
(oracle/_ref/bpmf_ref_cuda_k<K>, built by `make -C oracle ref` where /root/reference exists; travels prebuilt). Its runs
must reproduce what the reference's NO_COMM executable produced on the same inputs (tests/golden/refexe_chain_*.json):
every latent entry summary to 1e-10, the log's RMSE columns and "Final Avg RMSE" as printed."""
import json
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _read_ddm(path):
    raw = open(path, "rb").read()
    nr, nc = struct.unpack_from("<QQ", raw, 0)
    return np.frombuffer(raw, "<f8", nr * nc, 16).reshape(nc, nr)   # [item, k]


def _write_mtx(path, shape, rows, cols, vals):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n" + "%d %d %d\n" % (shape[0], shape[1], len(vals)))
        for r, c, v in zip(rows, cols, vals):
            f.write("%d %d %.17g\n" % (r + 1, c + 1, v))


@pytest.mark.parametrize("name", ["tiny_k10", "synth_k32"])
def test_reference_executable_with_the_b200_back_end(name, tmp_path):
    gold = json.load(open(os.path.join(GOLD, "refexe_chain_%s.json" % name)))
    spec = gold["spec"]
    exe = os.path.join(ROOT, "oracle", "_ref", "bpmf_ref_cuda_k%d" % spec["K"])
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/bpmf_ref_cuda_k%d was not built (needs the reference sources at build time)" % spec["K"])
    if spec["data"] == "tiny":
        train, test = util.TINY_TRAIN, util.TINY_TEST
    else:
        nr, nc, nnz, seed = spec["data"]
        train, test = util.synth_ratings(nr, nc, nnz, seed, skew=0.3)
    _write_mtx(tmp_path / "train.mtx", *train)
    _write_mtx(tmp_path / "test.mtx", *test)
    out = tmp_path / "out"
    out.mkdir()
    res = subprocess.run([exe, "-n", str(tmp_path / "train.mtx"), "-p", str(tmp_path / "test.mtx"), "-i", str(spec["nsims"]), "-b",
                          str(spec["burnin"]), "-a", "2.0", "-v", "-o", str(out) + "/"], capture_output=True, text=True, cwd=tmp_path,
                         timeout=300)
    assert res.returncode == 0, (res.stdout[-1500:], res.stderr[-1500:])
    lines = [l for l in res.stdout.splitlines() if " iteration " in l]
    assert len(lines) == spec["nsims"]
    for it, (line, g) in enumerate(zip(lines, gold["iterations"])):
        V, U = _read_ddm(out / ("V-%d.ddm" % it)), _read_ddm(out / ("U-%d.ddm" % it))
        np.testing.assert_allclose(V.mean(0), g["V_mean"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(U.mean(0), g["U_mean"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(V[0], g["V_first"], rtol=0, atol=1e-10 * max(1.0, np.abs(g["V_first"]).max()))
        np.testing.assert_allclose(U[-1], g["U_last"], rtol=0, atol=1e-10 * max(1.0, np.abs(g["U_last"]).max()))
        m = re.match(r"0: (Burnin|Sampling) iteration (\d+):\t RMSE: ([-\d.]+)\tavg RMSE: ([-\d.]+)\tFU\(\s*([\d.]+)\)\tFM\(\s*([\d.]+)\)", line)
        assert m, line
        log = g["log"]
        assert m.group(1) == log["phase"] and int(m.group(2)) == log["iter"]
        # both sides print 4 (RMSE) / 2 (FU, FM) decimals of values that agree to ~1e-10: equal, or one unit apart at a rounding edge
        assert abs(float(m.group(3)) - float(log["rmse"])) <= 1.01e-4 and abs(float(m.group(4)) - float(log["rmse_avg"])) <= 1.01e-4
        assert abs(float(m.group(5)) - float(log["FU"])) <= 1.01e-2 and abs(float(m.group(6)) - float(log["FM"])) <= 1.01e-2
    final = float(re.search(r"Final Avg RMSE: ([-\d.e+]+)", res.stdout).group(1))
    printed = float(gold["final_avg_rmse_printed"])
    assert abs(final - printed) <= 1e-5 * max(1.0, abs(printed))
