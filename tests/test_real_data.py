"""Parity on the REFERENCE'S OWN DATASETS (BASELINE.json configs[1..2]; CMakeLists.txt:174-182 runs ML-100K):
  * data/movielens ML-100K (943 x 1682, 80 000 train / 20 000 test), K=32, -i 20 -b 5 -a 2
  * data/chembl_20 (483 500 x 5 775, 818 931 train / 204 772 test; one column of 110 118 ratings, 84 815 empty rows, test
    columns without a single training rating), K=32, -i 12 -b 4 -a 2
against tests/golden/refexe_real_*.json, which the REFERENCE EXECUTABLE produced from the reference's own files
(tests/golden/make_real_data_golden.py). The inputs travel as tests/golden/data/*.sdm.gz (c++/io.cpp:256-288 layout,
entries in file order).

CPU (-m "not gpu"): the inputs have the structure SURVEY §8 lists; the oracle reproduces the fixtures (ML-100K: the whole
chain bit for bit; ChEMBL: its first iterations).
GPU (-m gpu): the chain through the C ABI (bpmf_gpu_load_coo / _sample / _predict), the `bpmf` executable on the
.sdm.gz files and on a .mtx.gz of them (the reference's bpmf_compressed ctest), and the reference's own executable with the
B200 back end — column means 1e-10, every pinned latent row 1e-10, RMSE 1e-6 / as printed, "Final Avg RMSE" as printed,
the -o outputs (Pavg / Pm2 / U-mu / V-mu); on ChEMBL the chunked heavy-item path must have engaged.
"""
import gzip
import json
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import util
from util import MOVIES, USERS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
HOST = os.path.join(ROOT, "bpmf_b200", "host")
TOL = 1e-10
TOL_RMSE = 1e-6
LINE = re.compile(r"0: (Burnin|Sampling) iteration (\d+):\t RMSE: ([-\d.]+)\tavg RMSE: ([-\d.]+)\tFU\(\s*([\d.]+)\)\tFM\(\s*([\d.]+)\)")


def fixture(name):
    return json.load(open(os.path.join(GOLD, "refexe_real_%s_k32.json" % name)))


def read_sdm(path):
    raw = (gzip.open(path, "rb") if str(path).endswith(".gz") else open(path, "rb")).read()
    nr, nc, n = struct.unpack_from("<QQQ", raw, 0)
    rows = np.frombuffer(raw, "<u4", n, 24).astype(np.int32) - 1
    cols = np.frombuffer(raw, "<u4", n, 24 + 4 * n).astype(np.int32) - 1
    vals = np.frombuffer(raw, "<f8", n, 24 + 8 * n).copy()
    return (int(nr), int(nc)), rows, cols, vals


def read_ddm(path):
    raw = open(path, "rb").read()
    nr, nc = struct.unpack_from("<QQ", raw, 0)
    return np.frombuffer(raw, "<f8", nr * nc, 16).reshape(nc, nr)   # [item, k]


def proj_weights(n):
    i = np.arange(n, dtype=np.float64)
    return np.cos(0.37 * i + 0.11) + 0.5 * np.sin(0.013 * i)


def inputs(gold):
    """train / test triplets; the test matrix is resized to the train shape like c++/sample.cpp:119-122"""
    train = read_sdm(os.path.join(GOLD, gold["inputs"]["train"]))
    test = read_sdm(os.path.join(GOLD, gold["inputs"]["test"]))
    shape = (max(train[0][0], test[0][0]), max(train[0][1], test[0][1]))
    return (shape,) + train[1:], (shape,) + test[1:]


def check_iteration(g, V, U, tol=TOL):
    """one iteration's latent matrices against the reference's: column means (the north-star gate), pinned rows, norms,
    and a projection that sees every item"""
    np.testing.assert_allclose(V.mean(0), g["V_mean"], rtol=0, atol=tol)
    np.testing.assert_allclose(U.mean(0), g["U_mean"], rtol=0, atol=tol)
    for M, rows, amax in ((V, g["V_rows"], g["V_absmax"]), (U, g["U_rows"], g["U_absmax"])):
        for i, ref in rows.items():
            # relative to the matrix' own scale: on ChEMBL the chain sits on the hyper-prior plateau (entries ~1e-3)
            np.testing.assert_allclose(M[int(i)], ref, rtol=0, atol=tol * min(1.0, max(amax, 1e-3)) if tol else 0)
    assert abs(np.sqrt((V * V).sum()) - g["V_norm"]) <= max(tol, 1e-12) * max(1.0, g["V_norm"])
    assert abs(np.sqrt((U * U).sum()) - g["U_norm"]) <= max(tol, 1e-12) * max(1.0, g["U_norm"])
    for M, ref, amax in ((V, g["V_proj"], g["V_absmax"]), (U, g["U_proj"], g["U_absmax"])):
        got = proj_weights(M.shape[0]) @ M
        # |sum_i w_i e_i| <= sqrt(N) * 1.5 * max|e_i| for independent errors; a single item off by 1e-9 * scale shows
        np.testing.assert_allclose(got, ref, rtol=0, atol=max(tol, 1e-13) * amax * 1.5 * np.sqrt(M.shape[0]) + 1e-12 * np.abs(ref).max())


def check_log_line(line, g):
    m = LINE.match(line)
    assert m, line
    log = g["log"]
    assert m.group(1) == log["phase"] and int(m.group(2)) == log["iter"]
    # both print 4 (RMSE) / 2 (FU, FM) decimals of values that agree to ~1e-10: equal, or one unit apart at a rounding edge
    assert abs(float(m.group(3)) - float(log["rmse"])) <= 1.01e-4 and abs(float(m.group(4)) - float(log["rmse_avg"])) <= 1.01e-4
    assert abs(float(m.group(5)) - float(log["FU"])) <= 1.01e-2 and abs(float(m.group(6)) - float(log["FM"])) <= 1.01e-2


def check_outputs(gold, out):
    """-o files (c++/bpmf.cpp:221-240) against the reference's"""
    for tag in ("Pavg", "Pm2"):
        g = gold[tag]
        _, r, c, v = read_sdm(os.path.join(out, tag + ".sdm"))
        assert len(v) == g["n"]
        step = g["sample_every"]
        assert [int(x) for x in r[::step]] == g["sample_rows"] and [int(x) for x in c[::step]] == g["sample_cols"]
        scale = max(1.0, g["absmax"])
        np.testing.assert_allclose(v[::step], g["sample"], rtol=0, atol=1e-8 * scale)
        assert abs(v.sum() - g["sum"]) <= 1e-8 * scale * np.sqrt(len(v)) + 1e-12 * abs(g["sum"])
        assert abs(proj_weights(len(v)) @ v - g["proj"]) <= 1e-8 * scale * np.sqrt(len(v))
    for tag in ("U-mu", "V-mu"):
        g = gold[tag]
        M = read_ddm(os.path.join(out, tag + ".ddm"))
        np.testing.assert_allclose(M.mean(0), g["mean"], rtol=0, atol=TOL)
        np.testing.assert_allclose(proj_weights(M.shape[0]) @ M, g["proj"], rtol=0, atol=TOL * max(1e-3, g["absmax"]) * 1.5 * np.sqrt(M.shape[0]) + 1e-12)


# --------------------------------------------------------------------------------------------------- CPU

@pytest.mark.parametrize("name", ["ml100k", "chembl20"])
def test_inputs_have_the_reference_structure(name):
    gold = fixture(name)
    train, test = inputs(gold)
    (nr, nc), rows, cols, vals = train
    st = gold["structure"]
    assert [nr, nc] == st["shape"] and len(vals) == st["nnz"]
    cnt_c, cnt_r = np.bincount(cols, minlength=nc), np.bincount(rows, minlength=nr)
    assert cnt_c.max() == st["max_col_nnz"] and cnt_r.max() == st["max_row_nnz"]
    assert (cnt_r == 0).sum() == st["empty_rows"] and (cnt_c == 0).sum() == st["empty_cols"]
    if name == "chembl20":       # SURVEY §8 sizes table
        assert (nr, nc, len(vals), len(test[3])) == (483500, 5775, 818931, 204772)
        assert cnt_c.max() == 110118 and (cnt_r == 0).sum() == 84815
        cold = cnt_c[test[2]] == 0                            # cold-start predictions: 204 527 of the 204 772 test entries
        assert cold.sum() == 204527 and (cnt_c == 0).sum() == 742   # sit in columns without a single training rating
    else:
        assert (nr, nc, len(vals), len(test[3])) == (943, 1682, 80000, 20000)


@pytest.mark.parametrize("name,iters", [("ml100k", 20), ("chembl20", 3)])
def test_oracle_reproduces_the_reference_on_its_own_data(name, iters):
    """the oracle port against the reference executable's run: bit for bit (same arithmetic in the same order; one thread,
    like the executable that made the fixtures — the sweep reductions are per thread, c++/sample.cpp:345-347,379-381, so
    the reference itself is reproducible only to round-off across thread counts)"""
    gold = fixture(name)
    spec = gold["spec"]
    train, test = inputs(gold)
    orc = util.make_oracle(spec["K"], train, test, alpha=2.0, burnin=spec["burnin"], nthreads=1)
    for it in range(iters):
        orc.iterate()
        g = gold["iterations"][it]
        check_iteration(g, orc.items(MOVIES), orc.items(USERS), tol=0)
        r = orc.rmse(MOVIES)
        assert "%3.4f" % r[0] == g["log"]["rmse"] and "%3.4f" % r[1] == g["log"]["rmse_avg"]
        assert "%6.2f" % np.sqrt(orc.stats(USERS)[3]) == "%6s" % g["log"]["FU"]
    if iters == spec["nsims"]:
        orc.finish()
        assert "%g" % orc.rmse(MOVIES)[1] == gold["final_avg_rmse_printed"]
        pa, pm = orc.pred(MOVIES)
        assert abs(pa.sum() - gold["Pavg"]["sum"]) <= 1e-9 * len(pa)


# --------------------------------------------------------------------------------------------------- GPU

@pytest.fixture(scope="module")
def gpu():
    import bpmf_b200
    bpmf_b200.load_library()
    return bpmf_b200


@pytest.fixture(scope="module")
def exe():
    path = os.path.join(HOST, "bpmf")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", HOST, "-s", "bpmf"])
    return path


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ml100k", "chembl20"])
def test_chain_through_the_c_abi(gpu, name):
    gold = fixture(name)
    spec = gold["spec"]
    K, burnin = spec["K"], spec["burnin"]
    (shape, rows, cols, vals), (_, trows, tcols, tvals) = inputs(gold)
    ctx = gpu.Context(K)
    ctx.load_coo(shape[0], shape[1], rows, cols, vals)          # device build of both sides' matrices (§8f N4)
    ctx.load_test_coo(trows, tcols, tvals)
    plain = None
    if name == "chembl20":      # the same chain with the chunked heavy-item path switched off, to see that it engaged
        plain = gpu.Context(K)
        plain.set_heavy_threshold(1 << 40)
        plain.load_coo(shape[0], shape[1], rows, cols, vals)
    engaged = []
    for it, g in enumerate(gold["iterations"]):
        before = ctx.launch_count()
        ctx.sample(MOVIES, 2.0, gpu.KERNEL_AUTO)
        engaged.append(ctx.launch_count() - before)
        ctx.sample(USERS, 2.0, gpu.KERNEL_AUTO)
        rm = ctx.predict(MOVIES, burnin)
        ctx.predict(USERS, burnin)
        V, U = ctx.get_items(MOVIES), ctx.get_items(USERS)
        check_iteration(g, V, U)
        assert rm[2] == len(tvals)
        assert abs(rm[0] - float(g["log"]["rmse"])) <= 5e-5 + TOL_RMSE and abs(rm[1] - float(g["log"]["rmse_avg"])) <= 5e-5 + TOL_RMSE
        assert abs(np.sqrt(ctx.get_stats(USERS)[3]) - float(g["log"]["FU"])) <= 5.01e-3
        if plain is not None and it < 2:
            before = plain.launch_count()
            plain.sample(MOVIES, 2.0, gpu.KERNEL_AUTO)
            # the chunked path adds two kernels per movies sweep: partial Grams of the chunks + the heavy items' tails
            assert engaged[it] == plain.launch_count() - before + 2
            plain.sample(USERS, 2.0, gpu.KERNEL_AUTO)
            check_iteration(g, plain.get_items(MOVIES), plain.get_items(USERS))
    final = ctx.predict(MOVIES, burnin)      # the extra predict of c++/bpmf.cpp:225|242
    printed = float(gold["final_avg_rmse_printed"])
    assert abs(final[1] - printed) <= 1e-5 * max(1.0, abs(printed))
    pa, pm = ctx.get_predictions(MOVIES)
    assert abs(pa.sum() - gold["Pavg"]["sum"]) <= 1e-8 * np.sqrt(len(pa)) * max(1.0, gold["Pavg"]["absmax"])
    if plain is not None:
        plain.close()
    ctx.close()


def _run_exe(exe_path, train, test, spec, out, extra=(), cwd=None, timeout=900):
    cmd = [exe_path, "-n", str(train), "-p", str(test), "-i", str(spec["nsims"]), "-b", str(spec["burnin"]), "-a", "2.0"] + list(extra) \
        + ["-v", "-o", str(out) + "/"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=cwd)
    assert res.returncode == 0, (res.stdout[-1500:], res.stderr[-1500:])
    return res.stdout


def _check_run(gold, log, out):
    spec = gold["spec"]
    lines = [l for l in log.splitlines() if " iteration " in l]
    assert len(lines) == spec["nsims"]
    for it, (line, g) in enumerate(zip(lines, gold["iterations"])):
        check_iteration(g, read_ddm(os.path.join(out, "V-%d.ddm" % it)), read_ddm(os.path.join(out, "U-%d.ddm" % it)))
        check_log_line(line, g)
    final = float(re.search(r"Final Avg RMSE: ([-\d.e+]+)", log).group(1))
    printed = float(gold["final_avg_rmse_printed"])
    assert abs(final - printed) <= 1e-5 * max(1.0, abs(printed))
    check_outputs(gold, out)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ml100k", "chembl20"])
def test_bpmf_executable_on_the_reference_data(exe, tmp_path, name):
    """`bpmf -n train -p test -i N -b B -a 2 -v -o out/` (the reference's command line) on the reference's data"""
    gold = fixture(name)
    out = tmp_path / "out"
    out.mkdir()
    log = _run_exe(exe, os.path.join(GOLD, gold["inputs"]["train"]), os.path.join(GOLD, gold["inputs"]["test"]), gold["spec"], out,
                   extra=["-d", str(gold["spec"]["K"])])
    _check_run(gold, log, str(out))
    m = re.search(r"num movs: (\d+)", log)
    if m:
        assert int(m.group(1)) == gold["structure"]["shape"][1]


@pytest.mark.gpu
def test_bpmf_executable_compressed_matrix_market(exe, tmp_path):
    """the reference's ctest pair bpmf_uncompressed / bpmf_compressed (CMakeLists.txt:174-182): `bpmf -i 4` on ML-100K as
    .mtx and as .mtx.gz must succeed — and here also reproduce the reference's first four iterations"""
    gold = fixture("ml100k")
    train, test = inputs(gold)
    spec = dict(gold["spec"], nsims=4)
    logs = []
    for ext, opener in ((".mtx", open), (".mtx.gz", gzip.open)):
        paths = []
        for tag, (shape, rows, cols, vals) in (("train", train), ("test", test)):
            p = tmp_path / (tag + ext)
            with opener(p, "wt") as f:
                f.write("%%MatrixMarket matrix coordinate real general\n% ML-100K\n" + "%d %d %d\n" % (shape[0], shape[1], len(vals)))
                f.write("".join("%d %d %.17g\n" % (r + 1, c + 1, v) for r, c, v in zip(rows, cols, vals)))
            paths.append(p)
        out = tmp_path / ("out" + ext)
        out.mkdir()
        log = _run_exe(exe, paths[0], paths[1], spec, out, extra=["-d", "32"])
        lines = [l for l in log.splitlines() if " iteration " in l]
        assert len(lines) == 4
        for it in range(4):
            check_iteration(gold["iterations"][it], read_ddm(out / ("V-%d.ddm" % it)), read_ddm(out / ("U-%d.ddm" % it)))
            check_log_line(lines[it], gold["iterations"][it])
        logs.append([LINE.match(l).groups() for l in lines])
    assert logs[0] == logs[1]


@pytest.mark.gpu
def test_reference_executable_with_the_b200_back_end_on_ml100k(tmp_path):
    """the reference's own main loop / host predict / file readers (compiled unmodified) driving libbpmf_b200.so through
    oracle/cuda_comm/cuda_comm.h, on its own ctest data"""
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "bpmf_ref_cuda_k32")
    if not os.path.exists(ref_exe):
        pytest.skip("oracle/_ref/bpmf_ref_cuda_k32 was not built (needs the reference sources at build time)")
    gold = fixture("ml100k")
    out = tmp_path / "out"
    out.mkdir()
    log = _run_exe(ref_exe, os.path.join(GOLD, gold["inputs"]["train"]), os.path.join(GOLD, gold["inputs"]["test"]), gold["spec"], out,
                   cwd=tmp_path)
    _check_run(gold, log, str(out))
