"""GPU test of the `bpmf` executable (bpmf_b200/host): the reference's command line (-n -p -i -b -a -o -v, c++/bpmf.cpp:83)
run end to end on small problems, compared with the CPU oracle: per-iteration U-i.ddm / V-i.ddm dumps to 1e-10, the
per-iteration log line and "Final Avg RMSE" to 1e-6, and the -o posterior files against the oracle's aggregates."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "bpmf_b200", "host")
EXE = os.path.join(HOST, "bpmf")


def write_mtx(path, shape, rows, cols, vals):
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (shape[0], shape[1], len(vals)))
        for r, c, v in zip(rows, cols, vals):
            f.write("%d %d %.17g\n" % (r + 1, c + 1, v))


def read_ddm(path):
    raw = open(path, "rb").read()
    nr, nc = struct.unpack_from("<QQ", raw, 0)
    return np.frombuffer(raw, "<f8", nr * nc, 16).reshape(nc, nr)   # [item, k]


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", HOST, "-s", "bpmf"])
    return EXE


@pytest.mark.parametrize("K,shape,nnz,extra", [(32, (300, 200), 6000, []), (10, (4, 2), None, []), (16, (150, 90), 2500, ["-x", "-H"]),
                                               (32, (500, 310), 9000, ["-g", "2"])])
def test_cli_matches_oracle(exe, tmp_path, K, shape, nnz, extra):
    if "-g" in extra:
        import torch
        if torch.cuda.device_count() < int(extra[extra.index("-g") + 1]):
            pytest.skip("needs more GPUs")
    if nnz is None:   # the reference's data/tiny
        train, test = util.TINY_TRAIN, util.TINY_TEST
    else:
        train, test = util.synth_ratings(shape[0], shape[1], nnz, 11 + K, skew=0.3)
    write_mtx(tmp_path / "train.mtx", *train)
    write_mtx(tmp_path / "test.mtx", *test)
    out = tmp_path / "out"
    out.mkdir()
    nsims, burnin = 6, 2
    cmd = [exe, "-n", str(tmp_path / "train.mtx"), "-p", str(tmp_path / "test.mtx"), "-i", str(nsims), "-b", str(burnin),
           "-a", "2.0", "-d", str(K), "-v", "-o", str(out)] + extra
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    log = res.stdout

    orc = util.make_oracle(K, train, test, alpha=2.0, burnin=burnin, keep_aggr=True)
    lines = [l for l in log.splitlines() if " iteration " in l]
    assert len(lines) == nsims
    for it in range(nsims):
        orc.iterate()
        U, V = read_ddm(out / ("U-%d.ddm" % it)), read_ddm(out / ("V-%d.ddm" % it))
        for got, side in ((V, util.MOVIES), (U, util.USERS)):
            ref = orc.items(side)
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), (it, side)
            assert np.abs(got.mean(0) - ref.mean(0)).max() <= 1e-10
        m = re.match(r"0: (Burnin|Sampling) iteration (\d+):\t RMSE: ([-\d.naif]+)\tavg RMSE: ([-\d.naif]+)\tFU\(\s*([\d.]+)\)\tFM\(\s*([\d.]+)\)", lines[it])
        assert m, lines[it]
        assert m.group(1) == ("Burnin" if it < burnin else "Sampling") and int(m.group(2)) == it
        rm = orc.rmse(util.MOVIES)
        assert abs(float(m.group(3)) - rm[0]) <= 1e-4 + 1e-6   # the log prints 4 decimals
        assert abs(float(m.group(4)) - rm[1]) <= 1e-4 + 1e-6
    orc.finish()
    final = float(re.search(r"Final Avg RMSE: ([-\d.e+]+)", log).group(1))
    assert abs(final - orc.rmse(util.MOVIES)[1]) <= 1e-5 * max(1.0, final)   # operator<< prints 6 significant digits

    # -o outputs: Pavg / Pm2 and the finalized posterior mean (aggrMu / nsamples, c++/bpmf.cpp:281-295)
    raw = open(out / "Pavg.sdm", "rb").read()
    nr, nc, n = struct.unpack_from("<QQQ", raw, 0)
    pavg = np.frombuffer(raw, "<f8", n, 24 + 8 * n)
    np.testing.assert_allclose(pavg, orc.pred(util.MOVIES)[0], rtol=0, atol=1e-6)
    nsamples = nsims - burnin
    for name, side in (("U", util.USERS), ("V", util.MOVIES)):
        amu, alam = orc.aggr(side)
        mu = read_ddm(out / (name + "-mu.ddm"))
        np.testing.assert_allclose(mu, amu / nsamples, rtol=0, atol=1e-10 * max(1.0, np.abs(amu).max()))
        lam = read_ddm(out / (name + "-Lambda.ddm"))
        assert lam.shape == (orc.num(side), K * K)
        # precision = inverse of the sample covariance of the post-burn-in draws; with fewer samples than K it is
        # singular to rounding, so compare through the covariance it was computed from where that is well defined
        if nsamples > K:
            cov = (alam - amu[:, :, None].repeat(K, 2).reshape(-1, K * K) * np.tile(amu, (1, K)) / nsamples) / (nsamples - 1)
            for i in range(0, orc.num(side), max(1, orc.num(side) // 5)):
                np.testing.assert_allclose(lam[i].reshape(K, K) @ cov[i].reshape(K, K), np.eye(K), atol=1e-6)


def test_cli_usage_errors(exe, tmp_path):
    assert subprocess.run([exe], capture_output=True).returncode != 0          # no arguments: usage + abort
    res = subprocess.run([exe, "-n", str(tmp_path / "nope.mtx"), "-p", str(tmp_path / "nope.mtx")], capture_output=True)
    assert res.returncode != 0


def test_cli_propagated_posterior(exe, tmp_path):
    """-m MU,LAMBDA (movies) and -l MU,LAMBDA (users): per-item priors read from .ddm files (c++/bpmf.cpp:134-135,
    c++/sample.cpp:157-174) — the chained-run mechanism of the reference — against the oracle with the same priors."""
    K = 16
    train, test = util.synth_ratings(80, 60, 1500, 77)
    write_mtx(tmp_path / "train.mtx", *train)
    write_mtx(tmp_path / "test.mtx", *test)
    out = tmp_path / "out"
    out.mkdir()
    orc = util.make_oracle(K, train, test, alpha=2.0, burnin=1)
    rng = np.random.default_rng(3)
    args = []
    for flag, side, name in (("-m", util.MOVIES, "V"), ("-l", util.USERS, "U")):
        n = orc.num(side)
        lam = np.stack([util.random_spd(K, 50 * side + i, scale=1.5).T.reshape(-1) for i in range(n)])
        mu = rng.normal(size=(n, K))
        orc.set_prop(side, mu, lam)
        for tag, a in (("mu", mu), ("Lambda", lam)):
            with open(tmp_path / ("%s-%s.ddm" % (name, tag)), "wb") as f:
                f.write(struct.pack("<QQ", a.shape[1], a.shape[0]) + np.ascontiguousarray(a, "<f8").tobytes())
        args += [flag, "%s,%s" % (tmp_path / (name + "-mu.ddm"), tmp_path / (name + "-Lambda.ddm"))]
    nsims = 3
    cmd = [exe, "-n", str(tmp_path / "train.mtx"), "-p", str(tmp_path / "test.mtx"), "-i", str(nsims), "-b", "1", "-d", str(K),
           "-v", "-o", str(out)] + args
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "with propagated posterior" in res.stdout
    for it in range(nsims):
        orc.iterate()
        for name, side in (("V", util.MOVIES), ("U", util.USERS)):
            got, ref = read_ddm(out / ("%s-%d.ddm" % (name, it))), orc.items(side)
            assert np.abs(got - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), (it, name)
