"""PINS THE ORACLE: oracle/bpmf_oracle.hpp (the restatement every parity test compares the CUDA path with) against the
REFERENCE'S OWN SOURCES for the hot path — c++/sample.cpp and c++/mvnormal.cpp compiled unmodified from /root/reference
into oracle/_ref/ (oracle/Makefile target `ref`, oracle/ref_harness.cpp) with stand-in headers for the two libraries
this image lacks (oracle/shim/: Eigen 3, Random123). Everything the reference's code decides is exercised by its own
text: RNG keying and consumption (junk normals included), the hyper draw, the per-item update, the sweep reductions, the
never-updated member `sum`, the running averages of predict and its double-counted last sample, the posterior
aggregation, the propagated-posterior quirk. The order of operations INSIDE Eigen's kernels is the one thing the
stand-in cannot reproduce; with the textbook orders on both sides the comparison is BIT-EXACT.

The libraries are built where /root/reference exists and travel prebuilt otherwise; without either the tests skip."""
import numpy as np
import pytest

import util
from oracle import oracle as orc_mod
from oracle import reference as ref_mod

MOVIES, USERS = 0, 1


def _pair(K, train, test, **kw):
    if not ref_mod.available(K):
        pytest.skip("no reference sources and no prebuilt oracle/_ref for K=%d" % K)
    (s, r, c, v), (_, tr, tc, tv) = train, test
    orc = util.make_oracle(K, train, test, nthreads=1, **kw)
    ref = ref_mod.Reference(K, s, r, c, v, tr, tc, tv, alpha=kw.get("alpha", 2.0), burnin=kw.get("burnin", 5),
                            keep_aggr=kw.get("keep_aggr", False))
    return orc, ref


def _same(a, b):
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    return a.shape == b.shape and a.tobytes() == b.tobytes()


@pytest.mark.parametrize("K", [10, 16, 32, 128])
def test_rng_streams(K):
    """rng_set_pos + randn of c++/mvnormal.cpp on the MicroURNG / Philox stand-in = the oracle's, for the counters the path
    uses: iter (hyper draw) and (idx + 1) * K * (iter + 1) (items), including a 32-bit wrap."""
    if not ref_mod.available(K):
        pytest.skip("no reference build")
    for c in (0, 1, 10, 32, 7 * K * 3, 2 ** 32 - 1, (123457 * K * 9) % 2 ** 32):
        assert _same(ref_mod.randn(K, c, 80), orc_mod.randn(c, 80)), c


@pytest.mark.parametrize("K,shape,nnz,kw", [
    (10, (40, 30), 300, {}),
    (16, (120, 90), 1500, dict(skew=1.0, empty_rows=10, heavy_col=100)),
    (32, (150, 110), 4000, dict(skew=0.8, empty_rows=12)),
    (48, (90, 70), 2500, dict(skew=1.0, heavy_col=60)),
    (64, (90, 70), 2500, {}),
    (128, (60, 50), 1800, dict(heavy_col=55)),
])
def test_chain_matches_reference_sources_bit_for_bit(K, shape, nnz, kw):
    """8 iterations of the reference's main loop (movies.sample(users); users.sample(movies); both predicts,
    c++/bpmf.cpp:184-190) with burnin 3, then the final predict(all) (c++/bpmf.cpp:242): latents, hyper-parameters, cov,
    norm, RMSEs, Pavg / Pm2."""
    train, test = util.synth_ratings(shape[0], shape[1], nnz, 100 + K, **kw)
    orc, ref = _pair(K, train, test, burnin=3)
    for side in (MOVIES, USERS):
        assert orc.mean_rating(side) == ref.scalars(side)["mean_rating"]
    for it in range(8):
        for side in (MOVIES, USERS):
            orc.sweep(side)
            ref.sample(side)
            assert _same(orc.items(side), ref.items(side)), (it, side)
            for a, b in zip(orc.hyper(side), ref.hyper(side)):
                assert _same(a, b), (it, side)
            s, p, cov, norm = orc.stats(side)
            assert _same(cov, ref.cov(side)) and norm == ref.scalars(side)["norm"], (it, side)
        for side in (MOVIES, USERS):
            orc.predict(side)
            ref.predict(side)
            sc = ref.scalars(side)
            assert orc.rmse(side) == (sc["rmse"], sc["rmse_avg"], sc["num_predict"]), (it, side)
            for a, b in zip(orc.pred(side), ref.pred(side)):
                assert _same(a, b), (it, side)
        assert ref.scalars(MOVIES)["iter"] == it == orc.iter(MOVIES)
    orc.finish()                          # movies.predict(users, true): the last sample once more (quirk Q4)
    ref.predict(MOVIES, all=True)
    sc = ref.scalars(MOVIES)
    assert orc.rmse(MOVIES) == (sc["rmse"], sc["rmse_avg"], sc["num_predict"])
    for a, b in zip(orc.pred(MOVIES), ref.pred(MOVIES)):
        assert _same(a, b)


def test_posterior_aggregation_matches_reference_sources():
    """-o: aggrMu / aggrLambda accumulated from iteration `burnin` on (c++/sample.cpp:364-368)."""
    K = 10
    train, test = util.synth_ratings(50, 35, 420, 77)
    orc, ref = _pair(K, train, test, burnin=2, keep_aggr=True)
    for it in range(5):
        for side in (MOVIES, USERS):
            orc.sweep(side)
            ref.sample(side)
    for side in (MOVIES, USERS):
        for a, b in zip(orc.aggr(side), ref.aggr(side)):
            assert _same(a, b) and np.abs(a).max() > 0


@pytest.mark.parametrize("K", [10, 32])
def test_propagated_posterior_matches_reference_sources(K):
    """-m / -l: per-item prior precisions, and the rhs built from the GLOBAL hp.mu (quirk Q5, c++/sample.cpp:272-285)."""
    train, test = util.synth_ratings(60, 45, 700, 5 + K)
    orc, ref = _pair(K, train, test)
    rng = np.random.default_rng(K)
    for side in (MOVIES, USERS):
        n = orc.num(side)
        lam = np.stack([util.random_spd(K, 1000 * side + i, scale=2.0).T.reshape(-1) for i in range(n)])
        mu = rng.normal(size=(n, K))
        orc.set_prop(side, mu, lam)
        ref.set_prop(side, mu, lam)
    for it in range(3):
        for side in (MOVIES, USERS):
            orc.sweep(side)
            ref.sample(side)
            assert _same(orc.items(side), ref.items(side)), (it, side)


def test_cholesky_failure_is_reported_by_both():
    """THROWERROR("Cholesky failed") (c++/sample.cpp:308): a prior precision that is not positive definite."""
    K = 10
    train, test = util.synth_ratings(30, 20, 150, 9)
    orc, ref = _pair(K, train, test)
    for side in (MOVIES, USERS):
        n = orc.num(side)
        lam = np.tile((-np.eye(K)).reshape(-1), (n, 1))
        orc.set_prop(side, np.zeros((n, K)), lam)
        ref.set_prop(side, np.zeros((n, K)), lam)
    with pytest.raises(RuntimeError, match="Cholesky failed"):
        ref.sample(MOVIES)
    with pytest.raises(RuntimeError):
        orc.sweep(MOVIES)


@pytest.mark.parametrize("name", ["tiny_k10", "synth_k32"])
def test_reference_executable_fixtures(name, tmp_path):
    """tests/golden/refexe_chain_*.json were produced by the REFERENCE EXECUTABLE — every translation unit of its `bpmf`
    target compiled unmodified against the stand-in headers (tests/golden/make_ref_exe_golden.py) — on the reference's
    own data/tiny with run_test.sh's arguments and on the synthetic K = 32 problem. The oracle (one thread, like that
    build) must reproduce them exactly: the -v dumps bit for bit, and the log's RMSE / avg RMSE / FU / FM / "Final Avg
    RMSE" fields as printed. Where the executable is available (this container) it is also re-run: the committed
    fixture must be what it produces."""
    import importlib.util
    import json
    import os
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

    def load(modname):
        spec_mod = importlib.util.spec_from_file_location(modname, os.path.join(gdir, modname + ".py"))
        mod = importlib.util.module_from_spec(spec_mod)
        spec_mod.loader.exec_module(mod)
        return mod

    gold = json.load(open(os.path.join(gdir, "refexe_chain_%s.json" % name)))
    spec = gold["spec"]
    train, test = load("make_chain_golden").problem(spec)
    m = util.make_oracle(spec["K"], train, test, alpha=2.0, burnin=spec["burnin"], nthreads=1)
    for it, g in enumerate(gold["iterations"]):
        m.iterate()
        V, U = m.items(MOVIES), m.items(USERS)
        assert [float(x) for x in V.mean(0)] == g["V_mean"] and [float(x) for x in U.mean(0)] == g["U_mean"], it
        assert [float(x) for x in V[0]] == g["V_first"] and [float(x) for x in U[-1]] == g["U_last"], it
        assert float(np.sqrt((V * V).sum())) == g["V_norm"] and float(np.sqrt((U * U).sum())) == g["U_norm"], it
        r = m.rmse(MOVIES)
        log = g["log"]
        assert log["iter"] == it and log["phase"] == ("Burnin" if it < spec["burnin"] else "Sampling")
        assert "%.4f" % r[0] == log["rmse"] and "%.4f" % r[1] == log["rmse_avg"], (it, r, log)
        assert "%.2f" % np.sqrt(m.stats(USERS)[3]) == log["FU"] and "%.2f" % np.sqrt(m.stats(MOVIES)[3]) == log["FM"], (it, log)
    m.finish()
    assert "%g" % m.rmse(MOVIES)[1] == gold["final_avg_rmse_printed"]
    # the oracle's own fixture (made with all threads: the sweep reductions then add in another order, as the
    # reference's do under OpenMP) agrees to round-off
    own = json.load(open(os.path.join(gdir, "chain_%s.json" % name)))
    for a, b in zip(own["iterations"], gold["iterations"]):
        for k in ("V_mean", "U_mean", "V_first", "U_last"):
            assert np.abs(np.array(a[k]) - np.array(b[k])).max() <= 1e-12
    exe = os.path.join(os.path.dirname(gdir), "..", "oracle", "_ref", "bpmf_ref_k%d" % spec["K"])
    if os.path.exists(exe) and os.path.exists("/root/reference/c++/bpmf.cpp"):
        gen = load("make_ref_exe_golden")
        log, out = gen.run_reference_exe(spec, str(tmp_path))
        assert gen.summarise(spec, log, out) == gold
