"""Regenerates tests/golden/chain_*.json: per-iteration K-vector column means, norms and RMSE of whole Gibbs chains as
the oracle computes them (the quantities BASELINE.json's parity gate names), for
  * the reference's own data/tiny (4 users x 2 movies, K=10, -i 9 -b 0: the run of data/tiny/run_test.sh), and
  * a seeded synthetic problem at K=32 (tests/util.synth_ratings(300, 200, 6000, 43, skew=0.3), -i 8 -b 3).
The reference cannot produce these (it does not build here: no Eigen3 / Random123), so the fixtures pin the ORACLE
against accidental change and give the GPU tests a committed target that does not depend on the oracle binary.
Run from the repository root:  python tests/golden/make_chain_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import util  # noqa: E402

CASES = {
    "tiny_k10": dict(K=10, data="tiny", nsims=9, burnin=0),
    "synth_k32": dict(K=32, data=(300, 200, 6000, 43), nsims=8, burnin=3),
}


def problem(spec):
    if spec["data"] == "tiny":
        return util.TINY_TRAIN, util.TINY_TEST
    nr, nc, nnz, seed = spec["data"]
    return util.synth_ratings(nr, nc, nnz, seed, skew=0.3)


def run(spec):
    train, test = problem(spec)
    m = util.make_oracle(spec["K"], train, test, alpha=2.0, burnin=spec["burnin"])
    its = []
    for _ in range(spec["nsims"]):
        m.iterate()
        V, U = m.items(util.MOVIES), m.items(util.USERS)
        r = m.rmse(util.MOVIES)
        its.append({"V_mean": [float(x) for x in V.mean(0)], "U_mean": [float(x) for x in U.mean(0)],
                    "V_norm": float(np.sqrt((V * V).sum())), "U_norm": float(np.sqrt((U * U).sum())),
                    "V_first": [float(x) for x in V[0]], "U_last": [float(x) for x in U[-1]],
                    "rmse": r[0], "rmse_avg": r[1]})
    m.finish()
    return {"spec": spec, "iterations": its, "final_avg_rmse": m.rmse(util.MOVIES)[1]}


if __name__ == "__main__":
    for name, spec in CASES.items():
        out = run(spec)
        json.dump(out, open(os.path.join(HERE, "chain_%s.json" % name), "w"), indent=1)
        print(name, "final avg rmse", out["final_avg_rmse"])
