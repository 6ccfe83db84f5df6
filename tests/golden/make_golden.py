"""Regenerates tests/golden/rng_kat.json from the oracle (which calls this container's libstdc++ <random>, the
code the reference links) — run here, where g++ 13 / libstdc++ is the toolchain the survey pinned.

The reference itself cannot produce fixtures (it does not build without Eigen3 + Random123, neither present),
so these vectors pin the ORACLE's stream against accidental change and carry the survey's derived KATs
 into the test-suite; the Philox core under them is pinned by published vectors in
tests/test_oracle_kat.py.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as o  # noqa: E402

out = {
    "randn": {str(c): [float(x) for x in o.randn(c, 8)] for c in (0, 10, 32, 320, 4294967295)},
    "words": {str(c): [int(x) for x in o.words(c, 8)] for c in (0, 10)},
}
g, n = o.gamma_then_randn(0, 516.0, 1)
out["gamma516_then_randn"] = [g, float(n[0])]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rng_kat.json"), "w"), indent=1)
print(json.dumps(out)[:300])
