"""The warp tail of the K = 32 stream kernel (bpmf_b200/csrc/stream_kernel.cu: blocked right-looking LDL^T on the matrix in
DMMA accumulator layout, the packed unit-lower factor, the two triangular solves as shuffle + FMA chains), emulated lane
by lane in numpy (bench_micro/emulate_block_ldlt.py, the model the kernel was written from): it must factor SPD matrices
like numpy's Cholesky and solve  x = L^-T (L^-1 b + z)  (c++/sample.cpp:321-323). CPU only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bench_micro"))


def test_k32_tail_model_factors_and_solves():
    from emulate_block_chol import to_layout
    from emulate_block_ldlt import block_ldlt, col_off1, scatter_unit_lower, solves
    rng = np.random.default_rng(1)
    for trial in range(12):
        A = rng.normal(size=(32, 64))
        M = A @ A.T / 64 + np.eye(32) * rng.uniform(0.01, 2.0)
        c, myd, myrinv, ok = block_ldlt(to_layout(M))
        assert ok
        Lref = np.linalg.cholesky(M)
        np.testing.assert_allclose(myd, np.diag(Lref) ** 2, rtol=1e-11)
        Lp = scatter_unit_lower(c, myrinv)
        Lu = np.eye(32)
        for k in range(31):
            Lu[k + 1:, k] = Lp[col_off1(k):col_off1(k) + 31 - k]
        assert np.abs(Lu * np.sqrt(myd)[None, :] - Lref).max() < 1e-11
        b, z = rng.normal(size=32), rng.normal(size=32)
        x = solves(Lp, myrinv, 1.0 / np.sqrt(myd), b, z)
        xref = np.linalg.solve(Lref.T, np.linalg.solve(Lref, b) + z)
        assert np.abs(x - xref).max() < 1e-11 * max(1.0, np.abs(xref).max()) * 10


def test_k32_tail_model_reports_a_non_positive_pivot():
    from emulate_block_chol import to_layout
    from emulate_block_ldlt import block_ldlt
    M = np.eye(32)
    M[5, 5] = -1.0                                         # "Cholesky failed" (c++/sample.cpp:308)
    assert not block_ldlt(to_layout(M))[3]
