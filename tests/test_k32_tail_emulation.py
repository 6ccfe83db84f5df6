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


def _spd(rng):
    A = rng.normal(size=(32, 64))
    return A @ A.T / 64 + np.eye(32) * rng.uniform(0.01, 2.0)


def test_shipped_tail_steps_match_the_first_model_bit_for_bit():
    """bench_micro/emulate_tail_r2b.py: the steps the second half of round 2 changed. Peeling columns 6 and 7 off the column loop
    changes no bit; the negated packed factor is the first model's factor with the sign flipped, bit for bit, and the solves on
    it give the same bits; the quad sums through the scratch are the shuffle tree's sums."""
    from emulate_block_chol import to_layout
    from emulate_block_ldlt import scatter_unit_lower, solves
    from emulate_tail_r2b import block_ldlt_r2b, quad_sums_shared, quad_sums_shuffle, solves_negated
    rng = np.random.default_rng(7)
    for trial in range(8):
        c0 = to_layout(_spd(rng))
        full = block_ldlt_r2b(c0, peel=False)
        Lpn = np.zeros(496)
        peeled = block_ldlt_r2b(c0, peel=True, Lp=Lpn)
        for a, b in zip(full[:3], peeled[:3]):
            assert a.tobytes() == b.tobytes()
        assert full[3] and peeled[3]
        c, myd, myrinv, _ = peeled
        Lp = scatter_unit_lower(c, myrinv)                  # the first model's scatter, from the same registers
        assert (-Lp).tobytes() == Lpn.tobytes()
        b, z = rng.normal(size=32), rng.normal(size=32)
        myrs = 1.0 / np.sqrt(myd)
        assert solves(Lp, myrinv, myrs, b, z).tobytes() == solves_negated(Lpn, myrinv, myrs, b, z).tobytes()
        rrp = rng.normal(size=(4, 32))
        assert quad_sums_shuffle(rrp).tobytes() == quad_sums_shared(rrp).tobytes()


def test_shipped_tail_factors_and_solves():
    from emulate_block_chol import to_layout
    from emulate_block_ldlt import col_off1
    from emulate_tail_r2b import block_ldlt_r2b, solves_negated
    rng = np.random.default_rng(11)
    for trial in range(8):
        M = _spd(rng)
        Lpn = np.zeros(496)
        c, myd, myrinv, ok = block_ldlt_r2b(to_layout(M), Lp=Lpn)
        assert ok
        Lref = np.linalg.cholesky(M)
        np.testing.assert_allclose(myd, np.diag(Lref) ** 2, rtol=1e-11)
        Lu = np.eye(32)
        for k in range(31):
            Lu[k + 1:, k] = -Lpn[col_off1(k):col_off1(k) + 31 - k]
        assert np.abs(Lu * np.sqrt(myd)[None, :] - Lref).max() < 1e-11
        b, z = rng.normal(size=32), rng.normal(size=32)
        x = solves_negated(Lpn, myrinv, 1.0 / np.sqrt(myd), b, z)
        xref = np.linalg.solve(Lref.T, np.linalg.solve(Lref, b) + z)
        assert np.abs(x - xref).max() < 1e-10 * max(1.0, np.abs(xref).max())
    M = np.eye(32)
    M[29, 29] = -1.0
    assert not block_ldlt_r2b(to_layout(M))[3]
