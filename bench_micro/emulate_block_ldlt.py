"""Lane-level numpy emulation of the K=32 warp tail of items_stream32_kernel v3 (bpmf_b200/csrc/stream_kernel.cu):
blocked right-looking LDL^T on the matrix held in DMMA accumulator layout (no square root and no column scaling on
the pivot chain), scatter of the unit-lower factor, and the two triangular solves as pure shuffle + FMA chains.

    A = Lu D Lu^T,  L = Lu D^(1/2)  =>  x = L^-T (L^-1 b + z) = Lu^-T (D^-1 Lu^-1 b + D^(-1/2) z)

Same layout as emulate_block_chol.py: lane l = 4g + t holds M[8I+g][8J+2t+e] in c[blk(I,J)][e].
"""
import numpy as np

from emulate_block_chol import G, LANES, T, blk, dmma, shfl, to_layout


def block_ldlt(c):
    c = c.copy()
    myd = np.zeros(32)
    myrinv = np.zeros(32)
    ok = True
    for kb in range(4):
        D = blk(kb, kb)
        for k2 in range(4):
            for e in range(2):
                k = 2 * k2 + e
                p = shfl(c[D, e], np.full(32, 4 * k + k2))          # pivot d_k
                ok = ok and bool(np.all(p > 0))
                rinv = 1.0 / p
                sel = LANES == 8 * kb + k
                myd = np.where(sel, p, myd)
                myrinv = np.where(sel, rinv, myrinv)
                bl0 = shfl(c[D, e], 4 * (2 * T + 0) + k2)           # a[2t][k], a[2t+1][k] (unscaled)
                bl1 = shfl(c[D, e], 4 * (2 * T + 1) + k2)
                bl0 = np.where(2 * T + 0 > k, -bl0 * rinv, 0.0)
                bl1 = np.where(2 * T + 1 > k, -bl1 * rinv, 0.0)
                for I in range(kb, 4):
                    b_ = blk(I, kb)
                    a = shfl(c[b_, e], (LANES & ~3) | k2)
                    c[b_, 0] = c[b_, 0] + a * bl0
                    c[b_, 1] = c[b_, 1] + a * bl1
        if kb < 3:
            frag = {}
            for I in range(kb + 1, 4):
                b_ = blk(I, kb)
                for kk in range(2):
                    src = (LANES & ~3) | (2 * kk + (T >> 1))
                    v0, v1 = shfl(c[b_, 0], src), shfl(c[b_, 1], src)
                    frag[I, kk] = np.where((T & 1) == 0, v0, v1)     # A~_I[g][4kk + t]
            rv = [shfl(myrinv, 8 * kb + 4 * kk + T) for kk in range(2)]
            for I in range(kb + 1, 4):
                for J in range(kb + 1, I + 1):
                    for kk in range(2):
                        c[blk(I, J)] = dmma(c[blk(I, J)], -frag[I, kk], frag[J, kk] * rv[kk])
    return c, myd, myrinv, ok


def col_off1(k):
    return 31 * k - k * (k - 1) // 2


def scatter_unit_lower(c, myrinv):
    Lp = np.zeros(496)
    for I in range(4):
        for J in range(I + 1):
            for e in range(2):
                i, k = 8 * I + G, 8 * J + 2 * T + e
                m = i > k
                v = c[blk(I, J), e] * shfl(myrinv, k)
                Lp[(col_off1(k) + i - k - 1)[m]] = v[m]
    return Lp


def solves(Lp, myrinv, myrs, b, z):
    bb = b.copy()
    for k in range(31):
        t = shfl(bb, np.full(32, k))
        Ljk = np.where(LANES > k, Lp[np.clip(col_off1(k) + LANES - k - 1, 0, 495)], 0.0)
        bb = bb - Ljk * t
    yv = bb * myrinv + myrs * z
    co = np.array([col_off1(k) for k in LANES])
    for i in range(31, 0, -1):
        xi = shfl(yv, np.full(32, i))
        Lik = np.where(LANES < i, Lp[np.clip(co + i - LANES - 1, 0, 495)], 0.0)
        yv = yv - Lik * xi
    return yv


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    worst = 0.0
    for trial in range(20):
        A = rng.normal(size=(32, 64))
        M = A @ A.T / 64 + np.eye(32) * rng.uniform(0.01, 2.0)
        c, myd, myrinv, ok = block_ldlt(to_layout(M))
        assert ok
        Lref = np.linalg.cholesky(M)
        assert np.allclose(myd, np.diag(Lref) ** 2, rtol=1e-11)
        Lp = scatter_unit_lower(c, myrinv)
        Lu = np.eye(32)
        for k in range(31):
            Lu[k + 1:, k] = Lp[col_off1(k):col_off1(k) + 31 - k]
        e1 = np.abs(Lu * np.sqrt(myd)[None, :] - Lref).max()
        b, z = rng.normal(size=32), rng.normal(size=32)
        x = solves(Lp, myrinv, 1.0 / np.sqrt(myd), b, z)
        xref = np.linalg.solve(Lref.T, np.linalg.solve(Lref, b) + z)
        e2 = np.abs(x - xref).max()
        worst = max(worst, e1, e2)
    print("max |Lu sqrt(D) - chol| / |x - xref| over 20 SPD matrices: %.3e" % worst)
    assert worst < 1e-11
