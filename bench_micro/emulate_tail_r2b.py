"""Lane-level numpy emulation of the K = 32 warp tail AS SHIPPED at the end of round 2 (tail32_warp / chol3_block_column in
bpmf_b200/csrc/stream_kernel.cu), next to the model the first version was written from (emulate_block_ldlt.py):

  * the trailing update without any lane exchange: DMMA number e of a block pair takes column 2t + e in its k-slot t, so the
    accumulator registers of the two panel blocks are its A and B fragments (the sum over the eight columns runs in another
    order than in the first model: results agree to rounding, not to the bit — that is the only difference);
  * columns 6 and 7 of every 8 x 8 block peeled off the column loop (their multipliers are all zero);
  * the factor's panel blocks stored from the trailing update's B fragments (bs = -Lu) and its diagonal blocks with -1 / d,
    i.e. the packed factor is kept NEGATED and the solves add instead of subtracting;
  * the quad sums of the right-hand side through a 4 x 32 scratch: lane 4g + t stores partial a at a * 32 + lane, lane r reads
    the four partials of ITS row side by side and adds them as (x0 + x1) + (x2 + x3).

tests/test_k32_tail_emulation.py holds each of these steps to the first model: the peeled loop, the negated factor and the
quad sums BIT FOR BIT, the whole tail to numpy's Cholesky.
"""
import numpy as np

from emulate_block_chol import G, LANES, T, blk, shfl, to_layout   # noqa: F401
from emulate_block_ldlt import col_off1


def dmma_regs(cacc, a_frag, b_frag):
    """mma.m8n8k4: lane 4g + t supplies A[g][t] = a_frag and B[t][g] = b_frag; C[g][2t + e] accumulates, k-slots in order."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    A[G, T] = a_frag
    B[T, G] = b_frag
    out = cacc.copy()
    for e in range(2):
        acc = cacc[e].copy()
        for k in range(4):                                  # fused multiply-adds in k order (numpy: multiply, then add)
            acc = acc + A[G, k] * B[k, 2 * T + e]
        out[e] = acc
    return out


def column_step(c, kb, k2, e, myd, myrinv, do0=True, do1=True, only_pivot=False):
    D = blk(kb, kb)
    k = 2 * k2 + e
    p = shfl(c[D, e], np.full(32, 4 * k + k2))
    rinv = 1.0 / p
    sel = LANES == 8 * kb + k
    myd[:] = np.where(sel, p, myd)
    myrinv[:] = np.where(sel, rinv, myrinv)
    if only_pivot:
        return
    bl0 = np.where(2 * T + 0 > k, -shfl(c[D, e], 4 * (2 * T + 0) + k2) * rinv, 0.0) if do0 else None
    bl1 = np.where(2 * T + 1 > k, -shfl(c[D, e], 4 * (2 * T + 1) + k2) * rinv, 0.0)
    for I in range(kb, 4):
        b_ = blk(I, kb)
        a = shfl(c[b_, e], (LANES & ~3) | k2)
        if do0:
            c[b_, 0] = c[b_, 0] + a * bl0
        if do1:
            c[b_, 1] = c[b_, 1] + a * bl1


def block_ldlt_r2b(c, peel=True, Lp=None):
    """-> c (L D in the lower triangle), d, 1 / d, ok; if Lp is given the NEGATED packed factor is stored on the way."""
    c = c.copy()
    myd, myrinv = np.zeros(32), np.zeros(32)
    for kb in range(4):
        D = blk(kb, kb)
        for k2 in range(3 if peel else 4):
            for e in range(2):
                column_step(c, kb, k2, e, myd, myrinv)
        if peel:
            column_step(c, kb, 3, 0, myd, myrinv, do0=False)          # column 6 updates column 7 only
            column_step(c, kb, 3, 1, myd, myrinv, only_pivot=True)    # column 7: its pivot
        nrv = [-shfl(myrinv, 8 * kb + 2 * T + e) for e in range(2)]   # -1 / d of column 2t + e
        if Lp is not None:
            for e in range(2):
                k = 8 * kb + 2 * T + e
                m = G > 2 * T + e
                Lp[(col_off1(k) + G - (2 * T + e) - 1)[m]] = (c[D, e] * nrv[e])[m]
        if kb < 3:
            bs = {(J, e): c[blk(J, kb), e] * nrv[e] for J in range(kb + 1, 4) for e in range(2)}
            if Lp is not None:
                for e in range(2):
                    k = 8 * kb + 2 * T + e
                    for J in range(kb + 1, 4):
                        Lp[col_off1(k) + 8 * J + G - k - 1] = bs[J, e]
            for I in range(kb + 1, 4):
                for J in range(kb + 1, I + 1):
                    for e in range(2):
                        c[blk(I, J)] = dmma_regs(c[blk(I, J)], c[blk(I, kb), e], bs[J, e])
    return c, myd, myrinv, bool(np.all(myd > 0))


def solves_negated(Lpn, myrinv, myrs, b, z):
    bb = b.copy()
    for k in range(31):
        t = shfl(bb, np.full(32, k))
        Ljk = np.where(LANES > k, Lpn[np.clip(col_off1(k) + LANES - k - 1, 0, 495)], 0.0)
        bb = bb + Ljk * t
    yv = bb * myrinv + myrs * z
    co = np.array([col_off1(k) for k in LANES])
    for i in range(31, 0, -1):
        xi = shfl(yv, np.full(32, i))
        Lik = np.where(LANES < i, Lpn[np.clip(co + i - LANES - 1, 0, 495)], 0.0)
        yv = yv + Lik * xi
    return yv


def quad_sums_shuffle(rrp):
    """the first version: two xor-shuffle rounds, lane t == 0 of a quad keeps row 8a + g"""
    r = rrp.copy()
    for a in range(4):
        r[a] = r[a] + shfl(r[a], LANES ^ 1)
        r[a] = r[a] + shfl(r[a], LANES ^ 2)
    out = np.zeros(32)
    for a in range(4):
        out[8 * a + G[T == 0]] = r[a][T == 0]
    return out


def quad_sums_shared(rrp):
    rs = np.zeros(128)
    for a in range(4):
        rs[a * 32 + LANES] = rrp[a]
    v = rs.reshape(32, 4)                                    # lane r reads rs[4r .. 4r + 3]
    return (v[:, 0] + v[:, 1]) + (v[:, 2] + v[:, 3])
