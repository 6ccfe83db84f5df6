"""Lane-level numpy emulation of the K=32 warp tail of items_k32_kernel (bpmf_b200/csrc/fast_kernels.cu):
blocked right-looking Cholesky on the matrix held in DMMA accumulator layout, scatter of L, and the two
triangular solves. Used to check the index arithmetic on the CPU before spending GPU time.

Layout: lower-triangle 8x8 blocks (I >= J), block index IJ = I*(I+1)/2 + J. Lane l = 4*g + t holds
M[8I+g][8J+2t+e] in c[IJ][e], e in {0,1}  (the C/D fragment of mma.m8n8k4.f64).
"""
import numpy as np

LANES = np.arange(32)
G, T = LANES >> 2, LANES & 3


def shfl(v, src):
    return v[src]


def blk(I, J):
    return I * (I + 1) // 2 + J


def to_layout(M):
    c = np.zeros((10, 2, 32))
    for I in range(4):
        for J in range(I + 1):
            for e in range(2):
                c[blk(I, J), e] = M[8 * I + G, 8 * J + 2 * T + e]
    return c


def dmma(cacc, a, b):
    """D = A*B + C with A[g][k] from lane 4g+k, B[k][n] from lane 4n+k, C[g][2t+e]."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    A[G, T] = a
    B[T, G] = b
    P = A @ B
    out = cacc.copy()
    for e in range(2):
        out[e] = cacc[e] + P[G, 2 * T + e]
    return out


def block_cholesky(c):
    """In place: returns (c with L in the lower triangle, rsq[4][32] = 1/L[8I+g][8I+g] per lane, ok)."""
    c = c.copy()
    rsq = np.zeros((4, 32))
    ok = True
    for kb in range(4):
        D = blk(kb, kb)
        for k2 in range(4):
            for e in range(2):
                k = 2 * k2 + e
                # pivot: lane (g=k, t=k2), register e
                p = shfl(c[D, e], np.full(32, 4 * k + k2))
                ok = ok and bool(np.all(p > 0))
                rs = 1.0 / np.sqrt(p)
                rsq[kb] = np.where(G == k, rs, rsq[kb])
                # scale column k of the diagonal block and of the panel blocks (lanes t == k2, register e)
                own = T == k2
                for I in range(kb, 4):
                    b_ = blk(I, kb)
                    c[b_, e] = np.where(own, c[b_, e] * rs, c[b_, e])
                # L[2t+e'][k] for this lane's two columns, from the diagonal block: lane (g = 2t+e', t = k2), register e
                bl0 = shfl(c[D, e], 4 * (2 * T + 0) + k2)
                bl1 = shfl(c[D, e], 4 * (2 * T + 1) + k2)
                for I in range(kb, 4):
                    b_ = blk(I, kb)
                    a = shfl(c[b_, e], (LANES & ~3) | k2)      # X[g][k] from the quad's lane t = k2
                    # columns 2t+e' > k only
                    c[b_, 0] = np.where(2 * T + 0 > k, c[b_, 0] - a * bl0, c[b_, 0])
                    c[b_, 1] = np.where(2 * T + 1 > k, c[b_, 1] - a * bl1, c[b_, 1])
        # trailing update A(I,J) -= X_I X_J^T for kb < J <= I, two k-chunks of 4
        frag = {}
        for I in range(kb + 1, 4):
            b_ = blk(I, kb)
            for kk in range(2):
                src = (LANES & ~3) | (2 * kk + (T >> 1))
                v0, v1 = shfl(c[b_, 0], src), shfl(c[b_, 1], src)
                frag[I, kk] = np.where((T & 1) == 0, v0, v1)   # X_I[g][4kk + t]
        for I in range(kb + 1, 4):
            for J in range(kb + 1, I + 1):
                for kk in range(2):
                    c[blk(I, J)] = dmma(c[blk(I, J)], -frag[I, kk], frag[J, kk])
    return c, rsq, ok


def col_off(k):
    return 32 * k - k * (k - 1) // 2


def scatter_L(c):
    Lp = np.zeros(528)
    for I in range(4):
        for J in range(I + 1):
            for e in range(2):
                i, k = 8 * I + G, 8 * J + 2 * T + e
                m = i >= k
                Lp[(col_off(k) + i - k)[m]] = c[blk(I, J), e][m]
    return Lp


def solves(Lp, rs_lane, b, z):
    """lane j owns row j: forward L y = b, y += z, backward L^T x = y. rs_lane[j] = 1 / L[j][j]."""
    b = b.copy()
    y = np.zeros(32)
    for k in range(32):
        yk = shfl(b * rs_lane, np.full(32, k))
        y = np.where(LANES == k, yk, y)
        Ljk = np.where(LANES > k, Lp[np.clip(col_off(k) + LANES - k, 0, 527)], 0.0)
        b = b - Ljk * yk
    y = y + z
    x = np.zeros(32)
    for i in range(31, -1, -1):
        xi = shfl(y * rs_lane, np.full(32, i))
        x = np.where(LANES == i, xi, x)
        Lik = np.where(LANES < i, Lp[np.clip(np.array([col_off(k) for k in LANES]) + i - LANES, 0, 527)], 0.0)
        y = y - Lik * xi
    return x


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    worst = 0.0
    for trial in range(20):
        A = rng.normal(size=(32, 64))
        M = A @ A.T / 64 + np.eye(32) * rng.uniform(0.01, 2.0)
        c, rsq, ok = block_cholesky(to_layout(M))
        assert ok
        Lp = scatter_L(c)
        L = np.zeros((32, 32))
        for k in range(32):
            L[k:, k] = Lp[col_off(k):col_off(k) + 32 - k]
        Lref = np.linalg.cholesky(M)
        e1 = np.abs(L - Lref).max()
        # rsq[I][lane] holds 1/L[8I+g][8I+g]; lane j wants 1/L[j][j]: from lane 4*(j&7) (any t), block j>>3
        rs_lane = np.array([rsq[j >> 3][4 * (j & 7)] for j in range(32)])
        assert np.allclose(rs_lane, 1 / np.diag(Lref))
        b, z = rng.normal(size=32), rng.normal(size=32)
        x = solves(Lp, rs_lane, b, z)
        xref = np.linalg.solve(Lref.T, np.linalg.solve(Lref, b) + z)
        e2 = np.abs(x - xref).max()
        worst = max(worst, e1, e2)
    print("max |L - chol| / |x - xref| over 20 SPD matrices: %.3e" % worst)
    assert worst < 1e-12
