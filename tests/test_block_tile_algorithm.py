"""The tail of the CTA-per-item kernel (bpmf_b200/csrc/block_kernel.cu), emulated lane by lane in numpy on the kernel's own
data layout and driven by the REAL trailing-update schedule of the product library (bpmf_gpu_debug_block_schedule, a
host-only entry point): swizzled 8 x 8 tiles of the lower block triangle, the right-hand side as block row NB, the
register-resident factorisation of a diagonal tile with the inverse of its unit-lower factor from the same row
operations, panel tiles as one 8x8x8 product with that inverse (DMMA fragment semantics), trailing quads sharing their
fragments, the backward solve. The result must be the solution numpy computes: what the schedule and the layout algebra
promise, checked without a GPU."""
import ctypes

import numpy as np
import pytest


def swz(r):
    return (r & 2) << 1


def tile(I, J):
    return (I * (I + 1) // 2 + J) * 64


def pos(r, c):
    return 8 * r + (c ^ swz(r))


G, T = np.arange(32) >> 2, np.arange(32) & 3            # lane = 4 g + t
FPOS = np.array([pos(g, t) for g, t in zip(G, T)])        # A / B fragment element (g, t); (g, t + 4) is at FPOS ^ 4
CPOS = np.array([pos(g, 2 * t) for g, t in zip(G, T)])    # C fragment pair (g, 2t), (g, 2t + 1)


def dmma(c0, c1, a, b):
    """mma.sync.m8n8k4.f64: D[m][n] += sum_k A[m][k] B[k][n]; lane (g, t) holds A[g][t], B[t][g], C[g][2t], C[g][2t+1]"""
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    A[G, T] = a
    B[T, G] = b
    D = A @ B
    return c0 + D[G, 2 * T], c1 + D[G, 2 * T + 1]


class Tiles:
    def __init__(self, NB):
        self.NB = NB
        self.ntile = NB * (NB + 1) // 2
        self.LINV = (self.ntile + NB) * 64
        self.m = np.zeros((self.ntile + 2 * NB) * 64)

    def store_matrix(self, MM, b):
        NB = self.NB
        for I in range(NB):
            for J in range(I + 1):
                for r in range(8):
                    for c in range(8):
                        self.m[tile(I, J) + pos(r, c)] = MM[8 * I + r, 8 * J + c]
        for e in range(NB * 64):
            self.m[tile(NB, 0) + e] = b[(e >> 6) * 8 + (e & 7)] if (e & 63) < 8 else 0.0
            er, ec = (e >> 3) & 7, e & 7
            if ec >= er:
                self.m[self.LINV + (e & ~7) + (ec ^ swz(er))] = 1.0 if ec == er else 0.0

    def factor_diag(self, kb, sd, srinv):
        tp, lp = tile(kb, kb), self.LINV + 64 * kb
        a = np.zeros((8, 8)); W = np.zeros((8, 8))
        for i in range(8):
            for j in range(i + 1):
                a[i, j] = self.m[tp + pos(i, j)]
        for k in range(8):
            d = a[k, k]
            assert d > 0
            sd[8 * kb + k], srinv[8 * kb + k] = d, 1.0 / d
            l = a[:, k] / d
            for j in range(k + 1, 8):
                for i in range(j, 8):
                    a[i, j] -= l[i] * a[j, k]
            for i in range(k + 1, 8):
                for j in range(k):
                    W[i, j] -= l[i] * W[k, j]
                W[i, k] = -l[i]
                self.m[tp + pos(i, k)] = l[i]
        for i in range(1, 8):
            for j in range(i):
                self.m[lp + pos(i, j)] = W[i, j]

    def panel(self, I, kb, srinv):
        tp, lp = tile(I, kb), self.LINV + 64 * kb
        c0, c1 = dmma(0.0, 0.0, self.m[tp + FPOS], self.m[lp + FPOS])           # B[k][n] = Linv[n][k]
        c0, c1 = dmma(c0, c1, self.m[tp + (FPOS ^ 4)], self.m[lp + (FPOS ^ 4)])
        self.m[tp + CPOS] = c0 * srinv[8 * kb + 2 * T]
        self.m[tp + CPOS + 1] = c1 * srinv[8 * kb + 2 * T + 1]

    def update(self, a_off, b_off, c_off, kb, sd):
        dc0, dc1 = sd[8 * kb + T], sd[8 * kb + 4 + T]
        c0, c1 = self.m[c_off + CPOS], self.m[c_off + CPOS + 1]
        c0, c1 = dmma(c0, c1, -self.m[a_off + FPOS], self.m[b_off + FPOS] * dc0)
        c0, c1 = dmma(c0, c1, -self.m[a_off + (FPOS ^ 4)], self.m[b_off + (FPOS ^ 4)] * dc1)
        self.m[c_off + CPOS], self.m[c_off + CPOS + 1] = c0, c1


@pytest.fixture(scope="module")
def lib():
    import bpmf_b200
    return bpmf_b200.load_library()


@pytest.mark.parametrize("K", [16, 48, 64, 128])
def test_tile_algorithm_with_the_library_schedule_solves_the_system(lib, K):
    NB, NWB = K // 8, K // 16
    rng = np.random.default_rng(K)
    Y = rng.normal(size=(3 * K, K))
    MM = 2.0 * (Y.T @ Y) / K + np.eye(K) + 0.1 * np.diag(rng.random(K))      # LambdaF + alpha G, well conditioned
    b, z = rng.normal(size=K), rng.normal(size=K)
    t = Tiles(NB)
    t.store_matrix(MM, b)
    sd, srinv = np.zeros(K), np.zeros(K)
    buf = np.empty(8 * 64, np.int32)
    t.factor_diag(0, sd, srinv)
    for kb in range(NB):
        for I in range(kb + 1, NB + 1):                   # block row NB = the right-hand side
            t.panel(I, kb, srinv)
        if kb + 1 < NB:
            t.update(tile(kb + 1, kb), tile(kb + 1, kb), tile(kb + 1, kb + 1), kb, sd)     # warp 0: the next diagonal tile ...
            quads = []
            for w in range(NWB):
                n = lib.bpmf_gpu_debug_block_schedule(K, kb, w, buf.ctypes.data_as(ctypes.c_void_p), 64)
                quads += [tuple(int(x) for x in q) for q in buf[: 8 * n].reshape(n, 8)]
            for a0, a1, b0, b1, c00, c01, c10, c11 in quads:  # ... the other warps: quads of the library's schedule
                for c_off, a_off, b_off in ((c00, a0, b0), (c01, a0, b1), (c10, a1, b0), (c11, a1, b1)):
                    if c_off >= 0:
                        t.update(a_off, b_off, c_off, kb, sd)
            t.factor_diag(kb + 1, sd, srinv)              # ... and its factorisation
    # backward solve on v = D^-1 Lu^-1 b + D^(-1/2) z (row 0 of block row NB)
    v = np.array([t.m[tile(NB, i >> 3) + (i & 7)] for i in range(K)]) + z / np.sqrt(sd)
    w_expected = np.linalg.solve(np.linalg.cholesky(MM), b) / np.sqrt(sd)   # = D^-1 Lu^-1 b since L = Lu D^(1/2)
    assert np.abs(v - z / np.sqrt(sd) - w_expected).max() <= 1e-11 * max(1.0, np.abs(w_expected).max())
    for k in range(K - 1, -1, -1):
        kbk, c = k >> 3, k & 7
        for i in range(k):
            v[i] -= t.m[tile(kbk, i >> 3) + pos(c, i & 7)] * v[k]           # Lu(k, i)
    L = np.linalg.cholesky(MM)
    x_ref = np.linalg.solve(L.T, np.linalg.solve(L, b) + z)                 # L^T \ (L \ b + z)  (sample.cpp:321-323)
    assert np.abs(v - x_ref).max() <= 1e-10 * max(1.0, np.abs(x_ref).max())
