"""ctypes binding of oracle/libbpmf_oracle.so (the CPU restatement of the reference). Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbpmf_oracle.so")
_lib = None

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the oracle with the committed Makefile (g++ only, no GPU needed)."""
    src = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "bpmf_oracle.hpp", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    L.bpmf_oracle_last_error.restype = C.c_char_p
    L.bpmf_oracle_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
    L.bpmf_oracle_words.argtypes = [C.c_uint32, C.c_int, _u32p]
    L.bpmf_oracle_randn.argtypes = [C.c_uint32, C.c_int, _f64p]
    L.bpmf_oracle_gamma_then_randn.argtypes = [C.c_uint32, C.c_double, C.c_int, C.POINTER(C.c_double), _f64p]
    L.bpmf_oracle_chol_lower.argtypes = [_f64p, C.c_int]
    L.bpmf_oracle_inverse.argtypes = [_f64p, C.c_int, _f64p]
    L.bpmf_oracle_hyper.argtypes = [C.c_int, C.c_int, C.c_uint32, _f64p, _f64p, _f64p, _f64p, _f64p]
    L.bpmf_oracle_create.restype = C.c_void_p
    L.bpmf_oracle_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _i32p, _i32p, _f64p, C.c_int, C.c_int,
                                     C.c_int64, _i32p, _i32p, _f64p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
    L.bpmf_oracle_destroy.argtypes = [C.c_void_p]
    for name in ("num", "iter"):
        getattr(L, "bpmf_oracle_" + name).argtypes = [C.c_void_p, C.c_int]
    for name in ("nnz", "nnz_test"):
        f = getattr(L, "bpmf_oracle_" + name)
        f.argtypes = [C.c_void_p, C.c_int]
        f.restype = C.c_int64
    L.bpmf_oracle_mean_rating.argtypes = [C.c_void_p, C.c_int]
    L.bpmf_oracle_mean_rating.restype = C.c_double
    L.bpmf_oracle_sweep.argtypes = [C.c_void_p, C.c_int]
    L.bpmf_oracle_sample_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.bpmf_oracle_predict.argtypes = [C.c_void_p, C.c_int]
    L.bpmf_oracle_iterate.argtypes = [C.c_void_p]
    L.bpmf_oracle_finish.argtypes = [C.c_void_p]
    L.bpmf_oracle_get_items.argtypes = [C.c_void_p, C.c_int, _f64p]
    L.bpmf_oracle_set_items.argtypes = [C.c_void_p, C.c_int, _f64p]
    L.bpmf_oracle_set_iter.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.bpmf_oracle_get_hyper.argtypes = [C.c_void_p, C.c_int, _f64p, _f64p, _f64p]
    L.bpmf_oracle_set_hyper.argtypes = [C.c_void_p, C.c_int, _f64p, _f64p]
    L.bpmf_oracle_get_stats.argtypes = [C.c_void_p, C.c_int, _f64p, _f64p, _f64p, C.POINTER(C.c_double)]
    L.bpmf_oracle_set_cov.argtypes = [C.c_void_p, C.c_int, _f64p]
    L.bpmf_oracle_get_rmse.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_int64)]
    L.bpmf_oracle_get_pred.argtypes = [C.c_void_p, C.c_int, _f64p, _f64p]
    L.bpmf_oracle_get_csc.argtypes = [C.c_void_p, C.c_int, C.c_int, _i64p, _i32p, _f64p]
    L.bpmf_oracle_get_aggr.argtypes = [C.c_void_p, C.c_int, _f64p, _f64p]
    L.bpmf_oracle_set_prop.argtypes = [C.c_void_p, C.c_int, _f64p, _f64p]
    _lib = L
    return L


def philox4x32_10(ctr, key):
    out = np.zeros(4, np.uint32)
    lib().bpmf_oracle_philox4x32_10(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out)
    return out


def words(c, n):
    out = np.zeros(n, np.uint32)
    lib().bpmf_oracle_words(c, n, out)
    return out


def randn(c, n):
    out = np.zeros(n, np.float64)
    lib().bpmf_oracle_randn(c, n, out)
    return out


def gamma_then_randn(c, alpha, n):
    g = C.c_double()
    out = np.zeros(max(n, 1), np.float64)
    lib().bpmf_oracle_gamma_then_randn(c, alpha, n, C.byref(g), out)
    return g.value, out[:n]


def hyper(K, N, it, cov, s=None):
    """rng_set_pos(it); hp.sample(N, sum, cov) -> (mu, LambdaU, LambdaF) with K x K matrices column-major."""
    s = np.zeros(K) if s is None else np.ascontiguousarray(s, np.float64)
    mu = np.zeros(K)
    LU = np.zeros(K * K)
    LF = np.zeros(K * K)
    rc = lib().bpmf_oracle_hyper(K, N, it, s, np.ascontiguousarray(cov, np.float64).ravel(), mu, LU, LF)
    if rc:
        raise RuntimeError(lib().bpmf_oracle_last_error().decode())
    return mu, LU, LF


MOVIES, USERS = 0, 1


class Oracle:
    """The two-factor model; rows of the input = users, columns = movies (sample.cpp:112-137)."""

    def __init__(self, K, shape, rows, cols, vals, tshape, trows, tcols, tvals, alpha=2.0, burnin=5, nthreads=0,
                 keep_aggr=False, no_covariance=False):
        L = lib()
        self.K = K
        self._h = L.bpmf_oracle_create(
            K, shape[0], shape[1], len(vals), np.ascontiguousarray(rows, np.int32), np.ascontiguousarray(cols, np.int32),
            np.ascontiguousarray(vals, np.float64), tshape[0], tshape[1], len(tvals),
            np.ascontiguousarray(trows, np.int32), np.ascontiguousarray(tcols, np.int32),
            np.ascontiguousarray(tvals, np.float64), alpha, burnin, nthreads, int(keep_aggr), int(no_covariance))
        if not self._h:
            raise RuntimeError(L.bpmf_oracle_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().bpmf_oracle_destroy(self._h)
            self._h = None

    def _ck(self, rc):
        if rc:
            raise RuntimeError(lib().bpmf_oracle_last_error().decode())

    def num(self, side): return lib().bpmf_oracle_num(self._h, side)
    def nnz(self, side): return lib().bpmf_oracle_nnz(self._h, side)
    def nnz_test(self, side): return lib().bpmf_oracle_nnz_test(self._h, side)
    def mean_rating(self, side): return lib().bpmf_oracle_mean_rating(self._h, side)
    def iter(self, side): return lib().bpmf_oracle_iter(self._h, side)
    def sweep(self, side): self._ck(lib().bpmf_oracle_sweep(self._h, side))
    def sample_range(self, side, lo, hi): self._ck(lib().bpmf_oracle_sample_range(self._h, side, lo, hi))
    def predict(self, side): self._ck(lib().bpmf_oracle_predict(self._h, side))
    def iterate(self): self._ck(lib().bpmf_oracle_iterate(self._h))
    def finish(self): self._ck(lib().bpmf_oracle_finish(self._h))
    def set_iter(self, side, it): lib().bpmf_oracle_set_iter(self._h, side, it)

    def items(self, side):
        out = np.zeros((self.num(side), self.K))
        lib().bpmf_oracle_get_items(self._h, side, out.reshape(-1))
        return out  # [item, k]

    def set_items(self, side, a):
        a = np.ascontiguousarray(a, np.float64)
        assert a.shape == (self.num(side), self.K)
        lib().bpmf_oracle_set_items(self._h, side, a.reshape(-1))

    def hyper(self, side):
        K = self.K
        mu, LU, LF = np.zeros(K), np.zeros(K * K), np.zeros(K * K)
        lib().bpmf_oracle_get_hyper(self._h, side, mu, LU, LF)
        return mu, LU, LF

    def set_hyper(self, side, mu, LF):
        lib().bpmf_oracle_set_hyper(self._h, side, np.ascontiguousarray(mu, np.float64),
                                    np.ascontiguousarray(LF, np.float64).reshape(-1))

    def stats(self, side):
        K = self.K
        s, p, c = np.zeros(K), np.zeros(K * K), np.zeros(K * K)
        n = C.c_double()
        lib().bpmf_oracle_get_stats(self._h, side, s, p, c, C.byref(n))
        return s, p, c, n.value

    def set_cov(self, side, cov):
        lib().bpmf_oracle_set_cov(self._h, side, np.ascontiguousarray(cov, np.float64).reshape(-1))

    def rmse(self, side):
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        lib().bpmf_oracle_get_rmse(self._h, side, C.byref(a), C.byref(b), C.byref(n))
        return a.value, b.value, n.value

    def pred(self, side):
        n = self.nnz_test(side)
        a, b = np.zeros(n), np.zeros(n)
        lib().bpmf_oracle_get_pred(self._h, side, a, b)
        return a, b

    def csc(self, side, which=0):
        n = self.nnz(side) if which == 0 else self.nnz_test(side)
        colptr = np.zeros(self.num(side) + 1, np.int64)
        rowidx = np.zeros(n, np.int32)
        val = np.zeros(n, np.float64)
        lib().bpmf_oracle_get_csc(self._h, side, which, colptr, rowidx, val)
        return colptr, rowidx, val

    def set_prop(self, side, mu, lam):
        """propagated posterior of -m / -l: mu [item, K], lam [item, K*K] (column-major K x K per item)"""
        mu = np.ascontiguousarray(mu, np.float64)
        lam = np.ascontiguousarray(lam, np.float64)
        assert mu.shape == (self.num(side), self.K) and lam.shape == (self.num(side), self.K * self.K)
        lib().bpmf_oracle_set_prop(self._h, side, mu.reshape(-1), lam.reshape(-1))

    def aggr(self, side):
        K, n = self.K, self.num(side)
        mu, lam = np.zeros((n, K)), np.zeros((n, K * K))
        lib().bpmf_oracle_get_aggr(self._h, side, mu.reshape(-1), lam.reshape(-1))
        return mu, lam
