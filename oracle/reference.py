"""ctypes wrapper of oracle/_ref/libbpmf_ref_k<K>.so: the REFERENCE's own hot-path sources (c++/sample.cpp,
c++/mvnormal.cpp) compiled unmodified against the stand-in headers of oracle/shim/ (oracle/Makefile, target `ref`;
oracle/ref_harness.cpp is the C API). Test infrastructure only — used by tests/test_oracle_vs_reference.py to pin the
restatement in bpmf_oracle.hpp. The libraries are built here when /root/reference exists and travel to the GPU box
prebuilt (oracle/_ref/ is git-ignored, not gpurun-ignored)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("BPMF_REFERENCE_SRC", "/root/reference/c++")
KS = (10, 16, 32, 48, 64, 128)

_f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def so_path(K):
    return os.path.join(HERE, "_ref", "libbpmf_ref_k%d.so" % K)


def build(force=False):
    """Compile the reference sources for every K in KS. Returns False when there are no reference sources here."""
    if not os.path.exists(os.path.join(REF_SRC, "sample.cpp")):
        return False
    if force:
        subprocess.check_call(["make", "-C", HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HERE, "ref-core", "REF=" + REF_SRC], stdout=subprocess.DEVNULL)
    # the reference executable with the B200 back end also needs the product library: best effort
    subprocess.call(["make", "-C", HERE, "-k", "ref-cuda", "REF=" + REF_SRC], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return True


def available(K):
    return os.path.exists(so_path(K)) or (K in KS and build())


_libs = {}


def lib(K):
    if K in _libs:
        return _libs[K]
    if not os.path.exists(so_path(K)):
        if not build() or not os.path.exists(so_path(K)):
            raise RuntimeError("no reference build for K=%d (needs %s or a prebuilt %s)" % (K, REF_SRC, so_path(K)))
    L = C.CDLL(so_path(K))
    L.bpmf_ref_create.restype = C.c_void_p
    L.bpmf_ref_create.argtypes = [C.c_int, C.c_int, C.c_long, _i32, _i32, _f64, C.c_long, _i32, _i32, _f64, C.c_int, C.c_int,
                                  C.c_double, C.c_int]
    L.bpmf_ref_error.restype = C.c_char_p
    L.bpmf_ref_error.argtypes = [C.c_void_p]
    L.bpmf_ref_destroy.argtypes = [C.c_void_p]
    for name in ("bpmf_ref_sample",):
        getattr(L, name).argtypes = [C.c_void_p, C.c_int]
    L.bpmf_ref_predict.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.bpmf_ref_num.argtypes = [C.c_void_p, C.c_int]
    L.bpmf_ref_nnz_test.argtypes = [C.c_void_p, C.c_int]
    L.bpmf_ref_nnz_test.restype = C.c_long
    for name in ("bpmf_ref_get_items", "bpmf_ref_set_items", "bpmf_ref_get_scalars", "bpmf_ref_get_cov"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_int, _f64]
    L.bpmf_ref_get_hyper.argtypes = [C.c_void_p, C.c_int, _f64, _f64, _f64]
    L.bpmf_ref_get_predictions.argtypes = [C.c_void_p, C.c_int, _f64, _f64]
    L.bpmf_ref_get_aggregates.argtypes = [C.c_void_p, C.c_int, _f64, _f64]
    L.bpmf_ref_set_prop.argtypes = [C.c_void_p, C.c_int, _f64, _f64]
    L.bpmf_ref_randn.argtypes = [C.c_uint, C.c_int, _f64]
    assert L.bpmf_ref_num_latent() == K
    _libs[K] = L
    return L


def randn(K, c, n):
    out = np.zeros(n)
    lib(K).bpmf_ref_randn(c, n, out)
    return out


class Reference:
    """movies + users `Sys` objects of the reference; rows of the input = users, columns = movies."""

    def __init__(self, K, shape, rows, cols, vals, trows, tcols, tvals, alpha=2.0, burnin=5, nsims=20, keep_aggr=False):
        self.K, self.L = K, lib(K)
        self._h = self.L.bpmf_ref_create(shape[0], shape[1], len(vals), np.ascontiguousarray(rows, np.int32),
                                         np.ascontiguousarray(cols, np.int32), np.ascontiguousarray(vals, np.float64), len(tvals),
                                         np.ascontiguousarray(trows, np.int32), np.ascontiguousarray(tcols, np.int32),
                                         np.ascontiguousarray(tvals, np.float64), burnin, nsims, alpha, int(keep_aggr))
        err = self.L.bpmf_ref_error(self._h).decode()
        if err:
            raise RuntimeError(err)

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.bpmf_ref_destroy(self._h)
            self._h = None

    def _ck(self, rc):
        if rc:
            raise RuntimeError(self.L.bpmf_ref_error(self._h).decode())

    def num(self, side): return self.L.bpmf_ref_num(self._h, side)
    def sample(self, side): self._ck(self.L.bpmf_ref_sample(self._h, side))
    def predict(self, side, all=False): self._ck(self.L.bpmf_ref_predict(self._h, side, int(all)))

    def items(self, side):
        out = np.zeros((self.num(side), self.K))
        self.L.bpmf_ref_get_items(self._h, side, out.reshape(-1))
        return out

    def set_items(self, side, a):
        a = np.ascontiguousarray(a, np.float64)
        assert a.shape == (self.num(side), self.K)
        self.L.bpmf_ref_set_items(self._h, side, a.reshape(-1))

    def scalars(self, side):
        out = np.zeros(6)
        self.L.bpmf_ref_get_scalars(self._h, side, out)
        return dict(rmse=out[0], rmse_avg=out[1], norm=out[2], mean_rating=out[3], iter=int(out[4]), num_predict=int(out[5]))

    def hyper(self, side):
        K = self.K
        mu, LU, LF = np.zeros(K), np.zeros(K * K), np.zeros(K * K)
        self.L.bpmf_ref_get_hyper(self._h, side, mu, LU, LF)
        return mu, LU, LF

    def cov(self, side):
        c = np.zeros(self.K * self.K)
        self.L.bpmf_ref_get_cov(self._h, side, c)
        return c

    def pred(self, side):
        n = self.L.bpmf_ref_nnz_test(self._h, side)
        a, b = np.zeros(n), np.zeros(n)
        self.L.bpmf_ref_get_predictions(self._h, side, a, b)
        return a, b

    def aggr(self, side):
        K, n = self.K, self.num(side)
        mu, lam = np.zeros((n, K)), np.zeros((n, K * K))
        self._ck(self.L.bpmf_ref_get_aggregates(self._h, side, mu.reshape(-1), lam.reshape(-1)))
        return mu, lam

    def set_prop(self, side, mu, lam):
        mu = np.ascontiguousarray(mu, np.float64)
        lam = np.ascontiguousarray(lam, np.float64)
        assert mu.shape == (self.num(side), self.K) and lam.shape == (self.num(side), self.K * self.K)
        self.L.bpmf_ref_set_prop(self._h, side, mu.reshape(-1), lam.reshape(-1))
