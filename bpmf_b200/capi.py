"""ctypes binding of libbpmf_b200.so — one method per entry point of include/bpmf_gpu.h.

There is no fallback: if the shared library is missing or no sm_100 GPU is present this raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("BPMF_B200_LIB") or os.path.join(_HERE, "libbpmf_b200.so")   # (the override: probe builds of bench_micro/)

MOVIES, USERS = 0, 1
KERNEL_AUTO, KERNEL_EXACT, KERNEL_STREAM, KERNEL_BLOCK = 0, 1, 3, 4
STREAM_KERNEL_NAME = "items_stream32v3_kernel<2,20>"   # what KERNEL_AUTO launches at K = 32 (csrc/stream_kernel.cu)

_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_opt = C.c_void_p  # nullable pointer arguments

# every symbol include/bpmf_gpu.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "bpmf_gpu_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "bpmf_gpu_destroy": (C.c_int, [C.c_void_p]),
    "bpmf_gpu_last_error": (C.c_char_p, [C.c_void_p]),
    "bpmf_gpu_num_latent": (C.c_int, [C.c_void_p]),
    "bpmf_gpu_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bpmf_gpu_sync": (C.c_int, [C.c_void_p]),
    "bpmf_gpu_load_side": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _i64p, _i32p, _f64p, C.c_double]),
    "bpmf_gpu_set_heavy_threshold": (C.c_int, [C.c_void_p, C.c_int64]),
    "bpmf_gpu_load_test": (C.c_int, [C.c_void_p, C.c_int, _i64p, _i32p, _f64p]),
    "bpmf_gpu_load_coo": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _i32p, _f64p]),
    "bpmf_gpu_load_test_coo": (C.c_int, [C.c_void_p, C.c_int64, _i32p, _i32p, _f64p]),
    "bpmf_gpu_get_side": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "bpmf_gpu_set_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "bpmf_gpu_bind_items": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "bpmf_gpu_set_peers": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "bpmf_gpu_items_device_ptr": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "bpmf_gpu_ipc_export": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p]),
    "bpmf_gpu_ipc_open": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "bpmf_gpu_enable_peer_access": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bpmf_gpu_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    "bpmf_gpu_host_free": (C.c_int, [C.c_void_p]),
    "bpmf_gpu_set_items": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "bpmf_gpu_get_items": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "bpmf_gpu_set_items_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "bpmf_gpu_get_items_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "bpmf_gpu_push_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "bpmf_gpu_get_iter": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "bpmf_gpu_set_iter": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "bpmf_gpu_sample": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int]),
    "bpmf_gpu_sample_host": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]),
    "bpmf_gpu_sample_hyper": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, _opt, _opt]),
    "bpmf_gpu_set_hyper": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p]),
    "bpmf_gpu_get_hyper": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p, _f64p]),
    "bpmf_gpu_sample_items": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_double, C.c_int]),
    "bpmf_gpu_reduce_stats": (C.c_int, [C.c_void_p, C.c_int]),
    "bpmf_gpu_get_stats": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p, _f64p, C.POINTER(C.c_double)]),
    "bpmf_gpu_predict": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.POINTER(C.c_int64)]),
    "bpmf_gpu_get_predictions": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p]),
    "bpmf_gpu_set_prop_posterior": (C.c_int, [C.c_void_p, C.c_int, _opt, _opt]),
    "bpmf_gpu_enable_aggregation": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "bpmf_gpu_aggregate": (C.c_int, [C.c_void_p, C.c_int]),
    "bpmf_gpu_get_aggregates": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p]),
    "bpmf_gpu_launch_count": (C.c_int64, [C.c_void_p]),
    "bpmf_gpu_last_items_kernel_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "bpmf_gpu_items_kernel_time": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "bpmf_gpu_debug_set_tuning": (C.c_int, [C.c_void_p, C.c_int]),
    "bpmf_gpu_upload_push_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "bpmf_gpu_sample_host_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]),
    "bpmf_gpu_sample_host_end": (C.c_int, [C.c_void_p, C.c_int]),
    "bpmf_gpu_load_side_slice": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i64p, _i32p, _f64p, C.c_double]),
    "bpmf_gpu_finalize_aggregates": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "bpmf_gpu_stats_block_items_for": (C.c_int, [C.c_int, C.c_int]),
    "bpmf_gpu_peer_barrier": (C.c_int, [C.c_void_p, C.c_int]),
    "bpmf_gpu_reduce_stats_partial": (C.c_int, [C.c_void_p, C.c_int]),
    "bpmf_gpu_reduce_stats_final": (C.c_int, [C.c_void_p, C.c_int]),
    "bpmf_gpu_stats_block_items": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "bpmf_gpu_stats_device_ptr": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "bpmf_gpu_ipc_export_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p]),
    "bpmf_gpu_set_stats_peers": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "bpmf_gpu_debug_set_roles": (C.c_int, [C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_int]),
    "bpmf_gpu_debug_block_schedule": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "bpmf_gpu_debug_randn": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, _f64p]),
}

_lib = None


class BpmfGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("bpmf_gpu error %d: %s" % (code, msg))
        self.code = code


def stats_block_items_for(num_latent, num_items):
    """items per statistics block of a side with num_items items (the granularity of multi-GPU item ranges)"""
    n = load_library().bpmf_gpu_stats_block_items_for(num_latent, num_items)
    if n < 1:
        raise ValueError("bad arguments")
    return n


def load_library():
    """dlopen the in-tree shared library and type every entry point. Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise FileNotFoundError(SO_PATH + " is missing: run `python -m bpmf_b200.build` (there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)  # AttributeError if the library does not export it
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU's BPMF state (both factors). Mirrors the call sequence of a `CUDA_Sys : Sys` backend."""

    def __init__(self, num_latent, device=0):
        self.L = load_library()
        self.K = int(num_latent)
        h = C.c_void_p()
        rc = self.L.bpmf_gpu_create(C.byref(h), device, self.K)
        if rc:
            raise BpmfGpuError(rc, self.L.bpmf_gpu_last_error(None).decode())
        self.h = h
        self.num = [0, 0]
        self.nnz_test = [0, 0]

    def close(self):
        if getattr(self, "h", None):
            self.L.bpmf_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise BpmfGpuError(rc, self.L.bpmf_gpu_last_error(self.h).decode())

    # ---- plumbing
    def set_stream(self, stream_handle): self._ck(self.L.bpmf_gpu_set_stream(self.h, C.c_void_p(stream_handle)))
    def sync(self): self._ck(self.L.bpmf_gpu_sync(self.h))
    def launch_count(self): return self.L.bpmf_gpu_launch_count(self.h)

    # ---- data
    def load_side(self, side, num_items, num_other, colptr, rowidx, val, mean_rating):
        colptr = np.ascontiguousarray(colptr, np.int64)
        rowidx = np.ascontiguousarray(rowidx, np.int32)
        val = np.ascontiguousarray(val, np.float64)
        assert colptr.shape == (num_items + 1,)
        self._ck(self.L.bpmf_gpu_load_side(self.h, side, num_items, num_other, colptr, rowidx, val, mean_rating))
        self.num[side] = num_items

    def load_side_slice(self, side, num_items, num_other, lo, hi, colptr, rowidx, val, mean_rating):
        """only the ratings of the items [lo, hi) become resident; colptr / rowidx / val are the FULL arrays of the side"""
        colptr = np.asarray(colptr)
        p0, p1 = int(colptr[lo]), int(colptr[hi])
        cs = np.ascontiguousarray(colptr[lo:hi + 1] - p0, np.int64)
        ri = np.ascontiguousarray(rowidx[p0:p1], np.int32) if p1 > p0 else np.zeros(1, np.int32)
        va = np.ascontiguousarray(val[p0:p1], np.float64) if p1 > p0 else np.zeros(1, np.float64)
        self._ck(self.L.bpmf_gpu_load_side_slice(self.h, side, num_items, num_other, lo, hi, cs, ri, va, float(mean_rating)))
        self.num[side] = num_items

    def finalize_aggregates(self, side, nsamples): self._ck(self.L.bpmf_gpu_finalize_aggregates(self.h, side, nsamples))

    def set_heavy_threshold(self, n): self._ck(self.L.bpmf_gpu_set_heavy_threshold(self.h, n))

    def load_test(self, side, colptr, rowidx, val):
        colptr = np.ascontiguousarray(colptr, np.int64)
        self._ck(self.L.bpmf_gpu_load_test(self.h, side, colptr, np.ascontiguousarray(rowidx, np.int32),
                                           np.ascontiguousarray(val, np.float64)))
        self.nnz_test[side] = int(colptr[-1])

    def load_coo(self, num_rows, num_cols, row, col, val):
        """both sides from one coordinate list, built on the device (rows = users, cols = movies)"""
        row, col = np.ascontiguousarray(row, np.int32), np.ascontiguousarray(col, np.int32)
        val = np.ascontiguousarray(val, np.float64)
        assert row.shape == col.shape == val.shape
        self._ck(self.L.bpmf_gpu_load_coo(self.h, num_rows, num_cols, len(val), row, col, val))
        self.num[0], self.num[1] = num_cols, num_rows

    def load_test_coo(self, row, col, val):
        row, col = np.ascontiguousarray(row, np.int32), np.ascontiguousarray(col, np.int32)
        val = np.ascontiguousarray(val, np.float64)
        self._ck(self.L.bpmf_gpu_load_test_coo(self.h, len(val), row, col, val))
        for side in (0, 1):
            self.nnz_test[side] = self.get_side(side, test=True, arrays=False)[0]

    def get_side(self, side, test=False, arrays=True):
        """(nnz, mean_rating, colptr, rowidx, val) of a side's (test) matrix as it sits on the device"""
        nnz, mean = C.c_int64(), C.c_double()
        self._ck(self.L.bpmf_gpu_get_side(self.h, side, int(test), C.byref(nnz), C.byref(mean), None, None, None))
        if not arrays:
            return nnz.value, mean.value
        colptr = np.empty(self.num[side] + 1, np.int64)
        rowidx, val = np.empty(nnz.value, np.int32), np.empty(nnz.value, np.float64)
        self._ck(self.L.bpmf_gpu_get_side(self.h, side, int(test), None, None, colptr.ctypes.data, rowidx.ctypes.data, val.ctypes.data))
        return nnz.value, mean.value, colptr, rowidx, val

    def set_range(self, side, lo, hi): self._ck(self.L.bpmf_gpu_set_range(self.h, side, lo, hi))
    def bind_items(self, side, dev_ptr): self._ck(self.L.bpmf_gpu_bind_items(self.h, side, C.c_void_p(dev_ptr)))

    def set_peers(self, side, ptrs):
        arr = (C.c_void_p * max(1, len(ptrs)))(*[C.c_void_p(p) for p in ptrs])
        self._ck(self.L.bpmf_gpu_set_peers(self.h, side, len(ptrs), arr))

    def items_device_ptr(self, side):
        p = C.c_void_p()
        self._ck(self.L.bpmf_gpu_items_device_ptr(self.h, side, C.byref(p)))
        return p.value

    def ipc_export(self, side):
        buf = C.create_string_buffer(64)
        self._ck(self.L.bpmf_gpu_ipc_export(self.h, side, buf))
        return buf.raw

    def ipc_open(self, handle):
        p = C.c_void_p()
        self._ck(self.L.bpmf_gpu_ipc_open(self.h, handle, C.byref(p)))
        return p.value

    def enable_peer_access(self, peer): self._ck(self.L.bpmf_gpu_enable_peer_access(self.h, peer.h))

    # ---- sliced sweep statistics (multi-GPU): every rank reduces the blocks of its own range into all ranks' buffers
    def stats_block_items(self, side):
        n = C.c_int()
        self._ck(self.L.bpmf_gpu_stats_block_items(self.h, side, C.byref(n)))
        return n.value

    def stats_device_ptr(self, side):
        p = C.c_void_p()
        self._ck(self.L.bpmf_gpu_stats_device_ptr(self.h, side, C.byref(p)))
        return p.value

    def ipc_export_stats(self, side):
        buf = C.create_string_buffer(64)
        self._ck(self.L.bpmf_gpu_ipc_export_stats(self.h, side, buf))
        return buf.raw

    def set_stats_peers(self, side, ptrs):
        arr = (C.c_void_p * max(1, len(ptrs)))(*[C.c_void_p(p) for p in ptrs])
        self._ck(self.L.bpmf_gpu_set_stats_peers(self.h, side, len(ptrs), arr))

    def peer_barrier(self, side): self._ck(self.L.bpmf_gpu_peer_barrier(self.h, side))
    def reduce_stats_partial(self, side): self._ck(self.L.bpmf_gpu_reduce_stats_partial(self.h, side))
    def reduce_stats_final(self, side): self._ck(self.L.bpmf_gpu_reduce_stats_final(self.h, side))

    def set_items(self, side, a):
        a = np.ascontiguousarray(a, np.float64)
        assert a.size == self.K * self.num[side]
        self._ck(self.L.bpmf_gpu_set_items(self.h, side, C.c_void_p(a.ctypes.data)))

    def set_items_ptr(self, side, host_ptr):
        """upload K * num doubles from a raw host address (e.g. a pinned torch tensor's data_ptr())"""
        self._ck(self.L.bpmf_gpu_set_items(self.h, side, C.c_void_p(host_ptr)))

    def get_items_ptr(self, side, host_ptr):
        self._ck(self.L.bpmf_gpu_get_items(self.h, side, C.c_void_p(host_ptr)))

    def get_items(self, side):
        out = np.empty((self.num[side], self.K), np.float64)
        self._ck(self.L.bpmf_gpu_get_items(self.h, side, C.c_void_p(out.ctypes.data)))
        return out

    def set_items_range_ptr(self, side, lo, hi, host_ptr):
        """upload items [lo, hi) from a raw host address of the FULL K x num matrix"""
        self._ck(self.L.bpmf_gpu_set_items_range(self.h, side, lo, hi, C.c_void_p(host_ptr)))

    def get_items_range_ptr(self, side, lo, hi, host_ptr):
        self._ck(self.L.bpmf_gpu_get_items_range(self.h, side, lo, hi, C.c_void_p(host_ptr)))

    def push_range(self, side, lo, hi): self._ck(self.L.bpmf_gpu_push_range(self.h, side, lo, hi))

    def get_iter(self, side):
        i = C.c_int()
        self._ck(self.L.bpmf_gpu_get_iter(self.h, side, C.byref(i)))
        return i.value

    def set_iter(self, side, it): self._ck(self.L.bpmf_gpu_set_iter(self.h, side, it))

    # ---- hot path
    def sample(self, side, alpha=2.0, variant=KERNEL_AUTO): self._ck(self.L.bpmf_gpu_sample(self.h, side, alpha, variant))

    def sample_host(self, side, host_other_ptr, host_items_ptr, alpha=2.0, variant=KERNEL_AUTO):
        """Sys::sample(other) with host-resident items on both sides; arguments are raw host addresses (or None)."""
        self._ck(self.L.bpmf_gpu_sample_host(self.h, side, alpha, variant, C.c_void_p(host_other_ptr),
                                             C.c_void_p(host_items_ptr)))

    def upload_push_range(self, side, lo, hi, host_ptr):
        self._ck(self.L.bpmf_gpu_upload_push_range(self.h, side, lo, hi, C.c_void_p(host_ptr)))

    def sample_host_begin(self, side, host_items_ptr, alpha=2.0, variant=KERNEL_AUTO):
        self._ck(self.L.bpmf_gpu_sample_host_begin(self.h, side, alpha, variant, C.c_void_p(host_items_ptr)))

    def sample_host_end(self, side): self._ck(self.L.bpmf_gpu_sample_host_end(self.h, side))

    def sample_hyper(self, side, it, sum_=None, cov=None):
        s = None if sum_ is None else np.ascontiguousarray(sum_, np.float64)
        c = None if cov is None else np.ascontiguousarray(cov, np.float64)
        self._ck(self.L.bpmf_gpu_sample_hyper(self.h, side, it, _ptr(s), _ptr(c)))

    def set_hyper(self, side, mu, LambdaF):
        self._ck(self.L.bpmf_gpu_set_hyper(self.h, side, np.ascontiguousarray(mu, np.float64),
                                           np.ascontiguousarray(LambdaF, np.float64).reshape(-1)))

    def get_hyper(self, side):
        K = self.K
        mu, LU, LF = np.empty(K), np.empty(K * K), np.empty(K * K)
        self._ck(self.L.bpmf_gpu_get_hyper(self.h, side, mu, LU, LF))
        return mu, LU, LF

    def sample_items(self, side, it, alpha=2.0, variant=KERNEL_AUTO):
        self._ck(self.L.bpmf_gpu_sample_items(self.h, side, it, alpha, variant))

    def reduce_stats(self, side): self._ck(self.L.bpmf_gpu_reduce_stats(self.h, side))

    def get_stats(self, side):
        K = self.K
        s, p, c = np.empty(K), np.empty(K * K), np.empty(K * K)
        n = C.c_double()
        self._ck(self.L.bpmf_gpu_get_stats(self.h, side, s, p, c, C.byref(n)))
        return s, p, c, n.value

    def predict(self, side, burnin):
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        self._ck(self.L.bpmf_gpu_predict(self.h, side, burnin, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def get_predictions(self, side):
        n = self.nnz_test[side]
        a, b = np.empty(n), np.empty(n)
        self._ck(self.L.bpmf_gpu_get_predictions(self.h, side, a, b))
        return a, b

    def set_prop_posterior(self, side, mu, lam):
        """per-item priors of -m / -l: mu [item, K] (unused by the draw, like the reference), lam [item, K*K]"""
        mu = None if mu is None else np.ascontiguousarray(mu, np.float64)
        lam = None if lam is None else np.ascontiguousarray(lam, np.float64)
        self._ck(self.L.bpmf_gpu_set_prop_posterior(self.h, side, _ptr(mu), _ptr(lam)))

    def enable_aggregation(self, side, burnin): self._ck(self.L.bpmf_gpu_enable_aggregation(self.h, side, burnin))
    def aggregate(self, side): self._ck(self.L.bpmf_gpu_aggregate(self.h, side))

    def get_aggregates(self, side):
        K, n = self.K, self.num[side]
        mu, lam = np.zeros((n, K)), np.zeros((n, K * K))      # only the columns of the aggregation range are written
        self._ck(self.L.bpmf_gpu_get_aggregates(self.h, side, mu.reshape(-1), lam.reshape(-1)))
        return mu, lam

    def last_items_kernel_ms(self):
        ms = C.c_float()
        self._ck(self.L.bpmf_gpu_last_items_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def items_kernel_time(self):
        """(total ms, count) of the item kernels launched since the last call."""
        t, n = C.c_double(), C.c_int()
        self._ck(self.L.bpmf_gpu_items_kernel_time(self.h, C.byref(t), C.byref(n)))
        return t.value, n.value

    def set_tuning(self, cfg): self._ck(self.L.bpmf_gpu_debug_set_tuning(self.h, cfg))
    def set_roles(self, gram_mask, stages, slots, warps=20): self._ck(self.L.bpmf_gpu_debug_set_roles(self.h, gram_mask, stages, slots, warps))

    def debug_randn(self, c, n):
        out = np.empty(n)
        self._ck(self.L.bpmf_gpu_debug_randn(self.h, c, n, out))
        return out
