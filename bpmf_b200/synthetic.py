"""Seeded synthetic rating matrices for the BASELINE.json configs.

Host-side input generation only (numpy/scipy); nothing here is on the sampling path. The generator is
deterministic in (shape, mean_nnz_per_row, seed), so bench.py's GPU arm, its CPU reference arm and the tests
all see the same matrix.

    R (nrows = users, ncols = movies): per-row nnz ~ Poisson(mean), uniformly random distinct columns,
    r_ij = u*_i . v*_j + 3.5 + N(0, 0.5^2) with u*, v* ~ N(0, I_rank / 4)  (planted rank-`rank` model)

Both compressed forms the two sweeps need are returned:
    users  side (BPMF_GPU_USERS):  column i of R^T = row i of R   -> (u_ptr, u_idx = movie ids, u_val)
    movies side (BPMF_GPU_MOVIES): column j of R                  -> (m_ptr, m_idx = user ids,  m_val)
exactly the `Sys::M` of each side (c++/sample.cpp:112-137): int32 inner indices ascending.
"""
import os
import sys
import time

import numpy as np

WORKLOADS = {
    # name: (nrows(users), ncols(movies), mean nnz per row, K, seed)      BASELINE.json configs
    "synthA-1Mx1M-100Mnnz-K32": (1_000_000, 1_000_000, 100.0, 32, 20260001),
    "synthB-200Kx200K-50Mnnz-K128": (200_000, 200_000, 250.0, 128, 20260002),
    "ml1m-shaped-6040x3952-1Mnnz-K32": (6040, 3952, 165.6, 32, 20260003),
    "small-20Kx20K-1Mnnz-K32": (20_000, 20_000, 50.0, 32, 20260004),
    # the skewed variant of Synthetic A: movie popularity ~ Zipf(s = 1) over a random permutation of the ids
    "synthA-zipf-1Mx1M-K32": (1_000_000, 1_000_000, 100.0, 32, 20260005),
    "small-zipf-50Kx50K-K32": (50_000, 50_000, 60.0, 32, 20260006),
}
ZIPF = {"synthA-zipf-1Mx1M-K32": 1.0, "small-zipf-50Kx50K-K32": 1.0}


class Ratings:
    """Train matrix in both orientations + a test set in coordinate form."""
    __slots__ = ("nrows", "ncols", "nnz", "mean_rating", "u_ptr", "u_idx", "u_val", "m_ptr", "m_idx", "m_val",
                 "t_rows", "t_cols", "t_vals")

    def side(self, side):
        """(num_items, num_other, colptr, rowidx, val) of `side` (0 = movies, 1 = users) as bpmf_gpu_load_side wants."""
        if side == 0:
            return self.ncols, self.nrows, self.m_ptr, self.m_idx, self.m_val
        return self.nrows, self.ncols, self.u_ptr, self.u_idx, self.u_val

    def test_side(self, side):
        """Test matrix T of `side` in compressed-column form (sample.cpp:117-123)."""
        n = self.ncols if side == 0 else self.nrows
        col = self.t_cols if side == 0 else self.t_rows
        row = self.t_rows if side == 0 else self.t_cols
        order = np.lexsort((row, col))
        ptr = np.zeros(n + 1, np.int64)
        np.cumsum(np.bincount(col, minlength=n), out=ptr[1:])
        return ptr, row[order].astype(np.int32), self.t_vals[order].astype(np.float64)


def _planted_values(rng, rows, cols, U, V, chunk=4_000_000):
    out = np.empty(len(rows), np.float64)
    for s in range(0, len(rows), chunk):
        e = min(len(rows), s + chunk)
        out[s:e] = np.einsum("ij,ij->i", U[rows[s:e]], V[cols[s:e]])
    out += 3.5
    out += rng.standard_normal(len(rows), dtype=np.float32) * 0.5
    return out


def generate(nrows, ncols, mean_nnz_row, seed, rank=16, test_frac=0.01, verbose=False, zipf=None):
    t0 = time.time()
    rng = np.random.Generator(np.random.PCG64(seed))
    counts = rng.poisson(mean_nnz_row, nrows).astype(np.int64)
    total = int(counts.sum())
    key = np.repeat(np.arange(nrows, dtype=np.int64), counts)
    key *= ncols
    if zipf is None:
        key += rng.integers(0, ncols, total, dtype=np.int64)
    else:
        # column popularity ~ 1 / rank^s, ranks assigned to a random permutation of the columns; duplicates inside a row
        # are dropped below, so the hottest columns saturate (rated by nearly every row) and nnz ends below the nominal
        cdf = np.cumsum(1.0 / np.arange(1, ncols + 1, dtype=np.float64) ** zipf)
        cdf /= cdf[-1]
        perm = rng.permutation(ncols).astype(np.int64)
        for s0 in range(0, total, 8_000_000):
            e0 = min(total, s0 + 8_000_000)
            rk = np.searchsorted(cdf, rng.random(e0 - s0), side="right")
            np.minimum(rk, ncols - 1, out=rk)
            key[s0:e0] += perm[rk]
    key.sort()
    keep = np.empty(total, bool)
    keep[0] = True
    np.not_equal(key[1:], key[:-1], out=keep[1:])
    key = key[keep]                      # distinct (row, col), row-major order, columns ascending in a row
    rows = (key // ncols).astype(np.int32)
    cols = (key - rows.astype(np.int64) * ncols).astype(np.int32)
    del key, keep
    U = (rng.standard_normal((nrows, rank), dtype=np.float32) * 0.5)
    V = (rng.standard_normal((ncols, rank), dtype=np.float32) * 0.5)
    vals = _planted_values(rng, rows, cols, U, V)
    r = Ratings()
    r.nrows, r.ncols, r.nnz = nrows, ncols, len(vals)
    r.mean_rating = float(vals.sum() / len(vals))      # Sys::init: M.sum() / M.nonZeros() (sample.cpp:183)
    r.u_ptr = np.zeros(nrows + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=nrows), out=r.u_ptr[1:])
    r.u_idx, r.u_val = cols, vals
    if verbose:
        print("  [synthetic] %d x %d, %d nnz by rows in %.1fs" % (nrows, ncols, r.nnz, time.time() - t0), flush=True, file=sys.stderr)
    # the other orientation: scipy's csr->csc is O(nnz) and keeps row ids ascending inside a column
    import scipy.sparse as sp
    csr = sp.csr_matrix((vals, cols, r.u_ptr.astype(np.int64)), shape=(nrows, ncols))
    csc = csr.tocsc()
    r.m_ptr = csc.indptr.astype(np.int64)
    r.m_idx = csc.indices.astype(np.int32)
    r.m_val = np.ascontiguousarray(csc.data, np.float64)
    del csr, csc
    nt = max(1, int(r.nnz * test_frac))
    r.t_rows = rng.integers(0, nrows, nt).astype(np.int32)
    r.t_cols = rng.integers(0, ncols, nt).astype(np.int32)
    r.t_vals = _planted_values(rng, r.t_rows, r.t_cols, U, V)
    if verbose:
        print("  [synthetic] both orientations + %d test entries in %.1fs" % (nt, time.time() - t0), flush=True, file=sys.stderr)
    return r


_FIELDS = ("u_ptr", "u_idx", "u_val", "m_ptr", "m_idx", "m_val", "t_rows", "t_cols", "t_vals")


def save(r, d):
    os.makedirs(d, exist_ok=True)
    for f in _FIELDS:
        np.save(os.path.join(d, f + ".npy"), getattr(r, f))
    np.save(os.path.join(d, "meta.npy"), np.array([r.nrows, r.ncols, r.nnz, r.mean_rating], np.float64))
    open(os.path.join(d, "DONE"), "w").write("ok\n")


def load(d, mmap=True):
    if not os.path.exists(os.path.join(d, "DONE")):
        return None
    r = Ratings()
    m = np.load(os.path.join(d, "meta.npy"))
    r.nrows, r.ncols, r.nnz, r.mean_rating = int(m[0]), int(m[1]), int(m[2]), float(m[3])
    for f in _FIELDS:
        setattr(r, f, np.load(os.path.join(d, f + ".npy"), mmap_mode="r" if mmap else None))
    return r


def workload(name, cache_dir=None, verbose=False):
    """The named BASELINE workload -> (Ratings, K). cache_dir (e.g. /dev/shm/...) lets the ranks of one box and the
    two bench arms share one generation."""
    nrows, ncols, mean, K, seed = WORKLOADS[name]
    if cache_dir:
        d = os.path.join(cache_dir, "bpmf_b200_%s_%d" % (name, seed))
        r = load(d)
        if r is not None:
            return r, K
    r = generate(nrows, ncols, mean, seed, verbose=verbose, zipf=ZIPF.get(name))
    if cache_dir:
        try:
            save(r, d)
        except OSError:
            pass
    return r, K
