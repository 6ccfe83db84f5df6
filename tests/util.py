"""Shared helpers for the tests: seeded synthetic ratings and oracle <-> GPU pairing."""
import numpy as np

MOVIES, USERS = 0, 1

# data/tiny of the reference (4 users x 2 movies, 6 train / 2 test ratings; data/tiny/train.mtx, test.mtx),
# 0-based (row=user, col=movie, value)
TINY_TRAIN = ((4, 2), [0, 1, 2, 3, 0, 2], [0, 0, 0, 0, 1, 1], [2., 3., 7., 4., 5., 1.])
TINY_TEST = ((4, 2), [1, 3], [1, 1], [5., 1.])


def synth_ratings(nrows, ncols, nnz, seed, rank=8, skew=0.0, test_frac=0.1, empty_rows=0, heavy_col=0):
    """Planted low-rank ratings r = u.v + 3.5 + noise on a random pattern.

    skew > 0 draws columns from a Zipf-like popularity; empty_rows leaves the last rows without ratings;
    heavy_col makes column 0 receive that many extra ratings (a ChEMBL-style hot column)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    live_rows = nrows - empty_rows
    rows = rng.integers(0, live_rows, size=nnz)
    if skew > 0:
        w = 1.0 / np.arange(1, ncols + 1) ** skew
        cols = rng.choice(ncols, size=nnz, p=w / w.sum())
    else:
        cols = rng.integers(0, ncols, size=nnz)
    if heavy_col:
        extra = rng.choice(live_rows, size=min(heavy_col, live_rows), replace=False)
        rows = np.concatenate([rows, extra])
        cols = np.concatenate([cols, np.zeros(len(extra), dtype=cols.dtype)])
    key = rows.astype(np.int64) * ncols + cols
    _, first = np.unique(key, return_index=True)   # distinct (row, col) pairs
    first.sort()
    rows, cols = rows[first], cols[first]
    U = rng.normal(0, 0.5, size=(nrows, rank))
    V = rng.normal(0, 0.5, size=(ncols, rank))
    vals = np.einsum("ij,ij->i", U[rows], V[cols]) + 3.5 + rng.normal(0, 0.5, size=len(rows))
    ntest = max(1, int(len(rows) * test_frac))
    perm = rng.permutation(len(rows))
    te, tr = perm[:ntest], perm[ntest:]
    train = ((nrows, ncols), rows[tr].astype(np.int32), cols[tr].astype(np.int32), vals[tr].copy())
    test = ((nrows, ncols), rows[te].astype(np.int32), cols[te].astype(np.int32), vals[te].copy())
    return train, test


def make_oracle(K, train, test, **kw):
    from oracle import oracle as o
    (s, r, c, v), (ts, tr, tc, tv) = train, test
    return o.Oracle(K, s, np.asarray(r, np.int32), np.asarray(c, np.int32), np.asarray(v, np.float64), ts,
                    np.asarray(tr, np.int32), np.asarray(tc, np.int32), np.asarray(tv, np.float64), **kw)


def make_gpu_from_oracle(orc, K, device=0, heavy_threshold=None):
    """A GPU context loaded with exactly the CSC structure / mean ratings the oracle built."""
    import bpmf_b200
    ctx = bpmf_b200.Context(K, device)
    if heavy_threshold:
        ctx.set_heavy_threshold(heavy_threshold)
    for side in (MOVIES, USERS):
        colptr, rowidx, val = orc.csc(side, 0)
        ctx.load_side(side, orc.num(side), orc.num(1 - side), colptr, rowidx, val, orc.mean_rating(side))
    for side in (MOVIES, USERS):
        colptr, rowidx, val = orc.csc(side, 1)
        ctx.load_test(side, colptr, rowidx, val)
    return ctx


def random_spd(K, seed, scale=1.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    A = rng.normal(size=(K, 2 * K))
    return scale * (A @ A.T) / (2 * K)
