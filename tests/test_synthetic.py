"""The synthetic workload generator (bench.py's inputs): deterministic in its seed, both orientations consistent,
indices ascending inside a column, and the Zipf variant skewed the way DESIGN.md says. CPU only."""
import numpy as np

from bpmf_b200 import synthetic


def _check_orientations(r):
    assert r.u_ptr[-1] == r.m_ptr[-1] == r.nnz
    for ptr, idx, n_other in ((r.u_ptr, r.u_idx, r.ncols), (r.m_ptr, r.m_idx, r.nrows)):
        assert idx.min() >= 0 and idx.max() < n_other
        ascending = np.diff(idx.astype(np.int64)) > 0
        starts = ptr[1:-1]                                               # where a new column begins, the index may drop
        ascending[starts[(starts > 0) & (starts < len(idx))] - 1] = True
        assert ascending.all()
    # the two orientations hold the same matrix
    rows = np.repeat(np.arange(r.nrows), np.diff(r.u_ptr))
    a = np.lexsort((rows, r.u_idx))
    assert np.array_equal(r.u_idx[a], np.repeat(np.arange(r.ncols), np.diff(r.m_ptr)))
    assert np.array_equal(rows[a], r.m_idx) and np.array_equal(r.u_val[a], r.m_val)
    assert abs(r.mean_rating - r.u_val.sum() / r.nnz) < 1e-12


def test_uniform_generator_is_deterministic_and_consistent():
    a = synthetic.generate(3000, 2000, 30.0, 7)
    b = synthetic.generate(3000, 2000, 30.0, 7)
    assert a.nnz == b.nnz and np.array_equal(a.u_idx, b.u_idx) and a.u_val.tobytes() == b.u_val.tobytes()
    assert np.array_equal(a.t_rows, b.t_rows) and a.t_vals.tobytes() == b.t_vals.tobytes()
    _check_orientations(a)
    c = synthetic.generate(3000, 2000, 30.0, 8)
    assert not np.array_equal(a.u_idx[:1000], c.u_idx[:1000])


def test_zipf_generator_is_skewed():
    r = synthetic.generate(4000, 4000, 40.0, 11, zipf=1.0)
    _check_orientations(r)
    per_movie = np.diff(r.m_ptr)
    per_user = np.diff(r.u_ptr)
    # popularity ~ 1 / rank: the hottest movie is rated by most users, the median movie by a handful
    assert per_movie.max() > 0.8 * r.nrows and np.median(per_movie) < 0.01 * r.nrows
    assert per_user.max() < 4 * per_user.mean()          # rows stay Poisson-like
    assert r.nnz < 4000 * 40                             # duplicates of hot movies inside a row are dropped


def test_named_workloads_have_the_baseline_shapes():
    w = synthetic.WORKLOADS
    assert w["synthA-1Mx1M-100Mnnz-K32"][:4] == (1_000_000, 1_000_000, 100.0, 32)
    assert w["synthB-200Kx200K-50Mnnz-K128"][:4] == (200_000, 200_000, 250.0, 128)
    assert set(synthetic.ZIPF) <= set(w)
