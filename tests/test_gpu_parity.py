"""GPU parity tests: every stage of the CUDA path, called through the C ABI (libbpmf_b200.so), against the CPU
oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): 1e-10 on latent vectors / per-iteration column means, 1e-6 on RMSE.
The EXACT kernel variant (reference summation order, no FMA) is held to a far tighter bound, which pins the
RNG stream logic: only log() (<= 1 ulp vs glibc) may differ.
"""
import numpy as np
import pytest

import util
from util import MOVIES, USERS

pytestmark = pytest.mark.gpu

TOL_ITEMS = 1e-10
TOL_RMSE = 1e-6
TOL_EXACT = 2e-12


@pytest.fixture(scope="module")
def gpu():
    import bpmf_b200
    bpmf_b200.load_library()
    return bpmf_b200


def test_library_is_the_cuda_one(gpu):
    # the product path must be the in-tree .so; no fallback module exists
    import os
    assert os.path.exists(gpu.SO_PATH)
    ctx = gpu.Context(32)
    assert ctx.launch_count() == 0
    ctx.close()


@pytest.mark.parametrize("c", [0, 10, 32, 12345, 0xFFFFFFF0])
def test_device_randn_matches_libstdcxx_stream(gpu, oracle_lib, c):
    ctx = gpu.Context(32)
    for n in (1, 6, 32, 128, 1000):
        got = ctx.debug_randn(c, n)
        ref = oracle_lib.randn(c, n)
        np.testing.assert_allclose(got, ref, rtol=4e-16 * 8, atol=0)
    ctx.close()


@pytest.mark.parametrize("K,N,it", [(10, 4, 0), (10, 2, 3), (16, 1000, 1), (32, 943, 7), (32, 1000000, 19), (64, 5000, 2),
                                    (128, 200000, 5)])
def test_hyper_draw(gpu, oracle_lib, K, N, it):
    cov = util.random_spd(K, 100 + K + it, scale=0.3)
    mu_o, LU_o, LF_o = oracle_lib.hyper(K, N, it, cov.T.copy())
    ctx = gpu.Context(K)
    # a side with N items: give it an empty pattern
    ctx.load_side(MOVIES, N, 8, np.zeros(N + 1, np.int64), np.zeros(0, np.int32), np.zeros(0), 0.0)
    ctx.sample_hyper(MOVIES, it, None, cov.T.copy().reshape(-1))
    mu, LU, LF = ctx.get_hyper(MOVIES)
    scale = np.abs(LF_o).max()
    assert np.abs(LU - LU_o).max() <= 1e-11 * np.abs(LU_o).max()
    assert np.abs(LF - LF_o).max() <= 1e-11 * scale
    assert np.abs(mu - mu_o).max() <= 1e-11 * max(1.0, np.abs(mu_o).max())
    ctx.close()


@pytest.mark.parametrize("walk", ["speculative", "sequential"])
def test_hyper_draw_gamma_walk_paths(gpu, oracle_lib, walk, monkeypatch):
    """The K gamma draws of the Bartlett factor (c++/mvnormal.cpp:64-73): positions first and the draws in parallel when every
    draw takes its first normal and uniform, the plain sequential walk otherwise (csrc/exact_kernels.cu, step 4). Tiny N
    makes the shape parameters small (Marsaglia-Tsang then rejects in a few per cent of the draws), so across these cases both
    outcomes of the speculation occur; BPMF_HYPER_SEQUENTIAL forces the fall-back for all of them."""
    if walk == "sequential":
        monkeypatch.setenv("BPMF_HYPER_SEQUENTIAL", "1")
    else:
        monkeypatch.delenv("BPMF_HYPER_SEQUENTIAL", raising=False)
    K = 32
    for N in (1, 2, 3, 7):
        ctx = gpu.Context(K)
        ctx.load_side(MOVIES, N, 8, np.zeros(N + 1, np.int64), np.zeros(0, np.int32), np.zeros(0), 0.0)
        for it in range(6):
            cov = util.random_spd(K, 500 + 10 * N + it, scale=0.3)
            mu_o, LU_o, LF_o = oracle_lib.hyper(K, N, it, cov.T.copy())
            ctx.sample_hyper(MOVIES, it, None, cov.T.copy().reshape(-1))
            mu, LU, LF = ctx.get_hyper(MOVIES)
            assert np.abs(LU - LU_o).max() <= 1e-11 * np.abs(LU_o).max(), (N, it)
            assert np.abs(LF - LF_o).max() <= 1e-11 * np.abs(LF_o).max(), (N, it)
            assert np.abs(mu - mu_o).max() <= 1e-11 * max(1.0, np.abs(mu_o).max()), (N, it)
        ctx.close()


def test_hyper_draw_zero_cov_first_sweep(gpu, oracle_lib):
    # sweep 0: cov == 0 -> X = I (quirk Q6)
    K, N = 32, 1682
    mu_o, LU_o, LF_o = oracle_lib.hyper(K, N, 0, np.zeros((K, K)))
    ctx = gpu.Context(K)
    ctx.load_side(MOVIES, N, 8, np.zeros(N + 1, np.int64), np.zeros(0, np.int32), np.zeros(0), 0.0)
    ctx.sample_hyper(MOVIES, 0)
    mu, LU, LF = ctx.get_hyper(MOVIES)
    np.testing.assert_allclose(LF, LF_o, rtol=0, atol=1e-11 * np.abs(LF_o).max())
    np.testing.assert_allclose(mu, mu_o, rtol=0, atol=1e-12)
    ctx.close()


def _prime(orc, ctx, K, seed):
    """Put identical non-trivial latents and hyper-parameters into oracle and GPU."""
    rng = np.random.Generator(np.random.PCG64(seed))
    for side in (MOVIES, USERS):
        x = rng.normal(0, 0.7, size=(orc.num(side), K))
        orc.set_items(side, x)
        ctx.set_items(side, x)
        LF = util.random_spd(K, seed + 10 + side, scale=2.0) + np.eye(K)
        mu = rng.normal(0, 0.3, size=K)
        orc.set_hyper(side, mu, LF)
        ctx.set_hyper(side, mu, LF)


CASES = [
    # K, rows, cols, nnz, kwargs
    (10, 40, 30, 300, {}),
    (32, 300, 200, 6000, {}),
    (32, 500, 64, 4000, dict(skew=1.2, empty_rows=40, heavy_col=450)),   # empty rows, a hot column > 32*k
    (16, 120, 90, 1500, {}),
    (64, 90, 70, 2500, {}),
    (48, 150, 40, 2000, dict(skew=1.0, empty_rows=10, heavy_col=120)),
    (128, 60, 50, 1800, dict(heavy_col=55)),
]


VARIANTS = {"exact": 1, "stream": 3, "block": 4}
BLOCK_K = (16, 48, 64, 80, 96, 112, 128)


@pytest.mark.parametrize("K,nr,nc,nnz,kw", CASES)
@pytest.mark.parametrize("variant", ["exact", "stream", "block"])
def test_item_update_one_sweep(gpu, K, nr, nc, nnz, kw, variant):
    if variant == "stream" and K != 32:
        pytest.skip("the warp-per-item tensor-core kernels are K == 32")
    if variant == "block" and K not in BLOCK_K:
        pytest.skip("the CTA-per-item tensor-core kernel is K = 16 m, K != 32")
    v = VARIANTS[variant]
    tol = TOL_EXACT if variant == "exact" else TOL_ITEMS
    train, test = util.synth_ratings(nr, nc, nnz, 7 + K, **kw)
    orc = util.make_oracle(K, train, test)
    ctx = util.make_gpu_from_oracle(orc, K)
    _prime(orc, ctx, K, 99)
    for side, it in ((MOVIES, 3), (USERS, 3), (MOVIES, 4)):
        orc.set_iter(side, it)
        orc.sample_range(side, 0, orc.num(side))
        ctx.sample_items(side, it, 2.0, v)
        got, ref = ctx.get_items(side), orc.items(side)
        assert np.abs(got - ref).max() <= tol * max(1.0, np.abs(ref).max()), (side, it)
        # keep both in lock-step for the next sweep
        ctx.set_items(side, ref)
    ctx.close()


def test_stats_and_cov(gpu):
    K = 32
    train, test = util.synth_ratings(700, 333, 5000, 3)
    orc = util.make_oracle(K, train, test)
    ctx = util.make_gpu_from_oracle(orc, K)
    orc.sweep(MOVIES)
    ctx.set_items(MOVIES, orc.items(MOVIES))
    ctx.reduce_stats(MOVIES)
    s, p, c, n = ctx.get_stats(MOVIES)
    so, po, co, no = orc.stats(MOVIES)
    np.testing.assert_allclose(s, so, rtol=0, atol=1e-11 * max(1, np.abs(so).max()))
    np.testing.assert_allclose(p, po, rtol=0, atol=1e-11 * np.abs(po).max())
    np.testing.assert_allclose(c, co, rtol=0, atol=1e-11 * np.abs(co).max())
    assert abs(n - no) <= 1e-11 * no
    ctx.close()


def _run_both(gpu, K, train, test, iters, burnin, variant, alpha=2.0):
    orc = util.make_oracle(K, train, test, alpha=alpha, burnin=burnin)
    ctx = util.make_gpu_from_oracle(orc, K)
    worst = 0.0
    for i in range(iters):
        orc.iterate()
        ctx.sample(MOVIES, alpha, variant)
        ctx.sample(USERS, alpha, variant)
        rm = ctx.predict(MOVIES, burnin)
        ru = ctx.predict(USERS, burnin)
        for side in (MOVIES, USERS):
            got, ref = ctx.get_items(side), orc.items(side)
            d = np.abs(got - ref).max()
            worst = max(worst, d)
            assert d <= TOL_ITEMS * max(1.0, np.abs(ref).max()), (i, side, d)
            # per-iteration posterior column means (the north-star gate)
            assert np.abs(got.mean(0) - ref.mean(0)).max() <= TOL_ITEMS
        ro = orc.rmse(MOVIES)
        assert abs(rm[0] - ro[0]) <= TOL_RMSE and abs(rm[1] - ro[1]) <= TOL_RMSE and rm[2] == ro[2], (i, rm, ro)
        rou = orc.rmse(USERS)
        assert abs(ru[0] - rou[0]) <= TOL_RMSE and abs(ru[1] - rou[1]) <= TOL_RMSE
        assert abs(np.sqrt(ctx.get_stats(USERS)[3]) - np.sqrt(orc.stats(USERS)[3])) <= 1e-9
    # final predict (quirk Q4: bpmf.cpp:225|242 re-applies the last sample)
    orc.finish()
    rm = ctx.predict(MOVIES, burnin)
    assert abs(rm[1] - orc.rmse(MOVIES)[1]) <= TOL_RMSE
    pa, pm = ctx.get_predictions(MOVIES)
    pao, pmo = orc.pred(MOVIES)
    np.testing.assert_allclose(pa, pao, rtol=0, atol=1e-9)
    np.testing.assert_allclose(pm, pmo, rtol=0, atol=1e-8)
    ctx.close()
    return worst


def test_full_run_tiny_k10(gpu):
    # BASELINE.json configs[0]: data/tiny, K=10; run_test.sh uses -i 9 -b 0
    _run_both(gpu, 10, util.TINY_TRAIN, util.TINY_TEST, 9, 0, gpu.KERNEL_AUTO)
    _run_both(gpu, 10, util.TINY_TRAIN, util.TINY_TEST, 20, 5, gpu.KERNEL_EXACT)


@pytest.mark.parametrize("variant", ["exact", "stream"])
def test_full_run_movielens_shaped_k32(gpu, variant):
    # ML-100K-shaped synthetic (943 x 1682, ~80k train / 20k test), K=32, 20 iterations, burn-in 5
    v = VARIANTS[variant]
    train, test = util.synth_ratings(943, 1682, 110000, 2026, rank=10, skew=0.7, test_frac=0.2)
    worst = _run_both(gpu, 32, train, test, 20, 5, v)
    print("worst latent deviation over 20 iterations (%s): %.3e" % (variant, worst))


@pytest.mark.parametrize("K", [16, 64, 128])
def test_full_run_block_kernel(gpu, K):
    # the CTA-per-item kernel (KERNEL_AUTO picks it for these K): whole chains incl. hyper draw, stats, predict
    train, test = util.synth_ratings(400, 250, 12000, 31 + K, rank=6, skew=0.6, empty_rows=15, heavy_col=200)
    worst = _run_both(gpu, K, train, test, 8, 3, gpu.KERNEL_AUTO)
    print("worst latent deviation over 8 iterations (K=%d): %.3e" % (K, worst))


def test_full_run_chembl_shaped_k32(gpu):
    # skewed like ChEMBL: many near-empty rows, empty rows, one very hot column
    train, test = util.synth_ratings(6000, 300, 14000, 77, rank=6, skew=1.1, empty_rows=800, heavy_col=3000)
    _run_both(gpu, 32, train, test, 8, 3, gpu.KERNEL_AUTO)


def test_range_and_peer_push(gpu):
    # two "ranks" on one GPU: each samples half of the items and pushes into the other's replica
    K = 32
    train, test = util.synth_ratings(400, 300, 9000, 11)
    orc = util.make_oracle(K, train, test)
    a = util.make_gpu_from_oracle(orc, K)
    b = util.make_gpu_from_oracle(orc, K)
    _prime(orc, a, K, 5)
    _prime(orc, b, K, 5)
    n = orc.num(MOVIES)
    half = n // 2 + 3
    pa, pb = a.items_device_ptr(MOVIES), b.items_device_ptr(MOVIES)
    for ctx, (lo, hi) in ((a, (0, half)), (b, (half, n))):
        ctx.set_range(MOVIES, lo, hi)
        ctx.set_peers(MOVIES, [pa, pb])
    orc.set_iter(MOVIES, 2)
    orc.sample_range(MOVIES, 0, n)
    a.sample_items(MOVIES, 2, 2.0, gpu.KERNEL_STREAM)
    b.sample_items(MOVIES, 2, 2.0, gpu.KERNEL_EXACT)
    a.sync(); b.sync()
    ref = orc.items(MOVIES)
    for ctx in (a, b):
        assert np.abs(ctx.get_items(MOVIES) - ref).max() <= TOL_ITEMS * np.abs(ref).max()
    a.close(); b.close()


@pytest.mark.parametrize("K", [32, 16])
def test_sliced_statistics_match_the_full_reduction(gpu, K):
    """Multi-GPU sweep statistics (replaces the three MPI_Allreduce of c++/mpi_common.h:44-50): three "ranks" on one GPU each
    reduce only the statistics blocks of their own block-aligned item range and store the partials into every rank's
    buffer; after that every rank's sum / prod / cov / norm are bit-identical to what one context reduces by itself."""
    train, test = util.synth_ratings(5000, 4100, 60000, 12)
    orc = util.make_oracle(K, train, test)
    ctxs = [util.make_gpu_from_oracle(orc, K) for _ in range(4)]
    single, ranks = ctxs[0], ctxs[1:]
    rng = np.random.default_rng(1)
    for side in (MOVIES, USERS):
        n = orc.num(side)
        x = rng.normal(0, 0.7, size=(n, K))
        for c in ctxs:
            c.set_items(side, x)
        single.reduce_stats(side)
        ref = single.get_stats(side)
        bi = single.stats_block_items(side)
        assert bi >= 1 and bi * 296 >= n
        cuts = [0, 5 * bi, min(n, 5 * bi + 40 * bi), n]           # ragged, block-aligned; the last range takes the tail blocks
        ptrs = [c.stats_device_ptr(side) for c in ranks]
        for r, c in enumerate(ranks):
            c.set_range(side, cuts[r], cuts[r + 1])
            c.set_stats_peers(side, ptrs)
        for c in ranks:
            c.reduce_stats_partial(side)
        for c in ranks:
            c.sync()
        for c in ranks:
            c.reduce_stats_final(side)
            got = c.get_stats(side)
            for a, b in zip(got[:3], ref[:3]):
                assert a.tobytes() == b.tobytes()
            assert got[3] == ref[3]
        # a range that does not sit on block boundaries is refused while peers are set
        ranks[0].set_range(side, 1, cuts[1])
        with pytest.raises(gpu.BpmfGpuError):
            ranks[0].reduce_stats_partial(side)
        ranks[0].set_stats_peers(side, [])
        ranks[0].reduce_stats(side)                                # without peers: the full replica, any range
        assert ranks[0].get_stats(side)[2].tobytes() == ref[2].tobytes()
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("K", [32, 16])
def test_slice_loading_sharded_aggregates_and_device_finalize(gpu, K):
    """One of G GPUs holds only the ratings and the posterior aggregates of ITS items (c++/bpmf.h:161-176):
    bpmf_gpu_load_side_slice must sample its range exactly like a context that loaded everything, the aggregates of -o
    (c++/sample.cpp:364-368) exist for the range only, and bpmf_gpu_finalize_aggregates (c++/bpmf.cpp:281-295: posterior mean
    and the inverse of the sample covariance, batched on the device) agrees with numpy."""
    train, test = util.synth_ratings(300, 260, 9000, 5)
    orc = util.make_oracle(K, train, test)
    full = util.make_gpu_from_oracle(orc, K)
    part = gpu.Context(K)
    rng_ = {}
    for side in (MOVIES, USERS):
        n = orc.num(side)
        lo, hi = n // 3, (2 * n) // 3 + 1
        rng_[side] = (lo, hi)
        colptr, rowidx, val = orc.csc(side, 0)
        part.load_side_slice(side, n, orc.num(1 - side), lo, hi, colptr, rowidx, val, orc.mean_rating(side))
        nnz, _, dptr, _, _ = part.get_side(side)
        assert nnz == colptr[hi] - colptr[lo]                    # only the slice's ratings are resident
    _prime(orc, full, K, 21)
    _prime(orc, part, K, 21)
    nsamples = K + 8
    acc = {side: None for side in (MOVIES, USERS)}
    for side in (MOVIES, USERS):
        part.enable_aggregation(side, 0)
    for it in range(nsamples):
        for side in (MOVIES, USERS):
            lo, hi = rng_[side]
            full.set_range(side, lo, hi)
            full.sample_items(side, it, 2.0, gpu.KERNEL_AUTO)
            part.sample_items(side, it, 2.0, gpu.KERNEL_AUTO)
            part.aggregate(side)
            a, b = part.get_items(side), full.get_items(side)
            assert a[lo:hi].tobytes() == b[lo:hi].tobytes(), (it, side)
            x = a[lo:hi]
            mu, lam = x.copy(), np.einsum("ia,ib->iba", x, x).reshape(len(x), K * K)
            acc[side] = (mu, lam) if acc[side] is None else (acc[side][0] + mu, acc[side][1] + lam)
            # both contexts continue from the same state: the other items keep their primed values
            part.set_items(side, b)
    with pytest.raises(gpu.BpmfGpuError):
        part.set_range(MOVIES, 0, rng_[MOVIES][1])              # outside the resident slice
    for side in (MOVIES, USERS):
        lo, hi = rng_[side]
        amu, alam = part.get_aggregates(side)
        assert not amu[:lo].any() and not amu[hi:].any()        # nothing outside the range
        np.testing.assert_allclose(amu[lo:hi], acc[side][0], rtol=0, atol=1e-11 * np.abs(acc[side][0]).max())
        np.testing.assert_allclose(alam[lo:hi], acc[side][1], rtol=0, atol=1e-11 * np.abs(acc[side][1]).max())
        part.finalize_aggregates(side, nsamples)
        fmu, flam = part.get_aggregates(side)
        np.testing.assert_allclose(fmu[lo:hi], amu[lo:hi] / nsamples, rtol=1e-14, atol=0)
        for i in range(lo, hi, max(1, (hi - lo) // 7)):
            s_, p_ = amu[i], alam[i].reshape(K, K)
            cov = (p_ - np.outer(s_, s_) / nsamples) / (nsamples - 1)
            prec = flam[i].reshape(K, K)
            np.testing.assert_allclose(prec @ cov, np.eye(K), rtol=0, atol=1e-7 * np.linalg.cond(cov))
            ref = np.linalg.inv(cov)
            assert np.abs(prec - ref).max() <= 1e-9 * np.linalg.cond(cov) * np.abs(ref).max()
    full.close(); part.close()


@pytest.mark.parametrize("reductions", ["auxiliary stream", "main stream"])
def test_two_ranks_on_one_gpu_whole_protocol(gpu, reductions, monkeypatch):
    """The multi-GPU sweep as ONE C call per rank (bpmf_gpu_sample with peers set): two "ranks" on one GPU, each on its own
    stream, each holding only the ratings of its own items; item kernels store into both replicas, every rank reduces its
    own statistics blocks into both buffers, the device-side barriers (peer_barrier_kernel) order it all. Whole chains,
    predict included, must be bit-identical to one context that does everything — with the reductions of a sweep on the
    auxiliary stream under the other side's sweep (the default with peers: statistics barrier there, latents barrier on the
    main stream) and with everything on the main stream (BPMF_STATS_MAIN, read when a context is created)."""
    torch = pytest.importorskip("torch")
    if reductions == "main stream":
        monkeypatch.setenv("BPMF_STATS_MAIN", "1")
    else:
        monkeypatch.delenv("BPMF_STATS_MAIN", raising=False)
    K = 32
    train, test = util.synth_ratings(5000, 4000, 70000, 9)
    orc = util.make_oracle(K, train, test, burnin=1)
    single = util.make_gpu_from_oracle(orc, K)
    ranks, streams = [gpu.Context(K), gpu.Context(K)], [torch.cuda.Stream(), torch.cuda.Stream()]
    for side in (MOVIES, USERS):
        n = orc.num(side)
        bi = gpu.capi.stats_block_items_for(K, n)
        cut = (n // 2 // bi) * bi
        colptr, rowidx, val = orc.csc(side, 0)
        for r, c in enumerate(ranks):
            lo, hi = (0, cut) if r == 0 else (cut, n)
            c.load_side_slice(side, n, orc.num(1 - side), lo, hi, colptr, rowidx, val, orc.mean_rating(side))
    for r, c in enumerate(ranks):
        c.set_stream(streams[r].cuda_stream)
        for side in (MOVIES, USERS):
            c.load_test(side, *orc.csc(side, 1))
    for side in (MOVIES, USERS):
        c0, c1 = ranks
        c0.set_peers(side, [c0.items_device_ptr(side), c1.items_device_ptr(side)])
        c1.set_peers(side, [c0.items_device_ptr(side), c1.items_device_ptr(side)])
        for c in ranks:
            c.set_stats_peers(side, [c0.stats_device_ptr(side), c1.stats_device_ptr(side)])
    for it in range(4):
        for side in (MOVIES, USERS):
            single.sample(side)
            for c in ranks:
                c.sample(side)                    # collective: both ranks enqueue, nobody synchronises in between
        ref = single.predict(MOVIES, 1)
        for c in ranks:
            assert c.predict(MOVIES, 1) == ref
            for side in (MOVIES, USERS):
                assert c.get_items(side).tobytes() == single.get_items(side).tobytes(), (it, side)
                a, b = c.get_stats(side), single.get_stats(side)
                assert a[2].tobytes() == b[2].tobytes() and a[3] == b[3]
        for c in ranks:
            c.peer_barrier(MOVIES)                # readers done before the next sweep's remote stores (what the sampler does)
    for c in ranks + [single]:
        c.close()


@pytest.mark.parametrize("cfg,what", [(3216, "16 warps"), (6220, "TMA 1-D bulk copies (UBLKCP)"), (14220, "TMA tile::gather4 (UTMALDG.2D.GATHER4)"),
                                      (14316, "gather4, 3 stages x 16 warps"), (15220, "rank-one DMMA column steps"),
                                      (900003220, "fixed bulk / tail claims instead of guided"),
                                      (17220, "pivots and reciprocals through shared memory"), (19220, "rank-one DMMA for block columns 2 and 3"),
                                      (18220, "factor's panel blocks by the separate scatter"), (21220, "stage indices by lane = row"),
                                      (24220, "neither of the two"), (27220, "rating weights computed per staged rating"),
                                      (28220, "right-hand side's quad sums by shuffles")])
def test_stream_kernel_variants_are_bit_identical(gpu, cfg, what):
    """The measured alternatives of the K = 32 kernel that stay selectable in the product build: other
    gather engines, claim schedules and the tensor-core column step compute the same bits as the default configuration —
    ragged items, empty items, a claim group's worth of items and more, two sweeps."""
    K = 32
    train, test = util.synth_ratings(9000, 2500, 160000, 19, skew=0.8, empty_rows=300)
    orc = util.make_oracle(K, train, test)
    ctx = util.make_gpu_from_oracle(orc, K)
    _prime(orc, ctx, K, 3)
    state = [ctx.get_items(MOVIES), ctx.get_items(USERS)]
    outs = []
    for c in (0, cfg):
        ctx.set_tuning(c)
        ctx.set_items(MOVIES, state[MOVIES]); ctx.set_items(USERS, state[USERS])
        for side, it in ((USERS, 2), (MOVIES, 2)):
            ctx.sample_items(side, it, 2.0, gpu.KERNEL_STREAM)
        outs.append((ctx.get_items(MOVIES), ctx.get_items(USERS)))
    assert outs[0][0].tobytes() == outs[1][0].tobytes() and outs[0][1].tobytes() == outs[1][1].tobytes(), what
    orc.set_iter(USERS, 2); orc.sample_range(USERS, 0, orc.num(USERS))
    assert np.abs(outs[1][1] - orc.items(USERS)).max() <= TOL_ITEMS * max(1.0, np.abs(orc.items(USERS)).max())
    ctx.close()


def test_cholesky_failure_is_reported(gpu):
    K = 32
    train, test = util.synth_ratings(50, 40, 600, 1)
    orc = util.make_oracle(K, train, test)
    for v in (gpu.KERNEL_EXACT, gpu.KERNEL_STREAM):
        ctx = util.make_gpu_from_oracle(orc, K)
        ctx.set_hyper(MOVIES, np.zeros(K), -np.eye(K))   # not positive definite -> "Cholesky failed" (sample.cpp:308)
        ctx.sample_items(MOVIES, 0, 2.0, v)
        with pytest.raises(gpu.BpmfGpuError) as ei:
            ctx.sync()
        assert ei.value.code == 3 and "Cholesky failed" in str(ei.value)
        ctx.close()


def test_bind_external_items_storage(gpu):
    torch = pytest.importorskip("torch")
    K = 32
    train, test = util.synth_ratings(100, 80, 1500, 4)
    orc = util.make_oracle(K, train, test)
    ctx = util.make_gpu_from_oracle(orc, K)
    _prime(orc, ctx, K, 8)
    ext = torch.zeros(orc.num(MOVIES), K, dtype=torch.float64, device="cuda")
    ctx.bind_items(MOVIES, ext.data_ptr())
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    orc.set_iter(MOVIES, 1)
    orc.sample_range(MOVIES, 0, orc.num(MOVIES))
    ctx.sample_items(MOVIES, 1, 2.0, gpu.KERNEL_AUTO)
    ctx.sync()
    assert np.abs(ext.cpu().numpy() - orc.items(MOVIES)).max() <= TOL_ITEMS * 10
    ctx.close()


@pytest.mark.parametrize("mode", ["allgather", "push"])
def test_two_gpus_match_one(gpu, mode):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "multi_gpu_worker.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "identical to single GPU" in r.stdout


@pytest.mark.parametrize("family", ["chain", "refexe_chain"])
@pytest.mark.parametrize("name", ["tiny_k10", "synth_k32"])
def test_chain_matches_committed_golden(gpu, name, family):
    """Whole chains through bpmf_gpu_sample / bpmf_gpu_predict against the committed fixtures (no oracle binary involved):
    per-iteration column means to 1e-10, RMSE to 1e-6 — BASELINE.json's parity gate. tests/golden/chain_*.json come from
    the oracle; tests/golden/refexe_chain_*.json from the REFERENCE EXECUTABLE (its sources compiled with the stand-in
    Eigen / Random123 headers, tests/golden/make_ref_exe_golden.py): latents in full precision from its -v dumps, RMSEs as
    its log prints them (4 decimals; "Final Avg RMSE" with 6 significant digits)."""
    import importlib.util
    import json
    import os
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gold = json.load(open(os.path.join(gdir, "%s_%s.json" % (family, name))))
    spec_mod = importlib.util.spec_from_file_location("make_chain_golden", os.path.join(gdir, "make_chain_golden.py"))
    mod = importlib.util.module_from_spec(spec_mod)
    spec_mod.loader.exec_module(mod)
    spec = gold["spec"]
    (shape, rows, cols, vals), (tshape, trows, tcols, tvals) = mod.problem(spec)
    import scipy.sparse as sp
    K, burnin = spec["K"], spec["burnin"]
    R = sp.coo_matrix((vals, (rows, cols)), shape=shape)
    T = sp.coo_matrix((tvals, (trows, tcols)), shape=tshape)
    mean = float(np.asarray(vals).sum() / len(vals))
    ctx = gpu.Context(K)
    for side, M in ((MOVIES, R.tocsc()), (USERS, R.T.tocsc())):
        M.sort_indices()
        ctx.load_side(side, M.shape[1], M.shape[0], M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data, mean)
    for side, M in ((MOVIES, T.tocsc()), (USERS, T.T.tocsc())):
        M.sort_indices()
        ctx.load_test(side, M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data)
    for g in gold["iterations"]:
        ctx.sample(MOVIES, 2.0, gpu.KERNEL_AUTO)
        ctx.sample(USERS, 2.0, gpu.KERNEL_AUTO)
        rm = ctx.predict(MOVIES, burnin)
        ctx.predict(USERS, burnin)
        V, U = ctx.get_items(MOVIES), ctx.get_items(USERS)
        np.testing.assert_allclose(V.mean(0), g["V_mean"], rtol=0, atol=TOL_ITEMS)
        np.testing.assert_allclose(U.mean(0), g["U_mean"], rtol=0, atol=TOL_ITEMS)
        np.testing.assert_allclose(V[0], g["V_first"], rtol=0, atol=TOL_ITEMS * max(1.0, np.abs(g["V_first"]).max()))
        np.testing.assert_allclose(U[-1], g["U_last"], rtol=0, atol=TOL_ITEMS * max(1.0, np.abs(g["U_last"]).max()))
        if "rmse" in g:
            assert abs(rm[0] - g["rmse"]) <= TOL_RMSE and abs(rm[1] - g["rmse_avg"]) <= TOL_RMSE
        else:                                 # as printed by the reference: "%3.4f"
            assert abs(rm[0] - float(g["log"]["rmse"])) <= 5e-5 + TOL_RMSE and abs(rm[1] - float(g["log"]["rmse_avg"])) <= 5e-5 + TOL_RMSE
    final = ctx.predict(MOVIES, burnin)      # the extra predict of c++/bpmf.cpp:225|242
    if "final_avg_rmse" in gold:
        assert abs(final[1] - gold["final_avg_rmse"]) <= TOL_RMSE
    else:                                     # operator<< of a double: 6 significant digits
        printed = float(gold["final_avg_rmse_printed"])
        assert abs(final[1] - printed) <= 1e-5 * max(1.0, abs(printed))
    ctx.close()


def test_sample_host_matches_device_resident_sample(gpu):
    """bpmf_gpu_sample_host (host-resident latent matrices, the e2e path of bench.py) against bpmf_gpu_sample on a second
    context: bit-identical, for a side small enough for the single-copy path and one large enough for the path that
    samples in parts and downloads each part while the next is sampled."""
    import scipy.sparse as sp
    K = 32
    train, _ = util.synth_ratings(70000, 500, 400000, 21, skew=0.2)
    (shape, rows, cols, vals) = train
    R = sp.coo_matrix((vals, (rows, cols)), shape=shape)
    mean = float(vals.mean())
    ctxs = [gpu.Context(K), gpu.Context(K)]
    for ctx in ctxs:
        for side, M in ((MOVIES, R.tocsc()), (USERS, R.T.tocsc())):
            M.sort_indices()
            ctx.load_side(side, M.shape[1], M.shape[0], M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data, mean)
    a, b = ctxs
    host = [np.zeros((500, K)), np.zeros((70000, K))]
    for it in range(3):
        for side in (MOVIES, USERS):
            a.sample(side, 2.0, gpu.KERNEL_AUTO)
            b.sample_host(side, host[1 - side].ctypes.data, host[side].ctypes.data, 2.0, gpu.KERNEL_AUTO)
            ref = a.get_items(side)
            assert host[side].tobytes() == ref.tobytes(), (it, side, np.abs(host[side] - ref).max())
            assert b.get_items(side).tobytes() == ref.tobytes()
    sa, sb = a.get_stats(USERS), b.get_stats(USERS)
    assert sa[2].tobytes() == sb[2].tobytes() and sa[3] == sb[3]
    for ctx in ctxs:
        ctx.close()


@pytest.mark.parametrize("K", [16, 32, 64])
def test_propagated_posterior_priors(gpu, K):
    """-m / -l (c++/sample.cpp:152-174,272-283): per-item prior precisions replace hp.LambdaF, the rhs still uses the
    GLOBAL hp.mu (quirk Q5). Against the oracle: the kernel KERNEL_AUTO picks (K = 32: the PROP instantiation of the
    stream kernel; K = 16 m: the CTA-per-item kernel, which reads the per-item precision too) and the any-K kernel asked
    for explicitly."""
    train, test = util.synth_ratings(120, 90, 2500, 5 + K)
    orc = util.make_oracle(K, train, test)
    ctx = util.make_gpu_from_oracle(orc, K)
    _prime(orc, ctx, K, 17)
    rng = np.random.default_rng(K)
    for side in (MOVIES, USERS):
        n = orc.num(side)
        lam = np.stack([util.random_spd(K, 1000 * side + i, scale=2.0).T.reshape(-1) for i in range(n)])
        mu = rng.normal(size=(n, K))
        orc.set_prop(side, mu, lam)
        ctx.set_prop_posterior(side, mu, lam)
    for side, it in ((MOVIES, 2), (USERS, 2), (MOVIES, 3)):
        orc.set_iter(side, it)
        orc.sample_range(side, 0, orc.num(side))
        ref = orc.items(side)
        for variant in ([gpu.KERNEL_AUTO, gpu.KERNEL_EXACT, gpu.KERNEL_STREAM] if K == 32 else [gpu.KERNEL_AUTO, gpu.KERNEL_BLOCK, gpu.KERNEL_EXACT]):
            before = ctx.get_items(side)
            ctx.sample_items(side, it, 2.0, variant)
            got = ctx.get_items(side)
            tol = (TOL_ITEMS if variant != gpu.KERNEL_EXACT else TOL_EXACT * 50) * max(1.0, np.abs(ref).max())
            assert np.abs(got - ref).max() <= tol, (side, it, variant)
            ctx.set_items(side, before)
        ctx.set_items(side, ref)
    with pytest.raises(gpu.BpmfGpuError):                 # variant 2 (round 1's first tensor-core kernel) no longer exists
        ctx.sample_items(MOVIES, 4, 2.0, 2)
    ctx.set_prop_posterior(MOVIES, None, None)          # priors removed: back to the shared hyper-parameters
    ctx.sample_items(MOVIES, 4, 2.0, gpu.KERNEL_STREAM if K == 32 else gpu.KERNEL_AUTO)
    ctx.sync()
    ctx.close()


def test_heavy_items_chunked_path(gpu):
    """Skew handling: items far heavier than the rest are cut into chunks (partial Grams by separate warps, added in a
    fixed order) and the stream kernel passes over them. Two hot movies of ~5000 and ~2600 ratings
    (threshold lowered to 2100 so that both take the chunked path: 3 and 2 chunks), against the oracle and against the
    plain path (threshold out of reach)."""
    K = 32
    rng = np.random.default_rng(8)
    nr, nc = 6000, 400
    rows = rng.integers(0, nr, size=30000); cols = rng.integers(0, nc, size=30000)
    hot_a = rng.choice(nr, size=5000, replace=False); hot_b = rng.choice(nr, size=2600, replace=False)
    rows = np.concatenate([rows, hot_a, hot_b]); cols = np.concatenate([cols, np.full(5000, 7), np.full(2600, 399)])
    key = rows.astype(np.int64) * nc + cols
    _, first = np.unique(key, return_index=True)
    rows, cols = rows[first].astype(np.int32), cols[first].astype(np.int32)
    vals = rng.normal(3.5, 1.0, size=len(rows))
    train = ((nr, nc), rows, cols, vals)
    test = ((nr, nc), rows[:50].copy(), cols[:50].copy(), vals[:50].copy())
    orc = util.make_oracle(K, train, test)
    chunked = util.make_gpu_from_oracle(orc, K, heavy_threshold=2100)
    plain = util.make_gpu_from_oracle(orc, K, heavy_threshold=1 << 40)
    for ctx in (chunked, plain):
        _prime(orc, ctx, K, 5)
    for it in (1, 2):
        orc.set_iter(MOVIES, it)
        orc.sample_range(MOVIES, 0, orc.num(MOVIES))
        ref = orc.items(MOVIES)
        outs = []
        for ctx in (chunked, plain):
            ctx.sample_items(MOVIES, it, 2.0, gpu.KERNEL_STREAM)
            got = ctx.get_items(MOVIES)
            assert np.abs(got - ref).max() <= TOL_ITEMS * max(1.0, np.abs(ref).max()), it
            outs.append(got)
            ctx.set_items(MOVIES, ref)
        # every item that is not heavy is computed by the same code in the same order: bit-identical
        mask = np.ones(nc, bool); mask[[7, 399]] = False
        assert outs[0][mask].tobytes() == outs[1][mask].tobytes()
    assert chunked.launch_count() > plain.launch_count()
    # a sub-range that ends between the two heavy items (multi-GPU ranges)
    chunked.set_range(MOVIES, 3, 200)
    chunked.sample_items(MOVIES, 3, 2.0, gpu.KERNEL_STREAM)
    orc.set_iter(MOVIES, 3)
    orc.sample_range(MOVIES, 3, 200)
    got, ref = chunked.get_items(MOVIES), orc.items(MOVIES)
    assert np.abs(got - ref).max() <= TOL_ITEMS * max(1.0, np.abs(ref).max())
    for ctx in (chunked, plain):
        ctx.close()


@pytest.mark.parametrize("K", [64, 16])
def test_heavy_items_in_the_cta_per_item_kernel(gpu, K):
    """Skew handling outside K = 32 (the reference's schedule(guided), c++/sample.cpp:352-355, covers every K): an item with
    far more ratings than the rest is cut into chunks of 4096 ratings, one CTA per chunk computes a partial Gram, the item
    kernel adds them in chunk order. A hot movie of ~9000 ratings (3 chunks) and one of ~4500 (2 chunks) against the oracle
    and against the unchunked path: every other item bit-identical."""
    rng = np.random.default_rng(80 + K)
    nr, nc = 12000, 200
    rows = rng.integers(0, nr, size=9000); cols = rng.integers(0, nc, size=9000)
    hot_a = rng.choice(nr, size=9100, replace=False); hot_b = rng.choice(nr, size=4500, replace=False)
    rows = np.concatenate([rows, hot_a, hot_b]); cols = np.concatenate([cols, np.full(9100, 3), np.full(4500, 199)])
    key = rows.astype(np.int64) * nc + cols
    _, first = np.unique(key, return_index=True)
    rows, cols = rows[first].astype(np.int32), cols[first].astype(np.int32)
    vals = rng.normal(3.5, 1.0, size=len(rows))
    train = ((nr, nc), rows, cols, vals)
    test = ((nr, nc), rows[:50].copy(), cols[:50].copy(), vals[:50].copy())
    orc = util.make_oracle(K, train, test)
    chunked = util.make_gpu_from_oracle(orc, K, heavy_threshold=3000)
    plain = util.make_gpu_from_oracle(orc, K, heavy_threshold=1 << 40)
    for ctx in (chunked, plain):
        _prime(orc, ctx, K, 5)
    for it in (1, 2):
        orc.set_iter(MOVIES, it)
        orc.sample_range(MOVIES, 0, nc)
        ref = orc.items(MOVIES)
        outs = []
        for ctx in (chunked, plain):
            before = ctx.launch_count()
            ctx.sample_items(MOVIES, it, 2.0, gpu.KERNEL_BLOCK)
            got = ctx.get_items(MOVIES)
            assert np.abs(got - ref).max() <= TOL_ITEMS * max(1.0, np.abs(ref).max()), it
            outs.append((got, ctx.launch_count() - before))
            ctx.set_items(MOVIES, ref)
        mask = np.ones(nc, bool); mask[[3, 199]] = False
        assert outs[0][0][mask].tobytes() == outs[1][0][mask].tobytes()
        assert outs[0][1] == outs[1][1] + 1                       # + the chunk kernel
    # a range that holds only the second heavy item (multi-GPU ranges)
    chunked.set_range(MOVIES, 120, nc)
    chunked.sample_items(MOVIES, 3, 2.0, gpu.KERNEL_BLOCK)
    orc.set_iter(MOVIES, 3)
    orc.sample_range(MOVIES, 120, nc)
    got, ref = chunked.get_items(MOVIES), orc.items(MOVIES)
    assert np.abs(got - ref).max() <= TOL_ITEMS * max(1.0, np.abs(ref).max())
    for ctx in (chunked, plain):
        ctx.close()


def test_propagated_posterior_with_heavy_items_k32(gpu):
    """-m / -l priors on a side WITH heavy items (K = 32): the stream kernel passes over the heavy items with the per-item
    precisions of the others, the chunked path finishes the heavy ones with theirs (c++/sample.cpp:272-283); against the
    oracle and against the any-K kernel."""
    K = 32
    rng = np.random.default_rng(28)
    nr, nc = 6000, 300
    rows = rng.integers(0, nr, size=20000); cols = rng.integers(0, nc, size=20000)
    hot = rng.choice(nr, size=5000, replace=False)
    rows = np.concatenate([rows, hot]); cols = np.concatenate([cols, np.full(5000, 11)])
    key = rows.astype(np.int64) * nc + cols
    _, first = np.unique(key, return_index=True)
    rows, cols = rows[first].astype(np.int32), cols[first].astype(np.int32)
    vals = rng.normal(3.5, 1.0, size=len(rows))
    train = ((nr, nc), rows, cols, vals)
    test = ((nr, nc), rows[:50].copy(), cols[:50].copy(), vals[:50].copy())
    orc = util.make_oracle(K, train, test)
    ctx = util.make_gpu_from_oracle(orc, K, heavy_threshold=2100)
    _prime(orc, ctx, K, 7)
    lam = np.stack([util.random_spd(K, 3000 + i, scale=2.0).T.reshape(-1) for i in range(nc)])
    mu = rng.normal(size=(nc, K))
    orc.set_prop(MOVIES, mu, lam)
    ctx.set_prop_posterior(MOVIES, mu, lam)
    for it in (2, 3):
        orc.set_iter(MOVIES, it)
        orc.sample_range(MOVIES, 0, nc)
        ref = orc.items(MOVIES)
        before = ctx.launch_count()
        ctx.sample_items(MOVIES, it, 2.0, gpu.KERNEL_AUTO)
        assert ctx.launch_count() - before == 3                  # stream kernel (SKIP + PROP), partial Grams, heavy tails
        got = ctx.get_items(MOVIES)
        assert np.abs(got - ref).max() <= TOL_ITEMS * max(1.0, np.abs(ref).max()), it
        ctx.set_items(MOVIES, ref)
    ctx.close()


def test_many_heavy_items_skipped_in_claim_groups(gpu):
    """Zipf-like skew: ~150 heavy movies among 40 000 (a sweep large enough for 16-item claim groups), placed so that a
    whole claim group is heavy (items 32..47), the first and the last item are heavy, and the rest are scattered. Against
    the oracle, and bit-identical to the plain path for every item that is not heavy."""
    K = 32
    rng = np.random.default_rng(18)
    nr, nc = 3000, 40000
    heavy = np.unique(np.concatenate([[0, nc - 1], np.arange(32, 48), [63, 64, 65], rng.choice(nc, size=130, replace=False)]))
    rows = [rng.integers(0, nr, size=200000)]; cols = [rng.integers(0, nc, size=200000)]
    for h in heavy:
        n = int(rng.integers(520, 1400))
        rows.append(rng.choice(nr, size=n, replace=False)); cols.append(np.full(n, h))
    rows = np.concatenate(rows); cols = np.concatenate(cols)
    key = rows.astype(np.int64) * nc + cols
    _, first = np.unique(key, return_index=True)
    rows, cols = rows[first].astype(np.int32), cols[first].astype(np.int32)
    vals = rng.normal(3.5, 1.0, size=len(rows))
    train = ((nr, nc), rows, cols, vals)
    test = ((nr, nc), rows[:50].copy(), cols[:50].copy(), vals[:50].copy())
    orc = util.make_oracle(K, train, test)
    chunked = util.make_gpu_from_oracle(orc, K, heavy_threshold=500)
    plain = util.make_gpu_from_oracle(orc, K, heavy_threshold=1 << 40)
    for ctx in (chunked, plain):
        _prime(orc, ctx, K, 5)
    orc.set_iter(MOVIES, 1)
    orc.sample_range(MOVIES, 0, nc)
    ref = orc.items(MOVIES)
    outs = []
    for ctx in (chunked, plain):
        ctx.sample_items(MOVIES, 1, 2.0, gpu.KERNEL_STREAM)
        got = ctx.get_items(MOVIES)
        assert np.abs(got - ref).max() <= TOL_ITEMS * max(1.0, np.abs(ref).max())
        outs.append(got)
    mask = np.ones(nc, bool); mask[heavy] = False
    assert outs[0][mask].tobytes() == outs[1][mask].tobytes()
    assert chunked.launch_count() == plain.launch_count() + 2      # + partial Grams + heavy tails
    for ctx in (chunked, plain):
        ctx.close()


def _host_build(n_major, major, minor, val):
    """compressed columns the way the host build does it (bpmf_b200/host/matrix.h from_triplets, after Eigen's
    setFromTriplets): stable order by (major, minor), duplicates summed sequentially in input order"""
    order = np.lexsort((minor, major))               # stable: equal keys keep input order
    M, m, v = major[order], minor[order], val[order]
    head = np.ones(len(v), bool)
    head[1:] = (M[1:] != M[:-1]) | (m[1:] != m[:-1])
    out_val = []
    for i in np.flatnonzero(head):
        acc, j = v[i], i + 1
        while j < len(v) and not head[j]:
            acc = acc + v[j]; j += 1
        out_val.append(acc)
    colptr = np.zeros(n_major + 1, np.int64)
    np.cumsum(np.bincount(M[head], minlength=n_major), out=colptr[1:])
    return colptr, m[head].astype(np.int32), np.array(out_val, np.float64)


@pytest.mark.parametrize("dups", [False, True])
def test_device_build_from_coordinates(gpu, dups):
    """bpmf_gpu_load_coo / _load_test_coo: both sides' compressed matrices built on the device from one
    shuffled coordinate list are bit-identical to the host build — indices, values (duplicates summed in input order),
    column pointers, mean_rating (sequential sum in storage order) — and a chain sampled from them is bit-identical to
    one loaded through bpmf_gpu_load_side."""
    K = 32
    rng = np.random.default_rng(5 + dups)
    nr, nc, n = 700, 450, 30000
    rows = rng.integers(0, nr, size=n).astype(np.int32); cols = rng.integers(0, nc, size=n).astype(np.int32)
    keep = (cols != 18) & (rows != 3)                # an empty column and an empty row
    rows, cols = rows[keep], cols[keep]
    if not dups:
        _, first = np.unique(rows.astype(np.int64) * nc + cols, return_index=True)
        first = rng.permutation(first)               # file order is arbitrary
        rows, cols = rows[first], cols[first]
    else:                                            # triples of the same entry far apart in the input
        rows = np.concatenate([rows, rows[:500], rows[:200]]); cols = np.concatenate([cols, cols[:500], cols[:200]])
    vals = rng.normal(3.5, 1.0, size=len(rows))
    t_rows = rng.integers(0, nr, size=800).astype(np.int32); t_cols = rng.integers(0, nc, size=800).astype(np.int32)
    _, first = np.unique(t_rows.astype(np.int64) * nc + t_cols, return_index=True)
    first = rng.permutation(first)
    t_rows, t_cols = t_rows[first], t_cols[first]
    t_vals = rng.normal(3.5, 1.0, size=len(t_rows))

    dev = gpu.Context(K)
    dev.load_coo(nr, nc, rows, cols, vals)
    dev.load_test_coo(t_rows, t_cols, t_vals)
    host = gpu.Context(K)
    for side, (n_major, major, minor, tmaj, tmin) in ((MOVIES, (nc, cols, rows, t_cols, t_rows)), (USERS, (nr, rows, cols, t_rows, t_cols))):
        colptr, idx, val = _host_build(n_major, major, minor, vals)
        mean = float(np.cumsum(val)[-1] / len(val))  # sequential sum in storage order (sample.cpp:183)
        nnz, dmean, dptr, didx, dval = dev.get_side(side)
        assert nnz == len(val) and dmean == mean
        assert np.array_equal(dptr, colptr) and np.array_equal(didx, idx) and dval.tobytes() == val.tobytes()
        tptr, tidx, tval = _host_build(n_major, tmaj, tmin, t_vals)
        tn, _, dtptr, dtidx, dtval = dev.get_side(side, test=True)
        assert tn == len(tval) and np.array_equal(dtptr, tptr) and np.array_equal(dtidx, tidx) and dtval.tobytes() == tval.tobytes()
        host.load_side(side, n_major, nr if side == MOVIES else nc, colptr, idx, val, mean)
        host.load_test(side, tptr, tidx, tval)
    for it in range(3):
        for side in (MOVIES, USERS):
            dev.sample(side, 2.0, gpu.KERNEL_AUTO)
            host.sample(side, 2.0, gpu.KERNEL_AUTO)
            assert dev.get_items(side).tobytes() == host.get_items(side).tobytes()
        assert dev.predict(MOVIES, 1) == host.predict(MOVIES, 1)
    # an index out of range is refused, and the context stays usable
    bad = rows.copy(); bad[5] = nr
    with pytest.raises(Exception):
        dev.load_coo(nr, nc, bad, cols, vals)
    dev.load_coo(nr, nc, rows, cols, vals)
    for ctx in (dev, host):
        ctx.close()
