#!/usr/bin/env python
"""bench.py — latent-vector samples/sec of the BPMF Gibbs sweep pair (movies sweep + users sweep) at K=32.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--exchange allgather|push]

A "step" = movies.sample(users); users.sample(movies)  (c++/bpmf.cpp:184-185) over the whole synthetic rating matrix:
hyper-parameter draw + fused per-item kernel + exchange (N > 1) + sweep reductions, per side.
`value`   = (N_users + N_movies) * steps / time, CUDA events, max over ranks, latents resident in HBM.
`e2e`     = the same through the C ABI with HOST-resident latent matrices (the reference's Sys::sample(Sys &other)
            reads other.items() from host memory and leaves items() in host memory): per sweep the other side is
            uploaded from pinned host memory and the fresh side downloaded, inside the timed region.
`roofline`= fused item kernel: nnz * K * 8 algorithmic gather bytes per launch / its CUDA-event duration, against
            MEASURED_PEAKS.json:hbm_gbs.
`cpu_baseline` = the CPU oracle (restated reference, OpenMP over items like sample.cpp:352) on a bounded row sample of
            the same matrix, on this box's host cores (N=1, rank 0 only).
--impl reference times only that CPU path (the reference itself cannot be built: no Eigen3/Random123).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"           # stdout carries exactly one JSON line: no "NCCL version" banner

METRIC = "latent-vector samples/sec (U+V sweep) at K=32"
UNIT = "samples/s"
DEFAULT_WORKLOAD = "synthA-1Mx1M-100Mnnz-K32"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.1] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[6]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def host_threads():
    """the host cores this process may use. Explicit, because torch.distributed.run exports OMP_NUM_THREADS=1 to its workers and
    the OpenMP runtime of the CPU arm would silently run on one core."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json:hbm_gbs"
        except (KeyError, ValueError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle on a bounded sample of the same matrix
# ------------------------------------------------------------------------------------------------------------------
class CpuBaseline:
    """Times the oracle's per-item loop (sample.cpp:352-372: OpenMP over items, schedule(guided)) on the first S user
    rows and the first S movie columns of the workload, the other side's full-size latent matrix resident in host
    memory, so every item does exactly the work it does in the full job. S is sized once from a calibration pass so that
    one step() is about budget_s seconds of CPU work (the whole matrix if that fits); the sub-models are built once and
    re-sampled by every step()."""

    def __init__(self, ratings, K, alpha, budget_s=12.0, threads=0):
        from oracle import oracle as o
        self.o, self.ratings, self.K, self.alpha = o, ratings, K, alpha
        # always passed explicitly (num_threads clause of the oracle): torchrun exports OMP_NUM_THREADS=1
        self.threads = host_threads() if threads <= 0 else threads
        self.rng = np.random.Generator(np.random.PCG64(7))
        cal = [self._sub_model(side, 4000) for side in (0, 1)]
        t, items, _ = self._run(cal)                  # calibration pass (also warms the OpenMP team)
        rate = items / t
        self.S = int(max(4000, min(rate * budget_s / 2, max(ratings.nrows, ratings.ncols))))
        self.models = cal if self.S == 4000 else [self._sub_model(side, self.S) for side in (0, 1)]
        self.iter = 3

    def _sub_model(self, side, S):
        # users sample: rows [0,S) of R against all movies; movies sample: columns [0,S) of R against all users
        o, K = self.o, self.K
        n, n_other, ptr, idx, val = self.ratings.side(side)
        S = min(S, n)
        e = int(ptr[S])
        own = np.repeat(np.arange(S, dtype=np.int32), np.diff(ptr[:S + 1]))
        oth = np.asarray(idx[:e], np.int32)
        v = np.asarray(val[:e], np.float64)
        if side == 1:     # users are file rows
            shape, rows, cols = (S, n_other), own, oth
        else:
            shape, rows, cols = (n_other, S), oth, own
        m = o.Oracle(K, shape, rows, cols, v, shape, rows[:1], cols[:1], v[:1], alpha=self.alpha, burnin=0, nthreads=self.threads)
        m.set_items(1 - side, self.rng.normal(0, 0.3, size=(m.num(1 - side), K)))
        m.set_hyper(side, self.rng.normal(0, 0.1, size=K), np.eye(K) * 2.0)
        return m, S, e

    def _run(self, models, it=3):
        tot, items, nnz = 0.0, 0, 0
        for side in (0, 1):
            m, s, e = models[side]
            m.set_iter(side, it)
            t0 = time.perf_counter()
            m.sample_range(side, 0, s)
            tot += time.perf_counter() - t0
            items += s; nnz += e
        return tot, items, nnz

    def step(self):
        """one bounded sample of a U+V sweep pair -> dict(value, seconds, ...)"""
        self.iter += 1
        t, items, nnz = self._run(self.models, self.iter)
        S, r = self.S, self.ratings
        whole = S >= max(r.nrows, r.ncols)
        return {"value": items / t, "unit": UNIT, "cores": int(self.threads), "kind": "port",
                "sample": "%s (%d ratings), per-item loop of the restated reference (oracle/, OpenMP schedule(guided), %d threads), "
                          "other side's full latent matrix in host memory; %.1f s of CPU time"
                          % ("the WHOLE workload: all %d user rows + all %d movie columns" % (r.nrows, r.ncols) if whole else
                             "first %d user rows + first %d movie columns of the workload" % (min(S, r.nrows), min(S, r.ncols)), nnz, self.threads, t),
                "seconds": t, "items": items, "whole_workload": whole}


def cpu_baseline(ratings, K, alpha, budget_s=12.0, threads=0):
    return CpuBaseline(ratings, K, alpha, budget_s, threads).step()


# ------------------------------------------------------------------------------------------------------------------
def make_config(args):
    from bpmf_b200 import synthetic
    nrows, ncols, mean_nnz, K, seed = synthetic.WORKLOADS[args.workload]
    return {"workload": args.workload, "users": nrows, "movies": ncols, "num_latent": K, "alpha": args.alpha,
            "generator": "bpmf_b200/synthetic.py seed %d: Poisson(%g) ratings per user row, %s, planted rank-16 values"
                         % (seed, mean_nnz, "movies drawn with Zipf(s=%g) popularity, duplicates dropped" % synthetic.ZIPF[args.workload]
                            if args.workload in synthetic.ZIPF else "uniform distinct movies"),
            "l2": "%s: per sweep %.0f MB of latent vectors gathered at random + %.0f MB of CSR"
                  % ("inputs larger than L2" if max(nrows, ncols) * K * 8 + nrows * mean_nnz * 12 > 126e6 else
                     "inputs FIT in the 126 MB L2 (no flush between steps: a cache-resident configuration, not the headline one)",
                     max(nrows, ncols) * K * 8 / 1e6, nrows * mean_nnz * 12 / 1e6)}


def reference_arm(args, config):
    """--impl reference: the reference's CPU path (kind "port": the OpenMP restatement in oracle/; the reference itself needs
    Eigen3 + Random123, absent here) on ALL host cores, each step one bounded sample of the workload. Rank 0 only."""
    from bpmf_b200 import synthetic
    nrows, ncols = config["users"], config["movies"]
    ratings, K = synthetic.workload(args.workload, cache_dir=args.cache_dir, verbose=True)
    config["nnz"] = int(ratings.nnz)
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    cb = CpuBaseline(ratings, K, args.alpha, budget_s=per_step_budget)
    vals, base = [], None
    for i in range(args.warmup + args.steps):
        base = cb.step()
        log("reference step %d: %.0f samples/s in %.2f s (%s)" % (i, base["value"], base["seconds"], base["sample"][:70]))
        if i >= args.warmup:
            vals.append(base)
    tot_items = sum(b["items"] for b in vals)
    tot_s = sum(b["seconds"] for b in vals)
    v = tot_items / tot_s
    base = dict(base, value=v)
    base.pop("items", None)
    full_ms = 1e3 * (nrows + ncols) / v
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, len(vals)), "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": base,
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
           "ms_per_whole_workload_step": full_ms,
           "note": "restated reference (oracle/, kind=port): the reference needs Eigen3 + Random123, absent here. ms_per_step is the "
                   "measured time of one step's sample (%s); %d host threads set explicitly"
                   % ("the whole workload" if base["whole_workload"] else "a bounded part of the workload; ms_per_whole_workload_step extrapolates it",
                      base["cores"])}
    print(json.dumps(out), flush=True)
    return 0


def latents_witness(gs, sides):
    """N-invariance witness of the chain state: the run is bit-identical for any GPU count iff these agree"""
    import hashlib
    w = {}
    h = hashlib.sha1()
    for name, side in sides:
        x = gs.items_host(side)
        h.update(np.ascontiguousarray(x).tobytes())
        w[name + "_abs_colmean_sum"] = float(np.abs(x.mean(0)).sum())
        w[name + "_frobenius"] = float(np.sqrt((x * x).sum()))
    w["sha1_V_then_U"] = h.hexdigest()
    return w


def secondary_block(args, rank, local_rank, world, timed_region, dist, workload="synthB-200Kx200K-50Mnnz-K128", steps=3, warmup=2):
    """BASELINE.json configs[4] in the same run: Synthetic B (200K x 200K, 50M ratings, K = 128) through the CTA-per-item
    kernel — ms per U+V step, the item kernel's time, and its fraction of the HBM roofline and of the fp64 tensor (DMMA) rate."""
    import bpmf_b200
    from bpmf_b200 import synthetic
    from bpmf_b200.sampler import GibbsSampler
    nrows, ncols, _, K, _ = synthetic.WORKLOADS[workload]
    if world > 1:
        if rank == 0:
            ratings, K = synthetic.workload(workload, cache_dir=args.cache_dir, verbose=True)
        dist.barrier(device_ids=[local_rank])
        if rank != 0:
            ratings, K = synthetic.workload(workload, cache_dir=args.cache_dir)
    else:
        ratings, K = synthetic.workload(workload, cache_dir=args.cache_dir, verbose=True)
    gs = GibbsSampler(ratings, K, device=local_rank, alpha=args.alpha, variant=bpmf_b200.KERNEL_AUTO, exchange=args.exchange, with_test=False)
    for _ in range(warmup):
        gs.step()
    gs.ctx.sync()
    gs.ctx.items_kernel_time()
    ms = timed_region(gs.step, steps)
    gs.ctx.sync()
    k_ms, k_n = gs.ctx.items_kernel_time()
    k_avg = k_ms / max(1, k_n)
    peak, _ = measured_peaks()
    gbs = ratings.nnz / world * K * 8.0 / (k_avg / 1e3) / 1e9
    tfl = ratings.nnz / world * (K * (K + 1) / 2 + K) * 2.0 / (k_avg / 1e3) / 1e12
    out = {"workload": workload, "num_latent": K, "nnz": int(ratings.nnz), "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms / steps, "value": (nrows + ncols) * steps / (ms / 1e3), "unit": UNIT,
           "kernel": "items_block_kernel<%d>" % (K // 8), "kernel_ms_avg": k_avg, "kernel_launches_timed": k_n,
           "hbm": {"achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak},
           "fp64_tensor": {"achieved": tfl, "peak": 37.0, "unit": "TFLOP/s", "frac": tfl / 37.0,
                           "peak_source": "profiles/r01_fp64_pipes_microbench.txt (DMMA m8n8k4, measured on B200)"}}
    gs.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--exchange", default="push", choices=["allgather", "push"])
    ap.add_argument("--variant", default="auto", choices=["auto", "exact", "stream"])
    ap.add_argument("--alpha", type=float, default=2.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the Synthetic B / K=128 block of the JSON line")
    ap.add_argument("--cache-dir", default="/dev/shm")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from bpmf_b200 import synthetic
    nrows, ncols, mean_nnz, K, seed = synthetic.WORKLOADS[args.workload]
    config = make_config(args)       # the same keys in both arms

    # ---------------------------------------------------------------- reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        return reference_arm(args, config) if rank == 0 else 0

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import bpmf_b200
    from bpmf_b200.sampler import GibbsSampler, MOVIES, USERS

    t_gen = time.time()
    if world > 1:
        if rank == 0:
            ratings, K = synthetic.workload(args.workload, cache_dir=args.cache_dir, verbose=True)
        dist.barrier(device_ids=[local_rank])
        if rank != 0:
            ratings, K = synthetic.workload(args.workload, cache_dir=args.cache_dir)
    else:
        ratings, K = synthetic.workload(args.workload, cache_dir=args.cache_dir, verbose=True)
    config["nnz"] = int(ratings.nnz)
    if rank == 0:
        log("workload %s ready in %.1fs (nnz %d)" % (args.workload, time.time() - t_gen, ratings.nnz))
    variant = {"auto": bpmf_b200.KERNEL_AUTO, "exact": bpmf_b200.KERNEL_EXACT, "stream": bpmf_b200.KERNEL_STREAM}[args.variant]
    gs = GibbsSampler(ratings, K, device=local_rank, alpha=args.alpha, variant=variant, exchange=args.exchange, with_test=True)
    parallelism = "1 gpu" if world == 1 else "items of both factors split over %d gpus, %s exchange" % (world, gs.exchange)
    n_samples = nrows + ncols

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def timed_region(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident sweep pairs
    for _ in range(args.warmup):
        gs.step()
    gs.ctx.sync()
    gs.ctx.items_kernel_time()                       # drop the warm-up launches
    launches0 = gs.ctx.launch_count()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.25)
    t0 = time.time()
    ms = timed_region(gs.step, args.steps)
    t1 = time.time()
    gs.ctx.sync()                                    # surfaces "Cholesky failed" etc.
    launches = gs.ctx.launch_count() - launches0
    k_ms, k_n = gs.ctx.items_kernel_time()
    clk = clocks.stop(t0, t1) if clocks else None
    value = n_samples * args.steps / (ms / 1e3)

    # ---- the reference's loop body also predicts every iteration (c++/bpmf.cpp:184-190): the same steps with both predict calls
    def step_with_predict():
        gs.step()
        gs.predict(burnin=args.warmup)
    p_steps = max(2, min(args.steps, 5))
    ms_pred = timed_region(step_with_predict, p_steps)
    rmse = gs.predict(burnin=args.warmup)[0]

    # ---- roofline of the fused item kernel (both sides' launches averaged)
    peak, peak_src = measured_peaks()
    bytes_per_launch = ratings.nnz / world * K * 8.0          # nnz * K * 8 per sweep, this rank's share
    k_avg_ms = k_ms / max(1, k_n)
    achieved = bytes_per_launch / (k_avg_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": ("items_exact_kernel" if args.variant == "exact" or K % 16 else "items_block_kernel<%d>" % (K // 8) if K != 32 else
                           bpmf_b200.STREAM_KERNEL_NAME),
                "kernel_ms_avg": k_avg_ms, "kernel_launches_timed": k_n, "kernel_share_of_step": k_ms / ms,
                "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peak_src + " (of measured, burst)"}
    if K != 32:
        # above K = 32 the Gram (nnz * K (K + 1) / 2 FMAs on the fp64 tensor cores) outweighs the gather: report it against the
        # measured DMMA peak too (bench_micro/fp64_pipes.cu, profiles/r01_fp64_pipes_microbench.txt: 37.0 TFLOP/s)
        gram_flop = ratings.nnz / world * (K * (K + 1) / 2 + K) * 2.0
        roofline["fp64_tensor"] = {"achieved": gram_flop / (k_avg_ms / 1e3) / 1e12, "peak": 37.0, "unit": "TFLOP/s",
                                   "frac": gram_flop / (k_avg_ms / 1e3) / 1e12 / 37.0,
                                   "peak_source": "profiles/r01_fp64_pipes_microbench.txt (DMMA m8n8k4, measured on B200)"}
    # DRAM bytes of one launch from the committed `ncu --set full` capture: valid for the launch it was taken on (N = 1, this
    # workload, this kernel); at N > 1 a launch covers 1/N of the ratings and no capture exists, so it stays null
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    roofline["traffic_note"] = "null: no ncu capture for this (workload, n_gpus)"
    if world == 1 and os.path.exists(tr):
        try:
            ent = json.load(open(tr)).get(args.workload, {})
            if ent.get("kernel", roofline["kernel"]) == roofline["kernel"] or "kernel" not in ent:
                roofline["traffic"] = ent.get("dram_bytes_per_launch")
                roofline["traffic_note"] = "dram__bytes_read.sum + dram__bytes_write.sum of one launch, %s" % ent.get("source", "profiles/traffic.json")
        except ValueError:
            pass

    # ---- end to end with host-resident latent matrices
    e2e = None
    if not args.no_e2e:
        host = [torch.empty(gs.num[s], K, dtype=torch.float64).pin_memory() for s in (MOVIES, USERS)]
        for s in (MOVIES, USERS):
            gs.ctx.get_items_ptr(s, host[s].data_ptr())
        torch.cuda.synchronize()

        def e2e_step():
            for side in (MOVIES, USERS):
                gs.sample_host(side, host[1 - side].data_ptr(), host[side].data_ptr())

        # iteration counters continue; the latents uploaded are the ones just downloaded, so the chain is unchanged
        e2e_step()
        e_steps = max(2, min(args.steps, 5))
        e_ms = timed_region(e2e_step, e_steps)
        hb = (gs.num[MOVIES] + gs.num[USERS]) * K * 8
        e2e = {"value": n_samples * e_steps / (e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": hb,
               "d2h_bytes_per_step": hb, "ms_per_step": e_ms / e_steps, "steps": e_steps,
               "api": "bpmf_gpu_sample_host (C ABI) with pinned host latent matrices" if world == 1 else
                      "per rank: its slice of the other side uploaded in chunks and pushed to the peers over NVLink while the upload "
                      "continues, sample stages, its fresh slice downloaded in parts while the rest is sampled; host latent matrices "
                      "pinned, one slice per rank (the sum over ranks is the bytes shown)"}

    total_launches = launches
    if world > 1:
        t = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        total_launches = int(t.item())
        # every rank's average item-kernel time: the spread is the load imbalance the cross-GPU barrier waits for
        kt = torch.zeros(world, dtype=torch.float64, device="cuda")
        kt[rank] = k_avg_ms
        dist.all_reduce(kt)
        roofline["kernel_ms_avg_per_rank"] = [round(float(x), 4) for x in kt.tolist()]
    witness = latents_witness(gs, (("V", MOVIES), ("U", USERS))) if rank == 0 else None
    gs.close()
    del gs

    secondary = None
    if not args.no_secondary and args.workload == DEFAULT_WORKLOAD:
        try:
            secondary = secondary_block(args, rank, local_rank, world, timed_region, dist)
        except Exception as e:                       # the headline line must not be lost to the second workload
            secondary = {"error": repr(e)[:300]}
            if world > 1:
                raise

    if rank == 0:
        base = None
        if not args.no_cpu_baseline:
            # rank 0's host cores (at N > 1 the other ranks wait at the barrier below); a shorter sample when other ranks wait
            base = cpu_baseline(ratings, K, args.alpha, budget_s=12.0 if world == 1 else 5.0)
            base.pop("seconds", None); base.pop("items", None)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": config, "parallelism": parallelism, "roofline": roofline,
               "cpu_baseline": base, "e2e": e2e, "gpu_launches": total_launches, "clocks": clk,
               "rmse_after_run": rmse[0], "ratings_per_s": 2.0 * ratings.nnz * args.steps / (ms / 1e3),
               "incl_predict": {"value": n_samples * p_steps / (ms_pred / 1e3), "unit": UNIT, "ms_per_step": ms_pred / p_steps, "steps": p_steps,
                                "what": "the same step plus movies.predict(users); users.predict(movies) — the reference's whole loop body "
                                        "(c++/bpmf.cpp:184-190), which its items/sec counts"},
               "witness": witness, "secondary": secondary,
               "note": "`value` times movies.sample(users); users.sample(movies) only (the U+V sweep pair BASELINE.json's metric names); "
                       "`incl_predict` adds both predict calls; cpu_baseline / --impl reference time the per-item loop of the sweeps without predict"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
