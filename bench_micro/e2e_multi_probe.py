"""Where the end-to-end step of a multi-GPU run spends its time (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29546 bench_micro/e2e_multi_probe.py

For the default placement of the ranks and again with every rank bound to the host cores of ITS GPU's NUMA node (pinned
buffers re-allocated after binding), all ranks at the same time:
  * H2D / D2H / both of the rank's 1/N slice of a latent matrix from / to pinned host memory (the PCIe + host memory floor),
  * the phases of bpmf_gpu_sample_host's multi-rank protocol one by one (synchronised after each: no overlap, shows each cost),
  * the e2e step as bench.py times it.
Rank r writes gpurun_out/e2e_probe/rank{r}.log; rank 0 also prints the topology."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bpmf_b200  # noqa: E402
from bpmf_b200 import synthetic  # noqa: E402
from bpmf_b200.sampler import GibbsSampler, MOVIES, USERS  # noqa: E402


def gpu_numa_node(local_rank):
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True).stdout.strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus            # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        cpus = open("/sys/devices/system/node/node%d/cpulist" % max(node, 0)).read().strip()
        return bus, node, cpus
    except Exception as e:      # noqa: BLE001
        return "?", -1, "(%r)" % e


def parse_cpulist(s):
    out = set()
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out.update(range(int(a), int(b) + 1))
        elif part.strip().isdigit():
            out.add(int(part))
    return out


def main():
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    outdir = os.path.join(ROOT, "gpurun_out", "e2e_probe")
    os.makedirs(outdir, exist_ok=True)
    logf = open(os.path.join(outdir, "rank%d.log" % rank), "w")

    def log(*a):
        print(*a, file=logf, flush=True)
        if rank == 0:
            print(*a, file=sys.stderr, flush=True)

    if rank == 0:
        log(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
        log(subprocess.run(["bash", "-c", "nproc; lscpu | grep -i -E 'numa|socket|model name'"], capture_output=True, text=True).stdout)
    bus, node, cpus = gpu_numa_node(local_rank)
    log("rank %d gpu %s numa node %d cpus %s; affinity now %d cpus, running on cpu %s"
        % (rank, bus, node, cpus, len(os.sched_getaffinity(0)), open("/proc/self/stat").read().split()[38]))

    wl = os.environ.get("TUNE_WORKLOAD", "synthA-1Mx1M-100Mnnz-K32")
    if rank == 0:
        ratings, K = synthetic.workload(wl, cache_dir="/dev/shm", verbose=True)
    dist.barrier(device_ids=[local_rank])
    if rank != 0:
        ratings, K = synthetic.workload(wl, cache_dir="/dev/shm")
    gs = GibbsSampler(ratings, K, device=local_rank, alpha=2.0, variant=bpmf_b200.KERNEL_AUTO, exchange="push", with_test=False)
    ctx = gs.ctx
    for _ in range(2):
        gs.step()
    ctx.sync()

    def barrier():
        dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def maxred(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / reps

    def wall(fn):
        """fn + device sync, host wall clock in ms"""
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0)

    def run_all(tag):
        host = [torch.empty(gs.num[s], K, dtype=torch.float64).pin_memory() for s in (MOVIES, USERS)]
        for s in (MOVIES, USERS):
            ctx.get_items_ptr(s, host[s].data_ptr())
        torch.cuda.synchronize()
        dev = torch.empty(max(gs.num), K, dtype=torch.float64, device="cuda")
        s2 = torch.cuda.Stream()
        for side in (MOVIES, USERS):
            lo, hi, _ = gs.range[side]
            mb = (hi - lo) * K * 8 / 1e6
            hs, ds = host[side][lo:hi], dev[lo:hi]
            up = timed(lambda: ds.copy_(hs, non_blocking=True), 20)
            dn = timed(lambda: hs.copy_(ds, non_blocking=True), 20)

            def both():
                s2.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s2):
                    hs.copy_(ds, non_blocking=True)
                dev[:hi - lo].copy_(host[1 - side][:hi - lo], non_blocking=True)
                torch.cuda.current_stream().wait_stream(s2)
            bd = timed(both, 20)
            log("[%s] rank %d side %d slice %.1f MB: H2D %.3f ms (%.1f GB/s)  D2H %.3f ms (%.1f GB/s)  both at once %.3f ms; max over ranks %.3f / %.3f / %.3f ms"
                % (tag, rank, side, mb, up, mb / up, dn, mb / dn, bd, maxred(up), maxred(dn), maxred(bd)))
        # the protocol's phases one by one (every phase synchronised, so nothing overlaps)
        for rep in range(3):
            for side in (MOVIES, USERS):
                o = 1 - side
                lo, hi, _ = gs.range[o]
                barrier()
                t_up = wall(lambda: ctx.upload_push_range(o, lo, hi, host[o].data_ptr()))
                t_b1 = wall(lambda: ctx.peer_barrier(side))
                t_sm = wall(lambda: ctx.sample_host_begin(side, host[side].data_ptr()))
                t_b2 = wall(lambda: ctx.peer_barrier(side))
                t_en = wall(lambda: ctx.sample_host_end(side))
                if rep:
                    log("[%s] rank %d sweep of side %d, phases synchronised: upload+push %.3f  barrier %.3f  sample parts + downloads + own statistics %.3f  barrier %.3f  sums + sync %.3f ms (total %.3f)"
                        % (tag, rank, side, t_up, t_b1, t_sm, t_b2, t_en, t_up + t_b1 + t_sm + t_b2 + t_en))

        def e2e_step():
            for side in (MOVIES, USERS):
                gs.sample_host(side, host[1 - side].data_ptr(), host[side].data_ptr())
        e2e_step()
        ms = maxred(timed(e2e_step, 5))
        dms = maxred(timed(gs.step, 5))
        log("[%s] e2e step %.3f ms, device-resident step %.3f ms (max over ranks)" % (tag, ms, dms))
        del host

    run_all("default placement")
    if node >= 0:
        try:
            os.sched_setaffinity(0, parse_cpulist(cpus))
            log("rank %d bound to node %d (%d cpus)" % (rank, node, len(os.sched_getaffinity(0))))
        except OSError as e:
            log("rank %d: sched_setaffinity failed: %r" % (rank, e))
        run_all("bound to the GPU's NUMA node")
    gs.close()
    barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
