"""CPU test of bench.py's reference arm: exactly ONE JSON line on stdout with the keys the driver reads (the GPU arm
prints the same line plus roofline / clocks; it needs a B200)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(tmp_path):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small-20Kx20K-1Mnnz-K32",
           "--steps", "1", "--warmup", "0", "--cache-dir", str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("latent-vector samples/sec") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["config"]["workload"] == "small-20Kx20K-1Mnnz-K32"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ms_per_step is the MEASURED time of a step's sample (never an extrapolation): consistent with value on that sample
    assert d["cpu_baseline"]["whole_workload"] is True
    assert abs(d["ms_per_step"] - 1e3 * 40000 / d["value"]) <= 1e-6 * d["ms_per_step"]
    # the same config keys as the GPU arm (the driver compares them): nothing arm-specific inside config
    assert sorted(d["config"]) == ["alpha", "generator", "l2", "movies", "nnz", "num_latent", "users", "workload"]


def test_reference_arm_uses_all_host_cores_under_torchrun(tmp_path):
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm must not inherit that"""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "small-20Kx20K-1Mnnz-K32",
           "--steps", "1", "--warmup", "0", "--cache-dir", str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and d["n_gpus"] == 2


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
