"""Builds bpmf_b200/libbpmf_b200.so (sm_100a only) with nvcc. No GPU needed: nvcc cross-compiles.

    python -m bpmf_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libbpmf_b200.so")
OBJ = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include")]
# (source, extra flags). exact_kernels.cu must not contract a*b+c into FMAs: it reproduces the rounding of the
# reference's x86-64 baseline build.
UNITS = [
    ("capi.cu", []),
    ("exact_kernels.cu", ["-fmad=false"]),
    ("block_kernel.cu", ["-DBPMF_BLOCK_PROF"] if os.environ.get("BPMF_BLOCK_PROF") else []),
    ("build_kernels.cu", []),
    ("stream_kernel.cu", ["-DBPMF_STREAM_PROBES"] if os.environ.get("BPMF_STREAM_PROBES") else []),
]
HEADERS = ["common.cuh", "rng.cuh", os.path.join(ROOT, "include", "bpmf_gpu.h")]
if os.environ.get("BPMF_STREAM_PROBES"):     # the experiments of bench_micro/ are compiled into probe builds only
    HEADERS += [os.path.join(ROOT, "bench_micro", f) for f in ("stream_experiments.cuh", "stream_roles.cuh", "stream_experiment_cases.inc")]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [_nvcc()] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    if force or _stale(SO, objs):
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", SO] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
