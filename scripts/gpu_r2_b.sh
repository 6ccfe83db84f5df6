#!/bin/bash
# round 2, call B: warp-role kernel A/B, fixed fp64 mixing microbenchmark
set -x
out=gpurun_out/r2b
mkdir -p $out
timeout 120 bench_micro/bin/fp64_mix > $out/fp64_mix.txt 2>&1; cat $out/fp64_mix.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_sweep or movielens" > $out/pytest.log 2>&1; tail -3 $out/pytest.log
timeout 900 python bench_micro/tune_roles.py > $out/tune_roles.log 2>&1; cat $out/tune_roles.log
