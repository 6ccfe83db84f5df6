#!/bin/bash
# round 2, call C: parity of the changed kernels (trailing update without shuffles, sliced statistics), gather4 + guided claims A/B
set -x
out=gpurun_out/r2c
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/pytest.log 2>&1; tail -5 $out/pytest.log
timeout 600 python bench_micro/tune_stream.py 3220 14220 14224 14216 14316 800003220 400003220 1600003220 800014220 > $out/tune.log 2>&1; cat $out/tune.log
TUNE_RANGE_DIV=8 timeout 600 python bench_micro/tune_stream.py 3220 800003220 400003220 1600003220 14220 800014220 > $out/tune_div8.log 2>&1; cat $out/tune_div8.log
