#!/bin/bash
# round 2, call G: rank-one DMMA column updates in the K=32 tail: parity + timing (whole range and an eighth)
set -x
out=gpurun_out/r2g
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_real_data.py -m gpu -x -q > $out/pytest.log 2>&1; tail -5 $out/pytest.log
timeout 600 python bench_micro/tune_stream.py 3220 3216 14220 > $out/tune.log 2>&1; cat $out/tune.log
TUNE_RANGE_DIV=8 timeout 600 python bench_micro/tune_stream.py 3220 > $out/tune_div8.log 2>&1; cat $out/tune_div8.log
