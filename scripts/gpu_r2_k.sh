#!/bin/bash
# round 2, call K: variant bit-identity test + K=32 priors with heavy items, real-data timing
set -x
out=gpurun_out/r2k
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants or heavy or propagated" > $out/pytest.log 2>&1; tail -4 $out/pytest.log
timeout 600 python bench_micro/real_data_timing.py > $out/real_data_timing.log 2>&1; grep -E "ml100k|chembl20" $out/real_data_timing.log
