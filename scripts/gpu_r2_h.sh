#!/bin/bash
# round 2, call H: same-box A/B of the rank-one DMMA column step (15220) against the default (3220), twice each
set -x
out=gpurun_out/r2h
mkdir -p $out
timeout 600 python bench_micro/tune_stream.py 3220 15220 3220 15220 > $out/tune.log 2>&1; grep cfg $out/tune.log
TUNE_RANGE_DIV=8 timeout 600 python bench_micro/tune_stream.py 3220 15220 3220 15220 > $out/tune_div8.log 2>&1; grep cfg $out/tune_div8.log
