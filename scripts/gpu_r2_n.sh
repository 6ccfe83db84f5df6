#!/bin/bash
# round 2: A/B of the stream kernel's tail variants (TV bit mask, chol3_block_column), all bit-identical
set -x
out=gpurun_out/r2n
mkdir -p $out
CFGS="${CFGS:-3220 17220 18220 20220 21220 22220 3220 17220 18220 20220 21220 22220}"
timeout 600 python bench_micro/tune_stream.py $CFGS > $out/tune_tv2.log 2>&1; grep -E "cfg|rror" $out/tune_tv2.log
TUNE_RANGE_DIV=8 timeout 600 python bench_micro/tune_stream.py $CFGS > $out/tune_tv2_eighth.log 2>&1; grep -E "cfg|rror" $out/tune_tv2_eighth.log
