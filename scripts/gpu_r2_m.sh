#!/bin/bash
# round 2: where the hyper draw's time goes (phase probes), and the A/B of the stream kernel's micro-optimised variant (17220)
set -x
out=gpurun_out/r2m
mkdir -p $out
BPMF_B200_LIB=$PWD/bpmf_b200/libbpmf_b200_prof.so timeout 300 python bench_micro/hyper_timing.py > $out/hyper_prof.log 2>&1
grep -E "^hyper K=(10|32|64|128) " $out/hyper_prof.log | sort | uniq -c | sort -k3,3 -k1,1nr | awk '!seen[$3]++' | cut -c1-400
timeout 600 python bench_micro/tune_stream.py 3220 17220 3220 17220 > $out/tune_tv2.log 2>&1; grep cfg $out/tune_tv2.log
TUNE_RANGE_DIV=8 timeout 600 python bench_micro/tune_stream.py 3220 17220 3220 17220 > $out/tune_tv2_eighth.log 2>&1; grep cfg $out/tune_tv2_eighth.log
