#!/bin/bash
# round 2, call J: probe build — throughput of the tail alone (no gather, no Gram); real-data timing needs the product build, so not here
set -x
out=gpurun_out/r2j
mkdir -p $out
timeout 600 python bench_micro/tune_stream.py 3220 63220 63216 143220 13220 3220 > $out/tune_tail_only.log 2>&1; grep cfg $out/tune_tail_only.log
