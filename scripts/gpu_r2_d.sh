#!/bin/bash
# round 2, call D: is the K=32 tail bound by its dependent chain or by its instruction count? (probe build)
set -x
out=gpurun_out/r2d
mkdir -p $out
timeout 600 python bench_micro/tune_stream.py 3220 643220 1283220 1923220 13220 163216 3216 > $out/tune_probes.log 2>&1; cat $out/tune_probes.log
