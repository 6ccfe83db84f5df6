#!/bin/bash
# round 2, call I: what a sweep costs besides its item kernel; launch list of a bench step
set -x
out=gpurun_out/r2i
mkdir -p $out
timeout 600 python bench_micro/sweep_overhead.py 1 8 > $out/sweep_overhead.log 2>&1; grep range $out/sweep_overhead.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > $out/bench_under_ncu.log 2>&1
grep -c "gpu__time_duration" $out/launches.csv
