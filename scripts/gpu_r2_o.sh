#!/bin/bash
# round 2: the reductions of a sweep on the auxiliary stream with reserved SMs — parity and what a sweep costs besides its item kernel
set -x
out=gpurun_out/r2o
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -5 $out/pytest.log
for mode in "BPMF_STATS_MAIN=1" "BPMF_RESERVE_SMS=1" "BPMF_RESERVE_SMS=2" "BPMF_RESERVE_SMS=0"; do
  echo "== $mode" >> $out/sweep_overhead.log
  env $mode timeout 300 python bench_micro/sweep_overhead.py 1 8 >> $out/sweep_overhead.log 2>&1
done
grep -E "==|range" $out/sweep_overhead.log
