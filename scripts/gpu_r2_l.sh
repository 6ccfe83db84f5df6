#!/bin/bash
# round 2: hyper kernel with its work arrays in shared memory — parity, A/B timing (threads x placement), the small datasets
set -x
out=gpurun_out/r2l
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -4 $out/pytest.log
for t in 128 256 512 1024; do
  echo "== threads $t" >> $out/hyper_timing.log
  BPMF_HYPER_THREADS=$t timeout 300 python bench_micro/hyper_timing.py >> $out/hyper_timing.log 2>&1
done
echo "== threads 1024, global scratch (before)" >> $out/hyper_timing.log
BPMF_HYPER_THREADS=1024 BPMF_HYPER_GLOBAL_SCRATCH=1 timeout 300 python bench_micro/hyper_timing.py >> $out/hyper_timing.log 2>&1
echo "== default" >> $out/hyper_timing.log
timeout 300 python bench_micro/hyper_timing.py >> $out/hyper_timing.log 2>&1
cat $out/hyper_timing.log
timeout 300 python bench_micro/real_data_timing.py > $out/real_data_timing.log 2>&1; cat $out/real_data_timing.log
BPMF_HYPER_THREADS=1024 BPMF_HYPER_GLOBAL_SCRATCH=1 timeout 300 python bench_micro/real_data_timing.py > $out/real_data_timing_before.log 2>&1; cat $out/real_data_timing_before.log
