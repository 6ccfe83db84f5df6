#!/bin/bash
# round 2: hyper draw with the speculative gamma walk — parity (hyper tests + chains), timing, phases
set -x
out=gpurun_out/r2r
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x -k "hyper or chain or golden or real or cli" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -4 $out/pytest.log
timeout 300 python bench_micro/hyper_timing.py > $out/hyper_timing.log 2>&1; cat $out/hyper_timing.log
BPMF_B200_LIB=$PWD/bpmf_b200/libbpmf_b200_prof.so timeout 300 python bench_micro/hyper_timing.py > $out/hyper_prof.log 2>&1
grep -E "^hyper K=" $out/hyper_prof.log | sort | uniq -c | sort -k3,3 -k1,1nr | awk '!seen[$3]++' | cut -c1-330
timeout 300 python bench_micro/real_data_timing.py > $out/real_data_timing.log 2>&1; cat $out/real_data_timing.log
