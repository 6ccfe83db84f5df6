#!/bin/bash
# round 2: A/B of two builds of the library (previous vs current) on the stream kernel + the RNG / parity tests
set -x
out=gpurun_out/r2s
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x -k "randn or philox or rng or hyper or chain or golden or sweep or real" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -3 $out/pytest.log
for rep in 1 2; do
  BPMF_B200_LIB=$PWD/bpmf_b200/libbpmf_b200_prev.so timeout 300 python bench_micro/tune_stream.py 3220 3220 2>&1 | grep cfg | sed 's/^/prev /' >> $out/ab.log
  timeout 300 python bench_micro/tune_stream.py 3220 3220 2>&1 | grep cfg | sed 's/^/curr /' >> $out/ab.log
done
cat $out/ab.log
