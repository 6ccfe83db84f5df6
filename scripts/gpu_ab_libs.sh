#!/bin/bash
# A/B of two builds of the library (bpmf_b200/libbpmf_b200_prev.so vs the current one) on the stream kernel, interleaved;
# whole range and an eighth of it (a rank's share at 8 GPUs).  CFGS = tuning configurations (default: the product kernel)
set -x
out=gpurun_out/ab
mkdir -p $out; rm -f $out/ab.log
CFGS="${CFGS:-3220 3220}"
for div in 1 8; do
  for rep in 1 2; do
    TUNE_RANGE_DIV=$div BPMF_B200_LIB=$PWD/bpmf_b200/libbpmf_b200_prev.so timeout 300 python bench_micro/tune_stream.py $CFGS 2>&1 | grep cfg | sed "s/^/prev 1\/$div /" >> $out/ab.log
    TUNE_RANGE_DIV=$div timeout 300 python bench_micro/tune_stream.py $CFGS 2>&1 | grep cfg | sed "s/^/curr 1\/$div /" >> $out/ab.log
  done
done
cat $out/ab.log
