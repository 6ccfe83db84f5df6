#!/bin/bash
# round 2, 8 GPUs: step time with the reductions on the main / auxiliary stream (sha1 of the latents must agree)
set -x
N=$(nvidia-smi -L | wc -l)
out=gpurun_out/r2q_n$N
mkdir -p $out
python -c "
import sys; sys.path.insert(0,'.')
from bpmf_b200 import synthetic
synthetic.workload('synthA-1Mx1M-100Mnnz-K32', cache_dir='/dev/shm', verbose=False)"
for mode in "BPMF_STATS_MAIN=1" "BPMF_RESERVE_SMS=2" "BPMF_RESERVE_SMS=3"; do
  env $mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench_micro/multi_step_timing.py 2>&1 | grep -E "GPUs|rror" >> $out/step_timing.log
done
cat $out/step_timing.log
