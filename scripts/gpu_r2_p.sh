#!/bin/bash
# round 2, multi-GPU: the reductions on the auxiliary stream — parity and step time.  gpurun --gpus N -- 'bash scripts/gpu_r2_p.sh'
set -x
N=$(nvidia-smi -L | wc -l)
out=gpurun_out/r2p_n$N
mkdir -p $out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -4 $out/pytest.log
fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 tests/multi_gpu_worker.py push > $out/worker_push.log 2>&1; tail -2 $out/worker_push.log
python -c "
import sys; sys.path.insert(0,'.')
from bpmf_b200 import synthetic
synthetic.workload('synthA-1Mx1M-100Mnnz-K32', cache_dir='/dev/shm', verbose=False)"
for mode in "BPMF_STATS_MAIN=1" "BPMF_RESERVE_SMS=2" "BPMF_RESERVE_SMS=1" "BPMF_RESERVE_SMS=3"; do
  env $mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench_micro/multi_step_timing.py 2>&1 | grep GPUs >> $out/step_timing.log
done
cat $out/step_timing.log
