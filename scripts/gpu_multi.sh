#!/bin/bash
# Multi-GPU evidence:  /usr/local/graft/bin/gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_multi.sh'
set -x
out=gpurun_out/multi
mkdir -p $out
N=$(nvidia-smi -L | wc -l)
for ex in push allgather; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 \
      tests/multi_gpu_worker.py $ex > $out/worker_$ex.log 2>&1; tail -2 $out/worker_$ex.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $N --steps 10 --warmup 3 > $out/bench_n${N}_push.json 2> $out/bench_n${N}_push.err
cat $out/bench_n${N}_push.json
