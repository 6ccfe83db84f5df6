#!/bin/bash
# round 2, multi-GPU evidence:  gpurun --gpus N --timeout 1800 -- 'bash scripts/gpu_r2_multi.sh'
set -x
N=$(nvidia-smi -L | wc -l)
out=gpurun_out/r2bmulti_n$N
mkdir -p $out
if [ "$N" = "2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_all.log 2>&1; echo "rc=$?" >> $out/pytest_all.log; tail -8 $out/pytest_all.log
fi
EXS="push allgather"; [ "$N" = "8" ] && EXS="push"
for ex in $EXS; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 \
      tests/multi_gpu_worker.py $ex > $out/worker_$ex.log 2>&1; tail -3 $out/worker_$ex.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $N --steps 10 --warmup 3 > $out/bench_n${N}.json 2> $out/bench_n${N}.err
tail -5 $out/bench_n${N}.err; cat $out/bench_n${N}.json
