set -x
mkdir -p gpurun_out/r1q
N=$(nvidia-smi -L | wc -l)
for ex in push allgather; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 --exchange $ex --no-e2e > gpurun_out/r1q/bench_n${N}_$ex.json 2> gpurun_out/r1q/bench_n${N}_$ex.err
cat gpurun_out/r1q/bench_n${N}_$ex.json; tail -2 gpurun_out/r1q/bench_n${N}_$ex.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 tests/multi_gpu_worker.py push > gpurun_out/r1q/worker_push.log 2>&1; tail -2 gpurun_out/r1q/worker_push.log
