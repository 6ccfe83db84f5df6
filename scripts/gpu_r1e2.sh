set -x
mkdir -p gpurun_out/r1e2
N=$(nvidia-smi -L | wc -l)
for ex in push allgather; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 tests/multi_gpu_worker.py $ex > gpurun_out/r1e2/worker_$ex.log 2>&1; tail -3 gpurun_out/r1e2/worker_$ex.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r1e2/bench_n${N}_push.json 2> gpurun_out/r1e2/bench_n${N}_push.err
cat gpurun_out/r1e2/bench_n${N}_push.json; tail -3 gpurun_out/r1e2/bench_n${N}_push.err
