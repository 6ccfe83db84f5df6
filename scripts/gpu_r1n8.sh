set -x
mkdir -p gpurun_out/r1n8
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r1n8/bench_n${N}_push.json 2> gpurun_out/r1n8/bench_n${N}_push.err
cat gpurun_out/r1n8/bench_n${N}_push.json; tail -3 gpurun_out/r1n8/bench_n${N}_push.err
