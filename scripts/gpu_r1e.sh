set -x
mkdir -p gpurun_out/r1e
nvidia-smi -L > gpurun_out/r1e/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1e/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1e/pytest.log
tail -5 gpurun_out/r1e/pytest.log
for ex in allgather push; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 --exchange $ex > gpurun_out/r1e/bench_n2_$ex.json 2> gpurun_out/r1e/bench_n2_$ex.err
cat gpurun_out/r1e/bench_n2_$ex.json; tail -3 gpurun_out/r1e/bench_n2_$ex.err
done
