set -x
mkdir -p gpurun_out/r1x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "block or one_sweep" > gpurun_out/r1x/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1x/pytest.log
tail -4 gpurun_out/r1x/pytest.log
timeout 900 python bench.py --workload synthB-200Kx200K-50Mnnz-K128 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1x/bench_synthB.json 2> gpurun_out/r1x/bench_synthB.err
cat gpurun_out/r1x/bench_synthB.json; tail -3 gpurun_out/r1x/bench_synthB.err
