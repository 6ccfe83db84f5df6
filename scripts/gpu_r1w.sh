set -x
mkdir -p gpurun_out/r1w
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1w/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1w/pytest.log
tail -25 gpurun_out/r1w/pytest.log
timeout 900 python bench.py --workload synthB-200Kx200K-50Mnnz-K128 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1w/bench_synthB.json 2> gpurun_out/r1w/bench_synthB.err
cat gpurun_out/r1w/bench_synthB.json; tail -3 gpurun_out/r1w/bench_synthB.err
timeout 600 python bench.py --workload ml1m-shaped-6040x3952-1Mnnz-K32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r1w/bench_ml1m.json 2> gpurun_out/r1w/bench_ml1m.err
cat gpurun_out/r1w/bench_ml1m.json
