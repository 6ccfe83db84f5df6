set -x
mkdir -p gpurun_out/r1v
timeout 900 python bench.py --workload synthB-200Kx200K-50Mnnz-K128 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1v/bench_synthB.json 2> gpurun_out/r1v/bench_synthB.err
cat gpurun_out/r1v/bench_synthB.json; tail -3 gpurun_out/r1v/bench_synthB.err
timeout 600 python bench.py --workload ml1m-shaped-6040x3952-1Mnnz-K32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r1v/bench_ml1m.json 2> gpurun_out/r1v/bench_ml1m.err
cat gpurun_out/r1v/bench_ml1m.json
