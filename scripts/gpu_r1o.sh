set -x
mkdir -p gpurun_out/r1o
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1o/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1o/pytest.log
tail -5 gpurun_out/r1o/pytest.log
timeout 300 python bench_micro/tune_stream.py 3216 3220 0 > gpurun_out/r1o/tune.log 2>&1
cat gpurun_out/r1o/tune.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1o/bench.json 2> gpurun_out/r1o/bench.err
cat gpurun_out/r1o/bench.json
