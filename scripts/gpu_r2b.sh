set -x
mkdir -p gpurun_out/r2b
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest.log
tail -12 gpurun_out/r2b/pytest.log
timeout 300 python bench_micro/tune_stream.py 0 3216 > gpurun_out/r2b/tune.log 2>&1; cat gpurun_out/r2b/tune.log
