set -x
mkdir -p gpurun_out/r1b
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1b/pytest.log
tail -15 gpurun_out/r1b/pytest.log
timeout 900 python bench_micro/tune_stream.py 2216 3216 3220 3218 3315 3411 2220 > gpurun_out/r1b/tune.log 2>&1
cat gpurun_out/r1b/tune.log
