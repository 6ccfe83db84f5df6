set -x
mkdir -p gpurun_out/r1m
BPMF_STREAM_CFG=10212 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r1m/pytest_v8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1m/pytest_v8.log
tail -5 gpurun_out/r1m/pytest_v8.log
timeout 300 python bench_micro/tune_stream.py 3216 10212 10211 10310 10208 > gpurun_out/r1m/tune.log 2>&1
cat gpurun_out/r1m/tune.log
