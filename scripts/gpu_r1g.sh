set -x
mkdir -p gpurun_out/r1g
BPMF_STREAM_CFG=6216 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r1g/pytest_bulk.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1g/pytest_bulk.log
tail -5 gpurun_out/r1g/pytest_bulk.log
timeout 600 python bench_micro/tune_stream.py 3216 6216 6220 6315 6411 16216 26216 586216 > gpurun_out/r1g/tune.log 2>&1
cat gpurun_out/r1g/tune.log
