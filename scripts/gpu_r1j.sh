set -x
mkdir -p gpurun_out/r1j
BPMF_STREAM_CFG=9124 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r1j/pytest_v7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1j/pytest_v7.log
tail -5 gpurun_out/r1j/pytest_v7.log
timeout 300 python bench_micro/tune_stream.py 3216 9124 9123 9084 9143 19124 > gpurun_out/r1j/tune.log 2>&1
cat gpurun_out/r1j/tune.log
