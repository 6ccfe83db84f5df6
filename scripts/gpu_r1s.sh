set -x
mkdir -p gpurun_out/r1s
BPMF_STREAM_CFG=9124 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r1s/pytest_v7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1s/pytest_v7.log
tail -4 gpurun_out/r1s/pytest_v7.log
timeout 300 python bench_micro/tune_stream.py 0 9124 9123 9084 9143 9086 9105 19124 19086 > gpurun_out/r1s/tune.log 2>&1
cat gpurun_out/r1s/tune.log
