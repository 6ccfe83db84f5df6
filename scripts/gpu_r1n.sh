set -x
mkdir -p gpurun_out/r1n
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r1n/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1n/pytest.log
tail -5 gpurun_out/r1n/pytest.log
timeout 300 python bench_micro/tune_stream.py 1003216 3216 3003216 9003216 17003216 33003216 5003220 5003218 > gpurun_out/r1n/tune.log 2>&1
cat gpurun_out/r1n/tune.log
