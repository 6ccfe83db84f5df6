set -x
mkdir -p gpurun_out/r1z
BPMF_STREAM_CFG=11220 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not block" > gpurun_out/r1z/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1z/pytest.log
tail -4 gpurun_out/r1z/pytest.log
timeout 300 python bench_micro/tune_stream.py 0 11216 11220 11315 11212 > gpurun_out/r1z/tune.log 2>&1
cat gpurun_out/r1z/tune.log
