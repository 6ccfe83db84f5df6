set -x
mkdir -p gpurun_out/r1f
BPMF_STREAM_CFG=48083 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r1f/pytest_v4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1f/pytest_v4.log
tail -5 gpurun_out/r1f/pytest_v4.log
timeout 900 python bench_micro/tune_stream.py 3216 48083 48082 48122 58122 58083 44123 46103 412082 148083 148122 144123 > gpurun_out/r1f/tune.log 2>&1
cat gpurun_out/r1f/tune.log
