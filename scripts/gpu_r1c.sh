set -x
mkdir -p gpurun_out/r1c
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1c/pytest.log
tail -15 gpurun_out/r1c/pytest.log
timeout 900 python bench_micro/tune_stream.py 3216 4216 5216 6216 3216 > gpurun_out/r1c/tune.log 2>&1
cat gpurun_out/r1c/tune.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:items_stream32v3 -s 4 -c 1 -o gpurun_out/r1c/v3_full python bench_micro/tune_stream.py 3216 > gpurun_out/r1c/ncu.log 2>&1
tail -3 gpurun_out/r1c/ncu.log
