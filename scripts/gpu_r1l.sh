set -x
mkdir -p gpurun_out/r1l
timeout 600 ncu --set full --clock-control none --import-source on -k regex:items_stream32v7 -s 0 -c 1 -o gpurun_out/r1l/v7_gramonly python bench_micro/tune_stream.py 19067 > gpurun_out/r1l/ncu.log 2>&1
tail -3 gpurun_out/r1l/ncu.log
