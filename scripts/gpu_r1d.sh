set -x
mkdir -p gpurun_out/r1d
timeout 900 python bench_micro/tune_stream.py 3216 13216 23216 23220 43216 103216 183216 343216 583216 83216 163216 323216 563216 > gpurun_out/r1d/tune.log 2>&1
cat gpurun_out/r1d/tune.log
