set -x
mkdir -p gpurun_out/r1i
timeout 600 python bench_micro/tune_stream.py 3216 7216 7220 7315 8216 8220 8315 > gpurun_out/r1i/tune.log 2>&1
cat gpurun_out/r1i/tune.log
