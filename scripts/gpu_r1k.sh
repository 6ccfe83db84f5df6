set -x
mkdir -p gpurun_out/r1k
timeout 300 python bench_micro/tune_stream.py 3216 9086 9105 9067 19086 19067 > gpurun_out/r1k/tune.log 2>&1
cat gpurun_out/r1k/tune.log
