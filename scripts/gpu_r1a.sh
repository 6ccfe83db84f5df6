set -x
mkdir -p gpurun_out/r1a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1a/smi.txt
nproc >> gpurun_out/r1a/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1a/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1a/bench.json 2> gpurun_out/r1a/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:items_stream32 -s 2 -c 2 -o gpurun_out/r1a/stream_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1a/ncu.log 2>&1
tail -3 gpurun_out/r1a/pytest.log; cat gpurun_out/r1a/bench.json
