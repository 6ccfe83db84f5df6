set -x
mkdir -p gpurun_out/r1r
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1r/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1r/pytest.log
tail -5 gpurun_out/r1r/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1r/bench.json 2> gpurun_out/r1r/bench.err
cat gpurun_out/r1r/bench.json
