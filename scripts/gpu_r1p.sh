set -x
mkdir -p gpurun_out/r1p
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1p/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1p/pytest.log
tail -5 gpurun_out/r1p/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1p/bench.json 2> gpurun_out/r1p/bench.err
cat gpurun_out/r1p/bench.json
BPMF_NO_HYPER_OVERLAP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1p/bench_nooverlap.json 2> gpurun_out/r1p/bench2.err
cat gpurun_out/r1p/bench_nooverlap.json
