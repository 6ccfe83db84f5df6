set -x
mkdir -p gpurun_out/r2a
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not block" > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
tail -3 gpurun_out/r2a/pytest.log
timeout 300 python bench_micro/tune_stream.py 0 3216 > gpurun_out/r2a/tune.log 2>&1; cat gpurun_out/r2a/tune.log
timeout 900 python bench_micro/run_cli_synth.py synthA-1Mx1M-100Mnnz-K32 6 1 > gpurun_out/r2a/cli_synthA.log 2>&1; tail -25 gpurun_out/r2a/cli_synthA.log
