set -x
mkdir -p gpurun_out/r1u
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1u/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1u/pytest.log
tail -15 gpurun_out/r1u/pytest.log
