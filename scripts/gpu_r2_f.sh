#!/bin/bash
# round 2, call F: whole GPU suite (block-kernel heavy items and priors, device load path of the executable, finalize kernel)
set -x
out=gpurun_out/r2f
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -30 $out/pytest.log
timeout 600 python bench.py --workload synthB-200Kx200K-50Mnnz-K128 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > $out/bench_synthB.json 2> $out/bench_synthB.err; cat $out/bench_synthB.json | head -c 900
