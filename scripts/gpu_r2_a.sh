#!/bin/bash
# round 2, call A: real-data parity tests + whole GPU suite, fp64 pipe mixing microbenchmark, ncu of the K=128 kernel, baseline bench
set -x
out=gpurun_out/r2a
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.max.memory,power.limit --format=csv > $out/smi.txt; nproc >> $out/smi.txt
timeout 900 python -m pytest tests/test_real_data.py -m gpu -x -q > $out/pytest_real.log 2>&1; echo "rc=$?" >> $out/pytest_real.log
tail -5 $out/pytest_real.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_real_data.py > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 120 bench_micro/bin/fp64_mix > $out/fp64_mix.txt 2>&1; cat $out/fp64_mix.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err; cat $out/bench_n1.json
timeout 600 python bench.py --workload synthB-200Kx200K-50Mnnz-K128 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > $out/bench_synthB.json 2> $out/bench_synthB.err; cat $out/bench_synthB.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:items_block -s 2 -c 1 -o $out/block_k128_full \
    python bench.py --workload synthB-200Kx200K-50Mnnz-K128 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $out/ncu_block.log 2>&1
ls -la $out
