#!/bin/bash
# What is run on the B200 box for a round's evidence (under gpurun, from the repository root):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh'
# Everything lands in gpurun_out/round/; the summaries that are judged are copied into profiles/ (scripts/ncu_summary.py).
set -x
out=gpurun_out/round
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.max.memory,power.limit --format=csv > $out/smi.txt; nproc >> $out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
# per-launch device times of the same command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > $out/bench_under_ncu.log 2>&1
# the dominant kernel, full set
timeout 900 ncu --set full --clock-control none --import-source on -k regex:items_stream32v3 -s 4 -c 1 -o $out/stream_v3_full \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > $out/ncu_full.log 2>&1
tail -3 $out/pytest.log; tail -2 $out/smoke.log; cat $out/bench_n1.json
