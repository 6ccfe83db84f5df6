#!/bin/bash
# round 2, call E: whole GPU suite after the multi-GPU / load-path / aggregates changes; bench line
set -x
out=gpurun_out/r2e
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -15 $out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err; tail -3 $out/bench_n1.err; cat $out/bench_n1.json
