#!/bin/bash
# round 2: compute-sanitizer (memcheck, racecheck, synccheck) over the K = 32 paths that changed in the second half of the round
set -x
out=gpurun_out/sanitize
mkdir -p $out
SEL="two_ranks_on_one_gpu or (item_update_one_sweep and 32) or full_run_movielens_shaped_k32 or heavy_items_chunked_path or stats_and_cov or hyper_draw_zero or sample_host_matches"
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 97 --target-processes application-only python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $out/$tool.log 2>&1
  echo "$tool rc=$?" >> $out/summary.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $out/$tool.log | tail -4 >> $out/summary.log
done
cat $out/summary.log
