set -x
mkdir -p gpurun_out/r1t
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1t/smoke.log 2>&1; tail -2 gpurun_out/r1t/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1t/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1t/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:items_stream32v3 -s 4 -c 2 -o gpurun_out/r1t/v3_2x20_full python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1t/ncu_full.log 2>&1
tail -2 gpurun_out/r1t/ncu_full.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r1t/smi.txt; nproc >> gpurun_out/r1t/smi.txt; lscpu | grep "Model name" >> gpurun_out/r1t/smi.txt
