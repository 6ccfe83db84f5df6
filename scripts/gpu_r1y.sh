set -x
mkdir -p gpurun_out/r1y
timeout 900 ncu --set full --clock-control none --import-source on -k regex:items_block -s 1 -c 1 -o gpurun_out/r1y/block128 python bench.py --workload synthB-200Kx200K-50Mnnz-K128 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1y/ncu.log 2>&1
tail -3 gpurun_out/r1y/ncu.log
