#!/bin/bash
# Builds bpmf_b200/libbpmf_b200_prof.so: the product library with the hyper kernel's phase probes (-DBPMF_HYPER_PROF: clock64 at
# the phase boundaries, printed by the kernel). Used by scripts/gpu_r2_m.sh / gpu_r2_r.sh through BPMF_B200_LIB (bpmf_b200/capi.py).
set -e
cd "$(dirname "$0")/.."
python -m bpmf_b200.build
mkdir -p /tmp/bpmf_prof
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I include -fmad=false -DBPMF_HYPER_PROF \
     -c bpmf_b200/csrc/exact_kernels.cu -o /tmp/bpmf_prof/exact_kernels.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o bpmf_b200/libbpmf_b200_prof.so bpmf_b200/build/capi.o /tmp/bpmf_prof/exact_kernels.o \
     bpmf_b200/build/block_kernel.o bpmf_b200/build/build_kernels.o bpmf_b200/build/stream_kernel.o
echo bpmf_b200/libbpmf_b200_prof.so
