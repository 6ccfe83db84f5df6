"""bpmf_b200 — B200 (sm_100a) implementation of the BPMF Gibbs sweep behind the reference's Sys interface.

The product is libbpmf_b200.so (hand-written CUDA, C ABI in include/bpmf_gpu.h) plus the C++ host code under
bpmf_b200/host (the `bpmf` executable: CLI, loaders, Sys / CUDA_Sys). This Python package only binds the
C ABI for tests and bench.py; nothing here computes on the CPU.
"""
from .capi import (Context, BpmfGpuError, load_library, SO_PATH, SYMBOLS, MOVIES, USERS,  # noqa: F401
                   KERNEL_AUTO, KERNEL_EXACT, KERNEL_STREAM, KERNEL_BLOCK, STREAM_KERNEL_NAME)
