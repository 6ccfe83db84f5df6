"""Host logic of the CTA-per-item kernel (K = 16 m, K != 32): the trailing-update schedule of the blocked LDL^T. For every
block column kb the quads dealt to the warps must cover each trailing tile exactly once — (I, J) with kb < J <= I < NB
except the next diagonal tile (warp 0 factorises it), plus block row NB (the right-hand side) for kb < J < NB — and name
the right panel tiles as their A / B operands. No GPU needed."""
import ctypes

import numpy as np
import pytest


def tile(I, J):
    return (I * (I + 1) // 2 + J) * 64


@pytest.fixture(scope="module")
def lib():
    import bpmf_b200
    return bpmf_b200.load_library()


@pytest.mark.parametrize("K", [16, 48, 64, 80, 96, 112, 128])
def test_every_trailing_tile_is_updated_exactly_once(lib, K):
    NB, NWB = K // 8, K // 16
    buf = np.empty(8 * 64, np.int32)
    for kb in range(NB):
        seen = {}
        loads = []
        for w in range(NWB):
            n = lib.bpmf_gpu_debug_block_schedule(K, kb, w, buf.ctypes.data_as(ctypes.c_void_p), 64)
            assert 0 <= n <= 64
            if NWB > 1 and w == 0:
                assert n == 0          # warp 0 is busy with the next diagonal tile
            loads.append(n)
            for q in buf[: 8 * n].reshape(n, 8):
                a0, a1, b0, b1 = (int(x) for x in q[:4])
                for u, c in enumerate(int(x) for x in q[4:]):
                    if c < 0:
                        continue
                    assert c not in seen, (kb, w, c)
                    seen[c] = (a1 if u >= 2 else a0, b1 if u & 1 else b0)
        want = {}
        for I in range(kb + 1, NB + 1):
            for J in range(kb + 1, min(I, NB - 1) + 1):
                if (I, J) != (kb + 1, kb + 1):
                    want[tile(I, J)] = (tile(I, kb), tile(J, kb))
        assert seen == want, (K, kb)
        if NWB > 2 and sum(loads) >= NWB - 1:
            assert max(loads[1:]) - min(loads[1:]) <= 1      # round-robin over warps 1 .. NWB-1


def test_bad_arguments(lib):
    buf = np.empty(8, np.int32)
    p = buf.ctypes.data_as(ctypes.c_void_p)
    assert lib.bpmf_gpu_debug_block_schedule(32, 0, 0, p, 1) == -1     # K = 32 has its own kernel
    assert lib.bpmf_gpu_debug_block_schedule(40, 0, 0, p, 1) == -1
    assert lib.bpmf_gpu_debug_block_schedule(64, 8, 0, p, 1) == -1
    assert lib.bpmf_gpu_debug_block_schedule(64, 0, 4, p, 1) == -1
