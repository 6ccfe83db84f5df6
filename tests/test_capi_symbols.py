"""The C-ABI library must load and export every symbol include/bpmf_gpu.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bpmf_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bpmf_gpu_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    from bpmf_b200 import build
    return build.build()


def test_header_declares_something():
    names = _declared()
    assert "bpmf_gpu_sample" in names and "bpmf_gpu_create" in names and len(names) >= 25


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    for name in _declared():
        assert hasattr(lib, name), name


def test_python_binding_covers_the_header(built):
    import bpmf_b200
    assert sorted(bpmf_b200.SYMBOLS) == _declared()
    bpmf_b200.load_library()


def test_no_silent_cpu_fallback(built):
    """Without a GPU the product refuses to run; with one it must create a context."""
    import bpmf_b200
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        bpmf_b200.Context(32).close()
    else:
        with pytest.raises(bpmf_b200.BpmfGpuError) as ei:
            bpmf_b200.Context(32)
        assert ei.value.code == 5 and "no CPU fallback" in str(ei.value)


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "bpmf_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                for needle in ("bpmf_oracle", "import oracle", "from oracle", "oracle/", "liboracle"):
                    assert needle not in txt, (fn, needle)
