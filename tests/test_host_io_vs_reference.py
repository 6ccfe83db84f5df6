"""The host loaders / writers (bpmf_b200/host/io.cpp, through io_tool) against the REFERENCE'S OWN file-format code:
c++/io.cpp + c++/gzstream.cpp compiled unmodified from /root/reference into oracle/_ref/libbpmf_ref_io.so against the
stand-in Eigen headers of oracle/shim/ (oracle/Makefile target `ref`, oracle/ref_io_harness.cpp). Every input file is
converted by both, and the outputs must be the same bytes (.gz outputs: the same bytes after decompression). CPU only."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

from oracle import reference as ref_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "bpmf_b200", "host")
TOOL = os.path.join(HOST, "io_tool")
REF_IO = os.path.join(ROOT, "oracle", "_ref", "libbpmf_ref_io.so")

MTX = """%%MatrixMarket matrix coordinate real general
% a comment

6 5 9
1 1 2.5
3 1 -1
6 5 1e2
2 2 0
2 3 4
2 3 0.5
4 4 7
5 2 3.25
1 5 -0.125
"""
PATTERN = """%%MatrixMarket matrix coordinate pattern general
4 3 5
1 1
4 3
2 2
3 1
1 3
"""
ARRAY = """%%MatrixMarket matrix array real general
3 2
1.5
-2
0.25
4
5e-3
6
"""


@pytest.fixture(scope="module")
def both():
    if not os.path.exists(REF_IO) and not ref_mod.build():
        pytest.skip("no reference sources and no prebuilt oracle/_ref/libbpmf_ref_io.so")
    subprocess.check_call(["make", "-C", HOST, "-s", "io_tool"])
    lib = C.CDLL(REF_IO)
    lib.bpmf_ref_io_convert.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    lib.bpmf_ref_io_error.restype = C.c_char_p
    lib.bpmf_ref_io_info.argtypes = [C.c_char_p] + [C.POINTER(C.c_longlong)] * 3 + [C.POINTER(C.c_double)]
    return lib


def _bytes(path):
    raw = open(path, "rb").read()
    return gzip.decompress(raw) if path.endswith(".gz") else raw


def _convert_both(lib, kind, src, tmp_path, out_ext):
    mine = str(tmp_path / ("mine" + out_ext))
    theirs = str(tmp_path / ("ref" + out_ext))
    r = subprocess.run([TOOL, kind, src, mine], capture_output=True, text=True)
    rc = lib.bpmf_ref_io_convert(src.encode(), theirs.encode(), int(kind == "dense"))
    assert (r.returncode == 0) == (rc == 0), (src, out_ext, r.stderr, lib.bpmf_ref_io_error())
    if rc == 0:
        assert _bytes(mine) == _bytes(theirs), (src, out_ext)
    return rc == 0


SPARSE_OUT = [".sdm", ".mtx", ".mm", ".sbm", ".sdm.gz", ".mtx.gz"]
DENSE_OUT = [".ddm", ".csv", ".mtx", ".ddm.gz", ".csv.gz"]


def test_sparse_files_convert_identically(both, tmp_path):
    """coordinate .mtx (unsorted, a duplicate, an explicit zero, comments), pattern .mtx, a larger random matrix; every
    sparse output format; then each output read back by both and converted once more."""
    srcs = []
    for name, text in (("a.mtx", MTX), ("p.mtx", PATTERN), ("p.mm", PATTERN)):
        p = tmp_path / name
        p.write_text(text)
        srcs.append(str(p))
    rng = np.random.default_rng(4)
    n = 3000
    r, c = rng.integers(1, 201, n), rng.integers(1, 151, n)
    v = np.round(rng.normal(3.5, 1.0, n), 6)
    big = tmp_path / "big.mtx"
    big.write_text("%%MatrixMarket matrix coordinate real general\n" + "200 150 %d\n" % n
                   + "".join("%d %d %.17g\n" % t for t in zip(r, c, v)))
    srcs.append(str(big))
    with gzip.open(tmp_path / "big.mtx.gz", "wb") as f:
        f.write(big.read_bytes())
    srcs.append(str(tmp_path / "big.mtx.gz"))
    for src in srcs:
        for ext in SPARSE_OUT:
            assert _convert_both(both, "sparse", src, tmp_path, ext), (src, ext)
            # second generation: the file just written (by the reference) read by both
            gen2 = str(tmp_path / ("gen2" + ext))
            os.replace(str(tmp_path / ("ref" + ext)), gen2)
            for ext2 in (".sdm", ".mtx"):
                assert _convert_both(both, "sparse", gen2, tmp_path, ext2), (src, ext, ext2)


def test_dense_files_convert_identically(both, tmp_path):
    srcs = []
    p = tmp_path / "d.mtx"
    p.write_text(ARRAY)
    srcs.append(str(p))
    rng = np.random.default_rng(5)
    X = rng.normal(size=(7, 5))
    csv = tmp_path / "x.csv"
    csv.write_text("7\n5\n" + "\n".join(",".join("%.17g" % x for x in row) for row in X) + "\n")
    srcs.append(str(csv))
    for src in srcs:
        for ext in DENSE_OUT:
            assert _convert_both(both, "dense", src, tmp_path, ext), (src, ext)
            gen2 = str(tmp_path / ("gen2" + ext))
            os.replace(str(tmp_path / ("ref" + ext)), gen2)
            for ext2 in (".ddm", ".csv"):
                assert _convert_both(both, "dense", gen2, tmp_path, ext2), (src, ext, ext2)


def test_both_refuse_the_same_inputs(both, tmp_path):
    """a dense file asked for as sparse and the other way round, an unknown extension, a missing file"""
    (tmp_path / "d.mtx").write_text(ARRAY)
    (tmp_path / "s.mtx").write_text(MTX)
    (tmp_path / "x.foo").write_text("1 2 3\n")
    cases = [("sparse", "d.mtx", ".sdm"), ("dense", "s.mtx", ".ddm"), ("sparse", "x.foo", ".sdm"), ("sparse", "nope.mtx", ".sdm"),
             ("sparse", "s.mtx", ".foo")]
    for kind, name, ext in cases:
        assert not _convert_both(both, kind, str(tmp_path / name), tmp_path, ext), (kind, name, ext)
