"""CPU tests of the host loaders / writers (bpmf_b200/host/io.cpp) through io_tool: the formats of c++/io.cpp
(.mtx/.mm text, .sdm/.sbm/.ddm binary, .csv, .gz on top) checked against scipy / numpy readings of the same bytes."""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "bpmf_b200", "host")
TOOL = os.path.join(HOST, "io_tool")


@pytest.fixture(scope="module")
def tool():
    subprocess.check_call(["make", "-C", HOST, "-s", "io_tool"])
    return TOOL


def run(tool, *args):
    return subprocess.run([tool, *args], capture_output=True, text=True)


def read_sdm(path):
    raw = (gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")).read()
    nr, nc, nnz = struct.unpack_from("<QQQ", raw, 0)
    rows = np.frombuffer(raw, "<u4", nnz, 24)
    cols = np.frombuffer(raw, "<u4", nnz, 24 + 4 * nnz)
    vals = np.frombuffer(raw, "<f8", nnz, 24 + 8 * nnz)
    assert len(raw) == 24 + 16 * nnz
    return nr, nc, rows, cols, vals


def read_ddm(path):
    raw = open(path, "rb").read()
    nr, nc = struct.unpack_from("<QQ", raw, 0)
    return np.frombuffer(raw, "<f8", nr * nc, 16).reshape(nc, nr).T   # column-major on disk


MTX = """%%MatrixMarket matrix coordinate real general
% a comment, then an empty line

5 4 7
1 1 2.5
3 1 -1
5 4 1e2
% comment between entries
2 2 0
2 3 4
2 3 0.5
4 4 7
"""


def test_mtx_to_sdm_sums_duplicates_keeps_zeros_and_sorts(tool, tmp_path):
    src = tmp_path / "a.mtx"
    src.write_text(MTX)
    dst = str(tmp_path / "a.sdm")
    assert run(tool, "sparse", str(src), dst).returncode == 0
    nr, nc, rows, cols, vals = read_sdm(dst)
    assert (nr, nc) == (5, 4)
    # column order, 1-based, duplicates (2,3) summed, explicit zero (2,2) kept (c++/io.cpp:282,521)
    assert list(zip(rows, cols, vals)) == [(1, 1, 2.5), (3, 1, -1.0), (2, 2, 0.0), (2, 3, 4.5), (4, 4, 7.0), (5, 4, 100.0)]
    out = run(tool, "info", dst).stdout.split()
    assert out[:3] == ["5", "4", "6"] and float(out[3]) == 113.0


def test_pattern_mtx_and_mm_extension(tool, tmp_path):
    src = tmp_path / "p.mm"
    src.write_text("%%MatrixMarket MATRIX COORDINATE PATTERN GENERAL\n3 3 2\n1 2\n3 3\n")
    dst = str(tmp_path / "p.sdm")
    assert run(tool, "sparse", str(src), dst).returncode == 0
    _, _, rows, cols, vals = read_sdm(dst)
    assert list(zip(rows, cols, vals)) == [(1, 2, 1.0), (3, 3, 1.0)]
    sbm = str(tmp_path / "p.sbm")
    assert run(tool, "sparse", dst, sbm).returncode == 0
    raw = open(sbm, "rb").read()
    assert struct.unpack_from("<QQQ", raw, 0) == (3, 3, 2) and len(raw) == 24 + 8 * 2
    back = str(tmp_path / "q.sdm")
    assert run(tool, "sparse", sbm, back).returncode == 0
    assert list(read_sdm(back)[4]) == [1.0, 1.0]


@pytest.mark.parametrize("gz", [False, True])
def test_random_sparse_round_trips_match_scipy(tool, tmp_path, gz):
    rng = np.random.default_rng(5)
    M = sp.random(60, 45, density=0.1, random_state=7, data_rvs=lambda n: rng.normal(size=n)).tocoo()
    src = str(tmp_path / "r.mtx")
    scipy.io.mmwrite(src, M, precision=17)
    if gz:
        with open(src, "rb") as f, gzip.open(src + ".gz", "wb") as g:
            g.write(f.read())
        src += ".gz"
    sdm = str(tmp_path / ("r.sdm.gz" if gz else "r.sdm"))
    assert run(tool, "sparse", src, sdm).returncode == 0
    nr, nc, rows, cols, vals = read_sdm(sdm)
    got = sp.coo_matrix((vals, (rows.astype(int) - 1, cols.astype(int) - 1)), shape=(nr, nc)).tocsc()
    assert (got != M.tocsc()).nnz == 0
    # .sdm -> .mtx (%g precision, like the reference's operator<<) -> scipy
    mtx2 = str(tmp_path / "r2.mtx")
    assert run(tool, "sparse", sdm, mtx2).returncode == 0
    M2 = scipy.io.mmread(mtx2).tocsc()
    assert abs(M2 - M.tocsc()).max() <= 1e-5 * abs(M).max()


def test_dense_formats(tool, tmp_path):
    A = np.random.default_rng(3).normal(size=(7, 5))
    raw = struct.pack("<QQ", 7, 5) + A.T.astype("<f8").tobytes()
    ddm = tmp_path / "a.ddm"
    ddm.write_bytes(raw)
    for ext in ("mtx", "csv"):
        mid = str(tmp_path / ("a." + ext))
        assert run(tool, "dense", str(ddm), mid).returncode == 0
        back = str(tmp_path / ("b_%s.ddm" % ext))
        assert run(tool, "dense", mid, back).returncode == 0
        np.testing.assert_allclose(read_ddm(back), A, rtol=1e-5)
    arr = scipy.io.mmread(str(tmp_path / "a.mtx"))
    np.testing.assert_allclose(arr, A, rtol=1e-5)
    copy = str(tmp_path / "c.ddm")
    assert run(tool, "dense", str(ddm), copy).returncode == 0
    assert open(copy, "rb").read() == raw


def test_errors(tool, tmp_path):
    assert run(tool, "sparse", str(tmp_path / "missing.mtx"), str(tmp_path / "o.sdm")).returncode == 2
    bad = tmp_path / "x.foo"
    bad.write_text("1")
    assert run(tool, "sparse", str(bad), str(tmp_path / "o.sdm")).returncode == 2
    oob = tmp_path / "oob.mtx"
    oob.write_text("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n")
    assert run(tool, "sparse", str(oob), str(tmp_path / "o.sdm")).returncode == 2
    ddm = tmp_path / "d.ddm"
    ddm.write_bytes(struct.pack("<QQ", 1, 1) + struct.pack("<d", 1.0))
    assert run(tool, "sparse", str(ddm), str(tmp_path / "o.sdm")).returncode == 2   # dense file asked as sparse
