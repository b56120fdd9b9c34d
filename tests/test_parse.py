"""The reference's I/O pair through the drop-in headers (include/mgcr/Parse.h): parse_data (MatrixMarket -> CRS text,
src/Parse.cpp:10-62) and read_data (CRS text -> Sparse<long>, :65-91), against the UNMODIFIED reference's output on the same
file (tests/golden/parse.npz, oracle/make_golden_parse.py).  Host only: no GPU is needed (the Sparse keeps its arrays on the
host until its first apply)."""
import os
import subprocess

import numpy as np

from conftest import ROOT

BIN = os.path.join(ROOT, "examples", "_build", "parse_check")


def test_parse_data_and_read_data_against_reference(tmp_path, golden):
    g = golden.parse
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "_build/parse_check"])
    d = str(tmp_path)
    with open(os.path.join(d, "in.mtx"), "wb") as f:
        f.write(g["mtx"].tobytes())
    p = subprocess.run([BIN, os.path.join(d, "in.mtx")], env=dict(os.environ, MGCR_DATA_DIR=d), capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stdout + p.stderr
    # parse_data: the CRS text, byte for byte what the reference writes (header, row offsets on one line, `col (re,im)` lines)
    assert open(os.path.join(d, "parsed.txt"), "rb").read() == g["parsed"].tobytes()
    # read_data: the arrays the reference reads back from it (ROW[nrow] = nnz is set by the constructor, src/Operator.h:61)
    assert np.array_equal(np.fromfile(os.path.join(d, "row.bin"), dtype=np.int64), g["row"])
    assert np.array_equal(np.fromfile(os.path.join(d, "col.bin"), dtype=np.int64), g["col"])
    assert np.array_equal(np.fromfile(os.path.join(d, "val.bin"), dtype=np.complex128), g["val"])
    n = len(g["row"]) - 1
    assert p.stdout.strip().splitlines()[-1] == "PARSED %d %d %d" % (n, n, len(g["col"]))
    # duplicates were summed and the entries are sorted row-major with ascending columns
    assert len(g["col"]) < 48 * 5 + 30
    for r in range(n):
        c = g["col"][g["row"][r]:g["row"][r + 1]]
        assert np.all(np.diff(c) > 0)
