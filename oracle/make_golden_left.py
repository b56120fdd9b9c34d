#!/usr/bin/env python
"""Golden fixture of the reference's LEFT-preconditioned GCR (src/GCR.h:201-204, 245-247), from the UNMODIFIED reference
(oracle/_ref/ref_oracle gcr-left): the shipped 4^4 sample operator as DiracOp(k = 0.15292), a Jacobi-like real diagonal as the
left preconditioner (a caller-defined subclass of the reference's Operator interface in oracle/ref_harness.cpp), rhs =
init_rand(0), x0 = 0, restart-5 and restart-10 modes, max_iter 4000, tol 1e-12.  Writes tests/golden/gcr_left.npz.

    python oracle/make_golden_left.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def left_diagonal(n):
    """the diagonal of L: 1 / (1 + 0.5 u), u uniform in [0, 1) from a fixed seed -- also used by the tests"""
    return 1.0 / (1.0 + 0.5 * np.random.default_rng(2024).random(n))


def main():
    from oracle import pyoracle as orc
    subprocess.check_call(["make", "-C", HERE, "_ref/ref_oracle"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    m = np.load(os.path.join(GOLD, "c1_matrix.npz"))
    n = 3072
    d = tempfile.mkdtemp(prefix="left_ref_")
    m["row"].astype(np.int64).tofile(os.path.join(d, "row.bin"))
    m["col"].astype(np.int64).tofile(os.path.join(d, "col.bin"))
    m["val"].astype(np.complex128).tofile(os.path.join(d, "val.bin"))
    orc.init_rand(0, n).tofile(os.path.join(d, "rhs.bin"))
    np.zeros(n, dtype=np.complex128).tofile(os.path.join(d, "x0.bin"))
    left_diagonal(n).tofile(os.path.join(d, "diag.bin"))
    k = 0.05 + 8 * ((0.17865 - 0.05) / 10.)
    out = {}
    # (truncation modes are left out: with this preconditioner they do not converge, and at max_iter the reference overruns its history, Q11)
    for tag, (tr, re) in (("r5", (0, 5)), ("r10", (0, 10))):
        js = subprocess.run([os.path.join(HERE, "_ref", "ref_oracle"), "gcr-left", d, str(n), str(m["val"].size), repr(k), "0", str(tr), str(re), "4000", "1e-12"],
                            check=True, capture_output=True, text=True).stdout
        js = json.loads([l for l in js.splitlines() if l.startswith("{")][-1])
        out[tag + "_hist"] = np.fromfile(os.path.join(d, "hist.bin"), dtype=np.float64)
        out[tag + "_x"] = np.fromfile(os.path.join(d, "x.bin"), dtype=np.complex128)
        print("golden left-preconditioned gcr %s: iters=%d final=%.10e" % (tag, js["iters"], js["final_rel_res"]))
    np.savez_compressed(os.path.join(GOLD, "gcr_left.npz"), **out)


if __name__ == "__main__":
    main()
