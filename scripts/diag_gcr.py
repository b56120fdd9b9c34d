"""diagnostic (not a test): GPU GCR history vs reference golden, per mode -- where does the deviation start?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mgpreconditionedgcr_b200 import host
from oracle import pyoracle as orc
g = np.load(os.path.join(ROOT, "tests/golden/gcr.npz"))
m = np.load(os.path.join(ROOT, "tests/golden/c1_matrix.npz"))
row, col, val = m["row"].astype(np.int64), m["col"].astype(np.int64), m["val"]
k = 0.05 + 8 * ((0.17865 - 0.05) / 10.)
ctx = host.Context(0)
A = host.DiracOp(ctx, host.Sparse(ctx, 3072, 3072, row, col, val), k)
Ao = orc.dirac(orc.csr(3072, 3072, row, col, val), k)
MODES = {"r5": (0, 5, 4000, 1e-13), "r2": (0, 2, 4000, 1e-13), "t5": (5, 0, 4000, 1e-13), "r10": (0, 10, 4000, 1e-10), "full100": (0, 0, 100, 1e-10), "smooth0": (0, 10, 0, 1e-8)}
rhs0 = orc.init_rand(0, 3072)
rng = np.random.default_rng(1)
for mode, (t, r, mi, tol) in MODES.items():
    x = ctx.field(3072).set_zero()
    it, hist = host.GCR(ctx, A, host.GCR_Param(t, r, mi, tol, False, None, None)).solve(ctx.from_numpy(rhs0), x)
    ref = g[mode + "_hist"]
    n = min(len(hist), len(ref))
    rel = np.abs(hist[:n] - ref[:n]) / ref[:n]
    first = int(np.argmax(rel > 1e-10)) if (rel > 1e-10).any() else -1
    xe = np.linalg.norm(x.numpy() - g[mode + "_x"]) / np.linalg.norm(g[mode + "_x"])
    # the oracle itself under a 1e-16 perturbation of rhs
    xo, ho, ito = orc.gcr_solve(Ao, orc.gcr_param(t, r, mi, tol), rhs0 * (1 + 1e-16 * rng.standard_normal(3072)))
    n2 = min(len(ho), len(ref))
    relo = np.abs(ho[:n2] - ref[:n2]) / ref[:n2]
    firsto = int(np.argmax(relo > 1e-10)) if (relo > 1e-10).any() else -1
    print("%-8s gpu it=%d ref it=%d first>1e-10 @%d (ref there %.2e) maxrel %.2e xerr %.2e | perturbed-oracle it=%d first @%d maxrel %.2e" % (
        mode, it, len(ref) - 1, first, ref[first] if first >= 0 else 0, rel.max(), xe, ito, firsto, relo.max()))
    print("   rel at 10,20,40,80:", [float("%.2e" % rel[i]) for i in (10, 20, 40, 80) if i < n])
