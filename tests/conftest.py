import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    class G:
        def __getattr__(self, name):
            d = np.load(os.path.join(GOLD, name + ".npz"))
            setattr(self, name, d)
            return d
    return G()


@pytest.fixture(scope="session")
def c1(golden):
    """The shipped sample operator D (4^4 x 4 x 3 hopping matrix) and k of src/main.cpp:845-847."""
    m = golden.c1_matrix
    row = m["row"].astype(np.int64)
    col = m["col"].astype(np.int64)
    val = m["val"]
    k = 0.05 + 8 * ((0.17865 - 0.05) / 10.)
    return dict(row=row, col=col, val=val, k=k, n=3072, dims=[4, 4, 4, 4, 4, 3])


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


HIST_TOL = 1e-10


ULP = 2.220446049250313e-16


def reference_envelope(solve_ref, ref_hist, it_ref, nper=3, eps=ULP):
    """How far the REFERENCE algorithm (the restatement, pinned bit-exactly to the reference) moves away from its own residual
    history under changes that leave it the same algorithm in exact arithmetic: its inner products summed right to left and in
    blocks instead of left to right (a sequential sum of n terms carries a rounding error ~sqrt(n) ulp that every parallel
    reduction replaces by another one), and its right-hand side changed by one unit in the last place (every element times
    1 +- 2^-52, random sign), `nper` times.  Restarted / truncated GCR amplifies such noise exponentially ("parity
    horizon"): on the shipped sample the history is reproducible to 1e-10 for ~40-200 iterations depending on the mode, on
    symmetric stencil operators for ~25.  This envelope is the tightest bar a parallel reduction can be held to.
    solve_ref(rng, eps) -> (hist, iters).  Returns (running-max envelope per iteration -- finite everywhere: beyond the end
    of a shorter run its last deviation is carried --, iteration count spread)."""
    from oracle import pyoracle
    env = np.zeros(len(ref_hist))
    spread = 0

    def account(h, it):
        nonlocal spread
        n = min(len(h), len(ref_hist))
        rel = np.maximum.accumulate(np.abs(h[:n] - ref_hist[:n]) / ref_hist[:n])
        env[:n] = np.maximum(env[:n], rel)
        env[n:] = np.maximum(env[n:], rel[-1])
        spread = max(spread, abs(it - it_ref))

    for mode in (1, 2):
        pyoracle.set_sum_order(mode)
        try:
            h, it = solve_ref(np.random.default_rng(0), 0.0)
        finally:
            pyoracle.set_sum_order(0)
        account(h, it)
    for s in range(nper):
        h, it = solve_ref(np.random.default_rng(100 + s), eps)
        account(h, it)
    return env, spread


def check_hist(hist, ref, it, it_ref, env=None, spread=0, tol=HIST_TOL, safety=25.0):
    """residual history within `tol` relative of the reference's, iteration count within +-1 -- relaxed only where, and
    only as far as (a finite multiple of), the reference's own history moves under another summation order / a one-ulp
    change of its input (see reference_envelope)."""
    m = min(len(hist), len(ref))
    rel = np.maximum.accumulate(np.abs(hist[:m] - ref[:m]) / ref[:m])
    bound = np.full(m, tol) if env is None else np.maximum(tol, safety * env[:m])
    bad = np.nonzero(rel > bound)[0]
    assert bad.size == 0, "residual history deviates at step %d: rel %.3e > bound %.3e" % (int(bad[0]), rel[bad[0]], bound[bad[0]])
    assert bound.size == 0 or np.all(np.isfinite(bound))
    # iteration count: +-1, or twice the spread the reference's own count shows over the handful of envelope runs (restarted GCR
    # on the 2-D Laplacian takes 619 iterations and moves by +-105 under a one-ulp change of its right-hand side)
    assert abs(it - it_ref) <= max(1, 2 * spread), (it, it_ref, spread)


def perturbed(orc, Ao, prm, rhs, x0=None, precond=None, left=None):
    def run(rng, eps):
        _, h, it = orc.gcr_solve(Ao, prm, rhs * (1 + eps * rng.choice([-1.0, 1.0], size=len(rhs))), x0=x0, precond=precond, left=left)
        return h, it
    return run


