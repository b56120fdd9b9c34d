"""The faster kernel forms chosen during round 2 compute the SAME BITS as the plain forms they replaced.

Every form is selected once per process from an environment knob, so each setting runs in its own
python process: a 64^3, 3-level MG-preconditioned GCR solve with bench.py's cycle parameters, every level on the host-driven
solver path (`small_gcr_rows = 0`, so the level-0 smoothers go through csrc/gcr.cu's blind solves like the 512^3 ones do), plus
one restrict and one prolong of a random vector; the process prints SHA-256 digests of the residual history, the solution
and the transfer results.  All settings must print the digests of the default one.

  MGCR_BLIND_LEAN=0        last iteration of a blind solve as a full x / r update, A p always stored, r = rhs and p0 = rhs copied
  MGCR_RESTRICT_FOLD=0     one shuffle tree per value instead of the halving butterfly (k_restrict_warp / k_restrict_warp4)
  MGCR_PROLONG_ROWS=0      one aggregate per warp (k_prolong) instead of two (k_prolong_rows<4,2,4>)
  MGCR_PROLONG_ROWS=2      two aggregates per warp with the number of near-null vectors at run time (k_prolong_rows<4,2,0>)
  MGCR_PROLONG_ONESHOT=0 / MGCR_RESTRICT_ONESHOT=1    persistent / one-shot grids
The lean form of the blind solves is also compared with the full one for other smoother / coarse-solver shapes (1, 3 and 4
iterations, restart shorter than the solve, truncation).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %(root)r)
from mgpreconditionedgcr_b200 import host
ctx = host.Context(0)
ctx.set_option("small_gcr_rows", 0)
dims = [64, 64, 64]
k = 1.0 / (6 + 0.01)
A = host.DiracOp(ctx, host.Hopping(ctx, dims), k)
lv, cur, ncol = [], list(dims), 1
for sub, ne in ((4, 4), (4, 4)):
    lv.append(dict(site_dims=[1] + cur, sub=[1, sub, sub, sub], n_spin=1, n_col=ncol, n_eigen=ne))
    cur = [d // sub for d in cur]
    ncol = ne
import os
smooth = eval(os.environ.get("VARIANT_SMOOTH", "(0, 4, 2, 1e-8)"))
coarse = eval(os.environ.get("VARIANT_COARSE", "(0, 10, 2, 1e-2)"))
mg = host.MG(ctx, A, lv, host.GCR_Param(0, 10, 10, 1e-8), host.GCR_Param(*coarse), host.GCR_Param(*smooth))
rhs = ctx.init_rand(0, A.get_dim())
x = ctx.field(A.get_dim()).set_zero()
it, hist = host.GCR(ctx, A, host.GCR_Param(0, 3, 1000, 1e-10, False, None, mg)).solve(rhs, x)
rng = np.random.default_rng(5)
v = rng.standard_normal(A.get_dim()) + 1j * rng.standard_normal(A.get_dim())
xc = mg.restrict(ctx.from_numpy(v), 0)
xf = mg.expand(xc, 0)
h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:24]
print("DIGEST it=%%d hist=%%s x=%%s restrict=%%s expand=%%s" %% (it, h(np.asarray(hist, dtype=np.float64)), h(x.numpy()), h(xc.numpy()), h(xf.numpy())))
"""

VARIANTS = [
    {},
    {"MGCR_BLIND_LEAN": "0"},
    {"MGCR_RESTRICT_FOLD": "0"},
    {"MGCR_PROLONG_ROWS": "0"},
    {"MGCR_PROLONG_ROWS": "2", "MGCR_PROLONG_ONESHOT": "0", "MGCR_RESTRICT_ONESHOT": "1"},
]


def run_variant(env_extra):
    env = dict(os.environ)
    for k in ("MGCR_BLIND_LEAN", "MGCR_RESTRICT_FOLD", "MGCR_PROLONG_ROWS", "MGCR_PROLONG_ONESHOT", "MGCR_RESTRICT_ONESHOT"):
        env.pop(k, None)
    env.update(env_extra)
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("DIGEST")]
    assert len(lines) == 1, out.stdout[-2000:]
    return lines[0]


# (smoother, coarse solver) parameter sets = (truncation, restart, max_iter, tol): the lean form has a case per shape of a blind solve
# -- one, two, three or four iterations; restart shorter than the solve (ring slot 0 is written again: p0 must be a copy);
# truncation (the ring wraps)
BLIND_SHAPES = [
    ("(0, 4, 3, 1e-8)", "(0, 10, 3, 1e-2)"),
    ("(0, 3, 4, 1e-8)", "(0, 2, 4, 1e-2)"),
    ("(2, 0, 3, 1e-8)", "(2, 0, 4, 1e-2)"),
    ("(0, 4, 1, 1e-8)", "(0, 10, 1, 1e-2)"),
]


@pytest.mark.parametrize("smooth,coarse", BLIND_SHAPES)
def test_lean_blind_solves_for_every_solve_shape(smooth, coarse):
    shape = {"VARIANT_SMOOTH": smooth, "VARIANT_COARSE": coarse}
    full = run_variant(dict(shape, MGCR_BLIND_LEAN="0"))
    lean = run_variant(shape)
    assert lean == full, "%s: lean %s / full %s" % (shape, lean, full)


def test_kernel_forms_compute_identical_bits():
    ref = run_variant(VARIANTS[0])
    assert " it=" in ref
    for v in VARIANTS[1:]:
        got = run_variant(v)
        assert got == ref, "%s:\n  %s\nthe default forms:\n  %s" % (v, got, ref)
