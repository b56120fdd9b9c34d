"""Microbenchmark of the matrix-free stencil kernels through the C ABI (no torch): DiracOp(Hopping) applies timed by the
library's pooled CUDA events; the TMA form (hopping_kernel = 2) is checked bit for bit against the register-marching form.
Knobs come from the environment (MGCR_HOP_STAGES, MGCR_HOP_ZC, MGCR_HOP_TILE); one process per knob setting."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mgpreconditionedgcr_b200 import host  # noqa: E402

args = [a for a in sys.argv[1:] if a != "--residual"]
residual = "--residual" in sys.argv          # time r = b - A x (the multigrid cycle's form) instead of y = A x
kernel = int(args[0])
lattices = [[int(v) for v in a.split("x")] for a in args[1:]] or [[256, 256, 256]]
ctx = host.Context(0)
tag = ("residual " if residual else "") + ("var " if os.environ.get("HOP_VAR") else "") + "kernel=%d stages=%s zc=%s tile=%s" % (kernel, os.environ.get("MGCR_HOP_STAGES", "-"), os.environ.get("MGCR_HOP_ZC", "-"), os.environ.get("MGCR_HOP_TILE", "-"))
for dims in lattices:
    V = int(np.prod(dims))
    x = ctx.init_rand(1, V)
    ctx.set_option("hopping_kernel", kernel)
    cls = "hopping_dirac"
    if os.environ.get("HOP_VAR"):   # variable bonds + diagonal (random values: only the traffic matters here)
        rng = np.random.default_rng(1)
        A = host.DiracOp(ctx, host.Hopping(ctx, dims, faces=[0.5 + rng.random(V) for _ in dims]), 1.0, diag=6.0 + rng.random(V))
        cls = "hopping_var_dirac"
    else:
        A = host.DiracOp(ctx, host.Hopping(ctx, dims), 1.0 / 6.01)
    y = ctx.field(V)
    b = ctx.init_rand(2, V) if residual else None
    run = (lambda out: A.residual(x, b, out=out)) if residual else (lambda out: A(x, out=out))
    for _ in range(3):
        run(y)
    ctx.sync()
    ctx.set_profile(True)
    for _ in range(20):
        run(y)
    ctx.sync()
    prof = ctx.profile()
    ctx.set_profile(False)
    p = prof[cls]
    line = "%s %s: %.1f us %.0f GB/s" % (tag, "x".join(map(str, dims)), 1e3 * p["ms"] / p["calls"], p["bytes"] / (p["ms"] * 1e-3) / 1e9)
    if kernel != 1 and V <= 2 ** 25:
        ctx.set_option("hopping_kernel", 1)
        yr = ctx.field(V)
        run(yr)
        line += " exact=%s" % bool(np.array_equal(y.numpy(), yr.numpy()))
    print(line, flush=True)
