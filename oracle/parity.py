"""Parity of the BENCHMARKED solve: bench.py's workload parameterisation (MG_DEFAULT cycle parameters, scalar_levels
hierarchy shapes, outer restart / tolerance) run at a reduced lattice size on the GPU path and on the CPU restatement
(oracle/mgcr_oracle.c), with the restatement's near-null vectors (every level) handed to both sides so that the
hierarchies are the same object.  TEST INFRASTRUCTURE: used by tests/test_gpu_mg.py, tests/test_gpu_var.py and by bench.py's
cpu_baseline leg (outside the timed region; the same oracle run is the CPU timing and the parity reference).

Reference pieces restated by what is compared: cycle structure src/MG.h:405-430, restrict / expand src/MG.h:347-383,
outer loop src/GCR.h:222-288.
"""
import time

import numpy as np

from . import pyoracle as orc


def levels(dims, subs, n_eigen):
    """MG level configs of a 3-D scalar lattice (bench.scalar_levels; duplicated here so that tests need not import bench)"""
    lv, cur, ncol = [], list(dims), 1
    for sub, ne in zip(subs, n_eigen):
        sub3 = [sub] * 3 if isinstance(sub, int) else list(sub)
        lv.append(dict(site_dims=[1] + cur, sub=[1] + sub3, n_spin=1, n_col=ncol, n_eigen=ne))
        cur = [d // q for d, q in zip(cur, sub3)]
        ncol = ne
    return lv


def operators(host, ctx, dims, m2=0.01, aniso=None):
    """(GPU operator, oracle operator) of a bench workload at lattice `dims`"""
    if aniso:
        faces, diag = host.synthetic_bonds(dims, eps=aniso["eps"], sigma=aniso["sigma"], m2=aniso["m2"], seed=aniso["seed"])
        A = host.DiracOp(ctx, host.Hopping(ctx, dims, faces=faces), 1.0, diag=diag) if ctx is not None else None
        return A, orc.dirac(orc.hopping(dims, faces), 1.0, diag)
    k = 1.0 / (2 * len(dims) + m2)
    A = host.DiracOp(ctx, host.Hopping(ctx, dims), k) if ctx is not None else None
    return A, orc.dirac(orc.hopping(dims), k)


ULP = 2.220446049250313e-16


def oracle_solve(Ao, lv, mgp, restart, max_iter, tol, rhs=None, nearnull=None):
    """the restatement's MG-GCR; nearnull: None (inverse iteration on every level), the level-0 vectors, or a per-level list.
    Returns dict(x, hist, iters, nearnull = the per-level vectors the hierarchy was built from, mg, seconds per stage)"""
    eig, coarse, smooth = (orc.gcr_param(*mgp[k]) for k in ("eigen", "coarse", "smooth"))
    t0 = time.perf_counter()
    if nearnull is None:
        nearnull = orc.arnoldi(Ao, eig, lv[0]["n_eigen"])
    t1 = time.perf_counter()
    mo = orc.MG(Ao, lv, eig, coarse, smooth, nearnull=nearnull)
    t2 = time.perf_counter()
    if rhs is None:
        rhs = orc.init_rand(0, Ao.n)
    prm = orc.gcr_param(0, restart, max_iter, tol)
    x, hist, it = orc.gcr_solve(Ao, prm, rhs, precond=mo.as_op())
    t3 = time.perf_counter()
    return dict(x=x, hist=hist, iters=it, nearnull=[mo.nearnull(l) for l in range(len(lv))], mg=mo, prm=prm, rhs=rhs, lv=lv, mgp=mgp,
                arnoldi_s=t1 - t0, setup_s=t2 - t1, solve_s=t3 - t2)


def oracle_envelope(Ao, ref, nper=2, inputs=True):
    """How far the RESTATEMENT moves away from its own residual history under changes that leave the algorithm the same in
    exact arithmetic -- the tightest bar an implementation with parallel reductions can be held to:
      * summation order: the reference adds the 10^6..10^9 terms of every inner product left to right (src/Fields.h:216-235),
        a rounding error of ~1e-13 per inner product that any tree reduction replaces by a different (smaller) one; the
        solve is repeated with the sums taken right to left and in blocks (the hierarchy rebuilt the same way);
      * one unit in the last place: every element of the right-hand side and (inputs=True) of the near-null vectors of every
        level multiplied by 1 +- 2^-52 with a random sign, `nper` times.
    What it shows on the benchmarked solves: the solutions stay together to 4e-16 and the iteration counts do not move, but
    residual norms 1e-6 below the start differ by 1e-10 .. 1e-9 relative and come back together as the solver reduces the
    difference.  Finite everywhere: beyond the end of a shorter run its last deviation is carried.
    Returns (running-max envelope per iteration, iteration-count spread)."""
    hist = ref["hist"]
    env = np.zeros(len(hist))
    spread = 0
    eig, coarse, smooth = (orc.gcr_param(*ref["mgp"][k]) for k in ("eigen", "coarse", "smooth"))

    def account(h, it):
        nonlocal spread
        m = min(len(h), len(hist))
        rel = np.maximum.accumulate(np.abs(h[:m] - hist[:m]) / hist[:m])
        env[:m] = np.maximum(env[:m], rel)
        env[m:] = np.maximum(env[m:], rel[-1])
        spread = max(spread, abs(it - ref["iters"]))

    for mode in (1, 2):
        orc.set_sum_order(mode)
        try:
            mo = orc.MG(Ao, ref["lv"], eig, coarse, smooth, nearnull=ref["nearnull"])
            _, h, it = orc.gcr_solve(Ao, ref["prm"], ref["rhs"], precond=mo.as_op())
        finally:
            orc.set_sum_order(0)
        account(h, it)
    for s in range(nper):
        rng = np.random.default_rng(100 + s)
        rhs = ref["rhs"] * (1 + ULP * rng.choice([-1.0, 1.0], size=len(ref["rhs"])))
        mo = ref["mg"]
        if inputs:
            mo = orc.MG(Ao, ref["lv"], eig, coarse, smooth, nearnull=[v * (1 + ULP * rng.choice([-1.0, 1.0], size=v.shape)) for v in ref["nearnull"]])
        _, h, it = orc.gcr_solve(Ao, ref["prm"], rhs, precond=mo.as_op())
        account(h, it)
    return env, spread


def gpu_solve(host, ctx, A, lv, mgp, restart, max_iter, tol, rhs, nearnull=None):
    mg = host.MG(ctx, A, lv, host.GCR_Param(*mgp["eigen"]), host.GCR_Param(*mgp["coarse"]), host.GCR_Param(*mgp["smooth"]), nearnull=nearnull)
    x = ctx.field(A.get_dim()).set_zero()
    it, hist = host.GCR(ctx, A, host.GCR_Param(0, restart, max_iter, tol, False, None, mg)).solve(ctx.from_numpy(rhs), x)
    out = dict(x=x.numpy(), hist=hist, iters=it)
    mg.destroy()
    return out


def compare(gpu, ref):
    """north-star quantities: per-iteration residual norms (max relative deviation over the common part of the two
    histories, and the iteration up to which it stays below 1e-10), iteration counts, relative solution difference"""
    m = min(len(gpu["hist"]), len(ref["hist"]))
    rel = np.abs(gpu["hist"][:m] - ref["hist"][:m]) / ref["hist"][:m]
    below = np.nonzero(np.maximum.accumulate(rel) > 1e-10)[0]
    return dict(iters_gpu=int(gpu["iters"]), iters_oracle=int(ref["iters"]), max_hist_rel=float(rel.max()),
                hist_within_1e10_until=int(below[0]) if below.size else int(m),
                x_rel=float(np.linalg.norm(gpu["x"] - ref["x"]) / np.linalg.norm(ref["x"])),
                final_gpu=float(gpu["hist"][-1]), final_oracle=float(ref["hist"][-1]))


def bench_parity(host, ctx, wl, dims, subs, n_eigen):
    """bench.py's workload `wl` at lattice `dims` with hierarchy (subs, n_eigen): GPU vs restatement on the restatement's
    near-null vectors.  Returns (parity dict for the bench line, oracle result with its timings)."""
    m = wl["mg"]
    lv = levels(dims, subs, n_eigen)
    A, Ao = operators(host, ctx, dims, m2=wl.get("m2", 0.01), aniso=wl.get("aniso"))
    ref = oracle_solve(Ao, lv, m, wl["restart"], wl["max_iter"], wl["tol"])
    gpu = gpu_solve(host, ctx, A, lv, m, wl["restart"], wl["max_iter"], wl["tol"], ref["rhs"], nearnull=ref["nearnull"])
    out = compare(gpu, ref)
    env, _ = oracle_envelope(Ao, ref, nper=1)
    out["oracle_own_envelope_max"] = float(env.max())
    out["size"] = "x".join(map(str, dims))
    out["levels"] = len(lv) + 1
    out["how"] = ("same workload parameters (cycle, hierarchy shape, restart, tolerance) at reduced size on the GPU path and on the CPU "
                  "restatement (oracle/mgcr_oracle.c), both hierarchies built from the restatement's near-null vectors on every level; "
                  "oracle_own_envelope_max = how far the restatement's own history moves when its inner products are summed in another order "
                  "or its inputs change by one unit in the last place")
    return out, ref
