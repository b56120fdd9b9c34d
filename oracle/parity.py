"""Parity of the BENCHMARKED solve: bench.py's workload parameterisation (MG_DEFAULT cycle parameters, scalar_levels
hierarchy shapes, outer restart / tolerance) run at a reduced lattice size on the GPU path and on the CPU restatement
(oracle/mgcr_oracle.c), with the restatement's level-0 near-null vectors handed to both sides so that the hierarchies
are the same object.  TEST INFRASTRUCTURE: used by tests/test_gpu_mg.py, tests/test_gpu_var.py and by bench.py's
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


def oracle_solve(Ao, lv, mgp, restart, max_iter, tol, rhs=None, nearnull=None):
    """the restatement's MG-GCR; returns dict(x, hist, iters, nearnull, mg, seconds per stage)"""
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
    return dict(x=x, hist=hist, iters=it, nearnull=nearnull, mg=mo, prm=prm, rhs=rhs, arnoldi_s=t1 - t0, setup_s=t2 - t1, solve_s=t3 - t2)


def oracle_envelope(Ao, ref, nper=2, eps=1e-16):
    """running-max relative deviation of the restatement's own history when its right-hand side is perturbed at the 1e-16
    level (less than one rounding of the input); finite everywhere: beyond the end of a shorter perturbed run its last
    deviation is carried.  Returns (envelope per iteration, iteration-count spread)."""
    hist = ref["hist"]
    env = np.zeros(len(hist))
    spread = 0
    for s in range(nper):
        rng = np.random.default_rng(100 + s)
        _, h, it = orc.gcr_solve(Ao, ref["prm"], ref["rhs"] * (1 + eps * rng.standard_normal(len(ref["rhs"]))), precond=ref["mg"].as_op())
        m = min(len(h), len(hist))
        rel = np.maximum.accumulate(np.abs(h[:m] - hist[:m]) / hist[:m])
        env[:m] = np.maximum(env[:m], rel)
        env[m:] = np.maximum(env[m:], rel[-1])
        spread = max(spread, abs(it - ref["iters"]))
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
    out["size"] = "x".join(map(str, dims))
    out["levels"] = len(lv) + 1
    out["how"] = ("same workload parameters (cycle, hierarchy shape, restart, tolerance) at reduced size on the GPU path and on the CPU "
                  "restatement (oracle/mgcr_oracle.c), both hierarchies built from the restatement's level-0 near-null vectors")
    return out, ref
