"""One rank of the multi-GPU parity check (spawned by tests/test_gpu_dist.py, or run under torchrun):
slab-partitioned stencil operator, GCR and the distributed multigrid hierarchy against the same problem solved on one
GPU by the same library.  Prints one JSON line per check on rank 0."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(rank, world, port, gather_dofs, mode="unit"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from mgpreconditionedgcr_b200 import host

    torch.cuda.set_device(rank)
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    ctx = host.Context(rank)
    ids = [host.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.init_dist(rank, world, ids[0])
    ctx.set_slab_align(16)
    if gather_dofs is not None:
        ctx.set_option("gather_dofs", gather_dofs)
    single = host.Context(rank)          # the same problem on one GPU, for comparison

    dims = [32 * world, 16, 24]
    V = int(np.prod(dims))
    plane = dims[1] * dims[2]
    k = 1. / 6.01
    b, e = host.slab_range(dims[0], 16, rank, world)
    sl = slice(b * plane, e * plane)
    out = {}

    def make(c, m2, z_range=None):
        """the operator of the run: I - k H (unit hopping), or the anisotropic variable-coefficient diag - H of configs[4]"""
        if mode == "var":
            faces, diag = host.synthetic_bonds(dims, m2=m2, z_range=z_range)
            return host.DiracOp(c, host.Hopping(c, dims, faces=faces), 1.0, diag=diag)
        return host.DiracOp(c, host.Hopping(c, dims), k if m2 == 0.01 else 0.12)

    A = make(ctx, 0.01, (b, e))
    A1 = make(single, 0.01)
    assert A.get_dim() == (e - b) * plane and A.global_dim() == V
    # operator apply with halo exchange
    f = single.init_rand(1, V)
    fl = ctx.init_rand(1, A.get_dim(), skip=b * plane)
    assert np.array_equal(fl.numpy(), f.numpy()[sl])
    out["apply_exact"] = bool(np.array_equal(A(fl).numpy(), A1(f).numpy()[sl]))
    # the TMA-staged stencil form with the neighbour planes coming from the halo buffers (lattice wide enough for its tiles)
    dims_w = [16 * world, 24, 72]
    pw = dims_w[1] * dims_w[2]
    for c in (ctx, single):
        c.set_option("hopping_tma_rows", 0)
    if mode == "var":
        fw, dw = host.synthetic_bonds(dims_w, m2=0.3)
        Aw_t = host.DiracOp(ctx, host.Hopping(ctx, dims_w, faces=[q[b // 2:e // 2] for q in fw]), 1.0, diag=dw[b // 2:e // 2])
        Aw_1 = host.DiracOp(single, host.Hopping(single, dims_w, faces=fw), 1.0, diag=dw)
    else:
        Aw_t = host.DiracOp(ctx, host.Hopping(ctx, dims_w), k)
        Aw_1 = host.DiracOp(single, host.Hopping(single, dims_w), k)
    fw1 = single.init_rand(7, int(np.prod(dims_w)))
    fwl = ctx.init_rand(7, Aw_t.get_dim(), skip=(b // 2) * pw)
    out["apply_tma_exact"] = bool(np.array_equal(Aw_t(fwl).numpy(), Aw_1(fw1).numpy()[(b // 2) * pw:(e // 2) * pw]))
    for c in (ctx, single):
        c.set_option("hopping_tma_rows", 1 << 19)
    # row-slab-partitioned CSR (general Sparse: gather list of off-slab columns, packed and exchanged per apply): the same
    # hopping matrix handed over as this rank's rows with global columns, and a random complex matrix whose rows reach into
    # every other rank's slab
    if mode == "unit":
        row, col, val = host.hopping_csr(dims)
        r0, r1 = b * plane, e * plane
        loc = slice(row[r0], row[r1])
        S = host.Sparse(ctx, r1 - r0, V, row[r0:r1 + 1] - row[r0], col[loc], val[loc], row_range=(r0, r1), nrow_global=V)
        Ad = host.DiracOp(ctx, S, k)
        out["csr_apply_exact"] = bool(np.array_equal(Ad(fl).numpy(), A1(f).numpy()[sl]))
        rng = np.random.default_rng(17)
        nn = 4000 * world
        rr, cc = [0], []
        for i in range(nn):
            c = np.unique(np.concatenate([rng.integers(0, nn, size=5), [i]]))
            cc += list(c)
            rr.append(len(cc))
        rr, cc = np.array(rr, np.int64), np.array(cc, np.int64)
        vv = rng.standard_normal(len(cc)) + 1j * rng.standard_normal(len(cc))
        q0, q1 = 4000 * rank, 4000 * (rank + 1)
        Sr = host.Sparse(ctx, q1 - q0, nn, rr[q0:q1 + 1] - rr[q0], cc[rr[q0]:rr[q1]], vv[rr[q0]:rr[q1]], row_range=(q0, q1), nrow_global=nn)
        S1 = host.Sparse(single, nn, nn, rr, cc, vv)
        xr = rng.standard_normal(nn) + 1j * rng.standard_normal(nn)
        out["csr_random_exact"] = bool(np.array_equal(Sr(xr[q0:q1]), S1(xr)[q0:q1]))
        # GCR on the distributed stored operator against the matrix-free one
        xs = ctx.field(A.get_dim()).set_zero()
        its, hs = host.GCR(ctx, host.DiracOp(ctx, S, 0.12), host.GCR_Param(0, 5, 2000, 1e-10, False, None, None)).solve(ctx.init_rand(0, A.get_dim(), skip=b * plane), xs)
        out["csr_gcr"] = [its, float(hs[-1])]
    # all-reduced inner products
    g = single.init_rand(3, V)
    gl = ctx.init_rand(3, A.get_dim(), skip=b * plane)
    out["dot_rel"] = abs(fl.dot(gl) - f.dot(g)) / abs(f.dot(g))
    # unpreconditioned GCR on a well-conditioned shift of the operator (converges inside the parity horizon of restarted
    # GCR, tests/test_parity_horizon.py), then on the ill-conditioned one for the iteration count MG has to beat
    p = host.GCR_Param(0, 5, 2000, 1e-10, False, None, None)
    rhs = single.init_rand(0, V)
    rhsl = ctx.init_rand(0, A.get_dim(), skip=b * plane)
    x1 = single.field(V).set_zero()
    it1, _ = host.GCR(single, A1, p).solve(rhs, x1)
    Aw = make(ctx, 0.5, (b, e))
    Aw1 = make(single, 0.5)
    x1.set_zero()
    itw1, h1 = host.GCR(single, Aw1, p).solve(rhs, x1)
    xl = ctx.field(A.get_dim()).set_zero()
    itd, hd = host.GCR(ctx, Aw, p).solve(rhsl, xl)
    m = min(len(h1), len(hd), 25)
    out["gcr_iters"] = [itw1, itd]
    out["gcr_hist_rel"] = float(np.max(np.abs(hd[:m] - h1[:m]) / h1[:m]))
    out["gcr_x_rel"] = float(np.linalg.norm(xl.numpy() - x1.numpy()[sl]) / np.linalg.norm(x1.numpy()[sl]))
    # multigrid: two coarse grids, 4^3 aggregates
    if mode == "var":   # line aggregates along the strongly coupled (fastest) direction first, as in bench.py's mg3d_aniso
        lv = [dict(site_dims=[1] + dims, sub=[1, 1, 1, 8], n_spin=1, n_col=1, n_eigen=2),
              dict(site_dims=[1, dims[0], dims[1], dims[2] // 8], sub=[1, 4, 2, 3], n_spin=1, n_col=2, n_eigen=4),
              dict(site_dims=[1, dims[0] // 4, dims[1] // 2, 1], sub=[1, 4, 4, 1], n_spin=1, n_col=4, n_eigen=4)]
    else:
        lv = [dict(site_dims=[1] + dims, sub=[1, 4, 4, 4], n_spin=1, n_col=1, n_eigen=4),
              dict(site_dims=[1] + [d // 4 for d in dims], sub=[1, 4, 2, 2], n_spin=1, n_col=4, n_eigen=4)]
    eig, coarse, smooth = host.GCR_Param(0, 10, 10, 1e-8), host.GCR_Param(0, 10, 8, 1e-2), host.GCR_Param(0, 4, 3, 1e-8)
    mg = host.MG(ctx, A, lv, eig, coarse, smooth)
    mg1 = host.MG(single, A1, lv, eig, coarse, smooth)
    i0, i1 = mg.info(0), mg1.info(0)
    out["nblocks"] = [i0["n_blocks"], i1["n_blocks"]]
    # Galerkin identity on the distributed objects: Ac xc == R A P xc (ghost prolongator rows + coarse halo exchange)
    nc = i0["n_blocks"] * i0["ne"]
    xc = ctx.init_rand(5, nc, skip=rank * nc)
    lhs = mg.coarse_op(0)(xc).numpy()
    rhs_c = mg.restrict(A(mg.expand(xc))).numpy()
    out["galerkin_rel"] = float(np.linalg.norm(lhs - rhs_c) / np.linalg.norm(rhs_c))
    # restrict / prolong are local: R P = 1
    out["rp_identity"] = float(np.linalg.norm(mg.restrict(mg.expand(xc)).numpy() - xc.numpy()) / np.linalg.norm(xc.numpy()))
    # MG-preconditioned GCR, distributed vs one GPU
    pm = host.GCR_Param(0, 10, 200, 1e-10, False, None, mg)
    pm1 = host.GCR_Param(0, 10, 200, 1e-10, False, None, mg1)
    y1 = single.field(V).set_zero()
    itm1, hm1 = host.GCR(single, A1, pm1).solve(rhs, y1)
    yl = ctx.field(A.get_dim()).set_zero()
    itmd, hmd = host.GCR(ctx, A, pm).solve(rhsl, yl)
    r = rhsl - A(yl)
    out["mg_iters"] = [itm1, itmd, it1]
    out["mg_true_res"] = r.norm() / rhsl.norm()
    out["mg_x_rel"] = float(np.linalg.norm(yl.numpy() - y1.numpy()[sl]) / np.linalg.norm(y1.numpy()[sl]))
    out["mg_hist_rel"] = float(np.max(np.abs(hmd[:min(len(hmd), len(hm1), 12)] - hm1[:min(len(hmd), len(hm1), 12)]) / hm1[:min(len(hmd), len(hm1), 12)]))
    # LOOSE inner tolerances: the smoother / coarse solves stop on their device-side test after different numbers of iterations;
    # every rank must take the same decision from the all-reduced norms (a partial ||r||^2 in the test once made ranks disagree)
    loose_c, loose_s = host.GCR_Param(0, 10, 4, 3e-1), host.GCR_Param(0, 4, 4, 3e-1)
    mgl = host.MG(ctx, A, lv, eig, loose_c, loose_s)
    mgl1 = host.MG(single, A1, lv, eig, loose_c, loose_s)
    y1.set_zero(); yl.set_zero()
    itl1, hl1 = host.GCR(single, A1, host.GCR_Param(0, 10, 200, 1e-10, False, None, mgl1)).solve(rhs, y1)
    itld, hld = host.GCR(ctx, A, host.GCR_Param(0, 10, 200, 1e-10, False, None, mgl)).solve(rhsl, yl)
    out["mg_loose_iters"] = [itl1, itld]
    out["mg_loose_x_rel"] = float(np.linalg.norm(yl.numpy() - y1.numpy()[sl]) / np.linalg.norm(y1.numpy()[sl]))
    rl = rhsl - A(yl)
    out["mg_loose_true_res"] = rl.norm() / rhsl.norm()
    # GPU-count-independent reduction shape (csrc/common.cuh, RedGeom): on a lattice whose length divides into the 8 virtual slabs
    # (>= 2^16 elements each) the distributed solve performs the same additions as the one-GPU solve -- histories and solutions are
    # IDENTICAL, bit for bit, for unpreconditioned GCR and for the MG-preconditioned solve (inverse iteration, hierarchy, cycle)
    if mode == "unit":
        dims_b = [max(64, 16 * world), 64, 256]   # >= 2^20 sites: above the persistent small-solve kernel's range (2^19), 8 virtual slabs of >= 2^17
        Vb = int(np.prod(dims_b))
        pb = dims_b[1] * dims_b[2]
        bb, eb = host.slab_range(dims_b[0], 16, rank, world)
        Ab, Ab1 = host.DiracOp(ctx, host.Hopping(ctx, dims_b), 1. / 6.3), host.DiracOp(single, host.Hopping(single, dims_b), 1. / 6.3)
        rb1 = single.init_rand(0, Vb)
        rbl = ctx.init_rand(0, Ab.get_dim(), skip=bb * pb)
        pg = host.GCR_Param(0, 5, 60, 1e-10, False, None, None)
        xb1, xbl = single.field(Vb).set_zero(), ctx.field(Ab.get_dim()).set_zero()
        ib1, hb1 = host.GCR(single, Ab1, pg).solve(rb1, xb1)
        ibl, hbl = host.GCR(ctx, Ab, pg).solve(rbl, xbl)
        out["bits_gcr"] = bool(ib1 == ibl and np.array_equal(hb1, hbl) and np.array_equal(xbl.numpy(), xb1.numpy()[bb * pb:eb * pb]))
        lvb = [dict(site_dims=[1] + dims_b, sub=[1, 4, 4, 4], n_spin=1, n_col=1, n_eigen=4)]
        mgb = host.MG(ctx, Ab, lvb, eig, host.GCR_Param(0, 10, 2, 1e-2), host.GCR_Param(0, 4, 2, 1e-8))
        mgb1 = host.MG(single, Ab1, lvb, eig, host.GCR_Param(0, 10, 2, 1e-2), host.GCR_Param(0, 4, 2, 1e-8))
        xb1.set_zero(); xbl.set_zero()
        jb1, gb1 = host.GCR(single, Ab1, host.GCR_Param(0, 3, 100, 1e-10, False, None, mgb1)).solve(rb1, xb1)
        jbl, gbl = host.GCR(ctx, Ab, host.GCR_Param(0, 3, 100, 1e-10, False, None, mgb)).solve(rbl, xbl)
        out["bits_mg"] = bool(jb1 == jbl and np.array_equal(gb1, gbl) and np.array_equal(xbl.numpy(), xb1.numpy()[bb * pb:eb * pb]))
        out["bits_mg_hist_rel"] = float(np.max(np.abs(gbl[:min(len(gbl), len(gb1))] - gb1[:min(len(gbl), len(gb1))]) / gb1[:min(len(gbl), len(gb1))]))
    res = [None] * world
    dist.all_gather_object(res, out)
    if rank == 0:
        print("DIST_RESULT " + json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), None if sys.argv[4] == "default" else int(sys.argv[4]),
         sys.argv[5] if len(sys.argv) > 5 else "unit")
