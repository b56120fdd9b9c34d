"""GPU parity tests of the variable-coefficient matrix-free operator (BASELINE.json configs[4]; run with
-m gpu): y = diag.x - H x with real symmetric bond coefficients, against the oracle's CSR restatement of the Sparse the
reference would hold for the same entries (src/Operator.h:64, 330-346) -- bit-exact for the apply, residual histories of
GCR within 1e-10, the multigrid hierarchy built on it against the oracle's."""
import numpy as np
import pytest

from conftest import check_hist, perturbed, reference_envelope, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from mgpreconditionedgcr_b200 import host
    c = host.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="module")
def host():
    from mgpreconditionedgcr_b200 import host
    return host


def random_faces(dims, seed):
    rng = np.random.default_rng(seed)
    return [0.25 + rng.random(int(np.prod(dims))) for _ in dims]


@pytest.mark.parametrize("dims", [[10, 12, 14], [3, 5, 70], [9, 17], [33], [1, 1, 40], [6, 1, 9]])
def test_variable_bond_apply_is_bit_exact(ctx, host, orc, dims):
    n = int(np.prod(dims))
    faces = random_faces(dims, 7)
    x = orc.init_rand(3, n)
    H = host.Hopping(ctx, dims, faces=faces)
    Ho = orc.hopping(dims, faces)
    ref = Ho(x)
    assert np.array_equal(H(x), ref)
    # the same entries handed over as the CSR a user of the reference would build
    row, col, val = host.hopping_csr(dims, faces)
    ro, co, vo = orc.csr_export(Ho)
    assert np.array_equal(row, ro) and np.array_equal(col, co) and np.array_equal(val, vo)
    assert np.array_equal(host.Sparse(ctx, n, n, row, col, val)(x), ref)
    # DiracOp with a real diagonal, real and complex k
    diag = 1.0 + np.random.default_rng(8).random(n)
    for k in (1.0, 0.11 + 0.07j):
        assert np.array_equal(host.DiracOp(ctx, H, k, diag=diag)(x), orc.dirac(Ho, k, diag)(x))
    assert np.array_equal(host.DiracOp(ctx, H, 0.3)(x), orc.dirac(Ho, 0.3)(x))
    assert H.apply_bytes() == (32 + 8 * len(dims)) * n


def test_variable_bonds_from_device_memory(ctx, host, orc):
    torch = pytest.importorskip("torch")
    dims = [12, 10, 18]
    n = int(np.prod(dims))
    faces, diag = host.synthetic_bonds(dims)
    tf = [torch.from_numpy(f.reshape(-1)).cuda() for f in faces]
    td = torch.from_numpy(diag.reshape(-1)).cuda()
    torch.cuda.synchronize()
    A = host.DiracOp(ctx, host.Hopping(ctx, dims, faces_dev=[t.data_ptr() for t in tf]), 1.0, diag_dev=td.data_ptr())
    x = orc.init_rand(5, n)
    assert np.array_equal(A(x), orc.dirac(orc.hopping(dims, faces), 1.0, diag)(x))


def test_gcr_on_the_anisotropic_operator(ctx, host, orc):
    dims = [12, 16, 20]
    n = int(np.prod(dims))
    faces, diag = host.synthetic_bonds(dims, m2=0.5)
    A = host.DiracOp(ctx, host.Hopping(ctx, dims, faces=faces), 1.0, diag=diag)
    Ao = orc.dirac(orc.hopping(dims, faces), 1.0, diag)
    rhs = orc.init_rand(0, n)
    for prm in ((0, 5, 400, 1e-10), (6, 0, 400, 1e-10)):
        xo, ho, ito = orc.gcr_solve(Ao, orc.gcr_param(*prm), rhs)
        env, spread = reference_envelope(perturbed(orc, Ao, orc.gcr_param(*prm), rhs), ho, ito)
        x = ctx.field(n).set_zero()
        it, hist = host.GCR(ctx, A, host.GCR_Param(*prm, False, None, None)).solve(ctx.from_numpy(rhs), x)
        check_hist(hist, ho, it, ito, env, spread)
        assert relerr(x.numpy(), xo) < 1e-8


def aniso_levels(dims, subs, n_eigen):
    lv, cur, ncol = [], list(dims), 1
    for sub, ne in zip(subs, n_eigen):
        lv.append(dict(site_dims=[1] + cur, sub=[1] + list(sub), n_spin=1, n_col=ncol, n_eigen=ne))
        cur = [d // s for d, s in zip(cur, sub)]
        ncol = ne
    return lv


def test_multigrid_on_the_anisotropic_operator(ctx, host, orc):
    """configs[4] at 1/16384 of its volume: semi-coarsened aggregates along the strongly coupled direction first"""
    dims = [32, 16, 32]
    n = int(np.prod(dims))
    faces, diag = host.synthetic_bonds(dims)
    A = host.DiracOp(ctx, host.Hopping(ctx, dims, faces=faces), 1.0, diag=diag)
    Ao = orc.dirac(orc.hopping(dims, faces), 1.0, diag)
    lv = aniso_levels(dims, [(1, 1, 8), (2, 2, 4), (4, 4, 1)], [2, 4, 4])
    eig, coarse, smooth = (0, 10, 10, 1e-8), (0, 10, 2, 1e-2), (0, 4, 2, 1e-8)
    mg = host.MG(ctx, A, lv, host.GCR_Param(*eig), host.GCR_Param(*coarse), host.GCR_Param(*smooth))
    mo = orc.MG(Ao, lv, orc.gcr_param(*eig), orc.gcr_param(*coarse), orc.gcr_param(*smooth))
    for l in range(3):
        assert np.array_equal(mg.block_map(l), mo.block_map(l))
        br, bc, bv = mg.coarse(l)
        bro, bco, bvo = mo.coarse(l)
        assert np.array_equal(br, bro) and np.array_equal(bc, bco)          # sparsity pattern: bit-exact
    # the level-0 prolongator comes from the same inverse iteration: equal to rounding
    assert relerr(mg.prolongator(0), mo.prolongator(0)) < 1e-6
    # Galerkin identity on the device objects
    v = ctx.init_rand(3, n)
    rv = mg.restrict(v, 0)
    assert relerr(mg.coarse_op(0)(rv).numpy(), mg.restrict(A(mg.expand(rv, 0)), 0).numpy()) < 1e-13
    rhs = orc.init_rand(0, n)
    x = ctx.field(n).set_zero()
    # solved to 1e-12 so that the two solutions can be compared at 1e-8 (condition number ~500)
    it, hist = host.GCR(ctx, A, host.GCR_Param(0, 3, 300, 1e-12, False, None, mg)).solve(ctx.from_numpy(rhs), x)
    assert hist[-1] <= 1e-12 and relerr(A(x).numpy(), rhs) < 1.5e-12
    xo, ho, ito = orc.gcr_solve(Ao, orc.gcr_param(0, 3, 300, 1e-12), rhs, precond=mo.as_op())
    assert abs(it - ito) <= 2
    assert np.max(np.abs(hist[:4] - ho[:4]) / ho[:4]) < 1e-6
    assert relerr(x.numpy(), xo) < 1e-8
    # and it beats plain GCR by a wide margin
    x0 = ctx.field(n).set_zero()
    it0, _ = host.GCR(ctx, A, host.GCR_Param(0, 10, 3000, 1e-10, False, None, None)).solve(ctx.from_numpy(rhs), x0)
    assert it * 3 < it0


def test_bench_parameterisation_aniso_against_oracle_64(ctx, host):
    """mg3d_aniso's parameters (MG_DEFAULT cycle, semi-coarsened aggregates 1x1x8 -> 2x2x4 -> ..., 2/4/4 near-null vectors,
    outer restart 3, 1e-10) at 64^3 with the level shapes bench.py's CPU arm uses, against the CPU restatement on the same
    near-null vectors: north-star gates (tests/test_gpu_mg.py::assert_north_star)."""
    import bench
    from oracle import parity
    from test_gpu_mg import assert_north_star
    wl = bench.WORKLOADS["mg3d_aniso"]
    dims = wl["cpu_sample"]
    subs, nes = wl["cpu_mg"]["subs"], wl["cpu_mg"]["n_eigen"]
    lv = bench.scalar_levels(dims, subs, nes)
    A, Ao = parity.operators(host, ctx, dims, aniso=wl["aniso"])
    ref = parity.oracle_solve(Ao, lv, wl["mg"], wl["restart"], wl["max_iter"], wl["tol"])
    env, spread = parity.oracle_envelope(Ao, ref, nper=1)
    gpu = parity.gpu_solve(host, ctx, A, lv, wl["mg"], wl["restart"], wl["max_iter"], wl["tol"], ref["rhs"], nearnull=ref["nearnull"])
    c = assert_north_star(gpu, ref, env, spread)
    print("parity aniso 64^3:", c, "own envelope max %.2e" % env.max())
