"""CPU tests of the variable-coefficient synthetic operator of BASELINE.json configs[4]: the generator
(host.synthetic_bonds, slab-independent counter-based hash), the oracle's restatement of it (hopping CSR with real bond
values wrapped as diag - k H), and -- when oracle/_ref is present -- the unmodified reference's GCR on the assembled
Sparse A = diag - H, which the DiracOp form reproduces to one reassociation of each row sum."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, check_hist, perturbed, reference_envelope
from mgpreconditionedgcr_b200 import host
from oracle import pyoracle as orc

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_oracle")


def assembled_csr(dims, faces, diag):
    """A = diag - H as the Sparse a user of the reference would assemble: ascending columns, diagonal in place"""
    row, col, val = host.hopping_csr(dims, faces)
    n = len(row) - 1
    r2, c2, v2 = [0], [], []
    d = diag.reshape(-1)
    for i in range(n):
        cs = list(col[row[i]:row[i + 1]]) + [i]
        vs = list(-val[row[i]:row[i + 1]]) + [complex(d[i])]
        order = np.argsort(cs)
        c2 += [cs[j] for j in order]
        v2 += [vs[j] for j in order]
        r2.append(len(c2))
    return np.array(r2, np.int64), np.array(c2, np.int64), np.array(v2, np.complex128)


def test_generator_symmetric_positive_row_sum_m2_and_slab_independent():
    dims = [6, 5, 7]
    faces, diag = host.synthetic_bonds(dims, m2=0.02)
    row, col, val = host.hopping_csr(dims, faces)
    n = int(np.prod(dims))
    Hd = np.zeros((n, n))
    for i in range(n):
        Hd[i, col[row[i]:row[i + 1]]] = val[row[i]:row[i + 1]].real
    assert np.array_equal(Hd, Hd.T)
    A = np.diag(diag.reshape(-1)) - Hd
    assert np.allclose(A.sum(axis=1), 0.02, atol=1e-14)
    assert np.linalg.eigvalsh(A)[0] > 0.0199
    # anisotropy of the bonds: slowest dim 1e-4, fastest 1
    assert faces[0].max() < 2e-4 and faces[2][:, :, :-1].min() > 0.5
    for zr in ((0, 2), (2, 5), (5, 6)):
        f2, d2 = host.synthetic_bonds(dims, m2=0.02, z_range=zr)
        assert all(np.array_equal(a[zr[0]:zr[1]], b) for a, b in zip(faces, f2)) and np.array_equal(diag[zr[0]:zr[1]], d2)


def test_bench_device_generator_equals_host_generator():
    torch = pytest.importorskip("torch")
    import bench
    dims = [16, 6, 10]
    for zr in ((0, 16), (4, 8), (12, 16)):
        f, d = host.synthetic_bonds(dims, z_range=zr)
        tf, td = bench.device_synthetic_bonds(torch, dims, zr[0], zr[1], (1e-4, 1e-2, 1.0), 0.5, 0.01, 12345, "cpu")
        assert all(np.allclose(a.reshape(-1), b.numpy(), rtol=1e-15, atol=0) for a, b in zip(f, tf))
        assert np.allclose(d.reshape(-1), td.numpy(), rtol=1e-15, atol=0)


@pytest.mark.parametrize("dims", [[5, 6, 7], [9, 11], [13]])
def test_oracle_variable_hopping_matches_python_csr(dims):
    rng = np.random.default_rng(3)
    faces = [0.5 + rng.random(int(np.prod(dims))) for _ in dims]
    row, col, val = host.hopping_csr(dims, faces)
    ro, co, vo = orc.csr_export(orc.hopping(dims, faces))
    assert np.array_equal(row, ro) and np.array_equal(col, co) and np.array_equal(val, vo)
    # unit bonds reproduce the unit hopping matrix
    ones = [np.ones(int(np.prod(dims))) for _ in dims]
    r1, c1, v1 = orc.csr_export(orc.hopping(dims, ones))
    r0, c0, v0 = orc.csr_export(orc.hopping(dims))
    assert np.array_equal(r1, r0) and np.array_equal(c1, c0) and np.array_equal(v1, v0)


def test_dirac_form_equals_assembled_sparse_to_rounding():
    dims = [6, 8, 10]
    n = int(np.prod(dims))
    faces, diag = host.synthetic_bonds(dims)
    A = orc.dirac(orc.hopping(dims, faces), 1.0, diag)
    As = orc.csr(n, n, *assembled_csr(dims, faces, diag))
    x = orc.init_rand(4, n)
    a, b = A(x), As(x)
    assert np.max(np.abs(a - b)) <= 8 * np.finfo(float).eps * np.max(np.abs(b))


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="compiled reference (oracle/_ref) not present")
def test_anisotropic_gcr_against_live_reference():
    """the unmodified reference's GCR on the assembled Sparse A = diag - H vs the restatement's diag - k H form"""
    dims = [6, 8, 10]
    n = int(np.prod(dims))
    faces, diag = host.synthetic_bonds(dims, m2=0.5)
    row, col, val = assembled_csr(dims, faces, diag)
    rhs = orc.init_rand(0, n)
    x0 = np.zeros(n, dtype=np.complex128)
    A = orc.dirac(orc.hopping(dims, faces), 1.0, diag)
    with tempfile.TemporaryDirectory() as d:
        for name, a in (("row", row), ("col", col), ("val", val), ("rhs", rhs), ("x0", x0)):
            a.tofile(os.path.join(d, name + ".bin"))
        for trunc, restart, max_iter in ((0, 5, 200), (4, 0, 200)):
            subprocess.check_call([REF_BIN, "gcr-file", d, str(n), str(len(col)), "0", "0", str(trunc), str(restart), str(max_iter), "1e-10"],
                                  stdout=subprocess.DEVNULL)
            ref_hist = np.fromfile(os.path.join(d, "hist.bin"))
            ref_x = np.fromfile(os.path.join(d, "x.bin"), dtype=np.complex128)
            # bit for bit on the same assembled matrix
            As = orc.csr(n, n, row, col, val)
            prm = orc.gcr_param(trunc, restart, max_iter, 1e-10)
            xs, hs, its = orc.gcr_solve(As, prm, rhs)
            assert np.array_equal(hs, ref_hist) and np.array_equal(xs, ref_x)
            # within the north-star gates in the matrix-free form (row sums reassociated), as far as the reference's own
            # 1e-16-perturbed histories stay inside them (conftest.reference_envelope)
            env, spread = reference_envelope(perturbed(orc, As, prm, rhs), ref_hist, its)
            x, hist, it = orc.gcr_solve(A, prm, rhs)
            check_hist(hist, ref_hist, it, its, env, spread)
            assert np.linalg.norm(x - ref_x) / np.linalg.norm(ref_x) < 1e-8


def test_multigrid_on_anisotropic_operator_oracle():
    """line aggregates along the strong direction first: the restatement's MG-GCR converges and beats plain GCR"""
    dims = [16, 8, 32]
    n = int(np.prod(dims))
    faces, diag = host.synthetic_bonds(dims)
    A = orc.dirac(orc.hopping(dims, faces), 1.0, diag)
    lv = [dict(site_dims=[1, 16, 8, 32], sub=[1, 1, 1, 8], n_spin=1, n_col=1, n_eigen=2),
          dict(site_dims=[1, 16, 8, 4], sub=[1, 2, 2, 4], n_spin=1, n_col=2, n_eigen=4)]
    mg = orc.MG(A, lv, orc.gcr_param(0, 10, 10, 1e-8), orc.gcr_param(0, 10, 2, 1e-2), orc.gcr_param(0, 4, 2, 1e-8))
    rhs = orc.init_rand(0, n)
    x, hist, it = orc.gcr_solve(A, orc.gcr_param(0, 3, 300, 1e-10), rhs, precond=mg.as_op())
    assert hist[-1] <= 1e-10 and np.linalg.norm(A(x) - rhs) / np.linalg.norm(rhs) < 1.5e-10
    _, _, it0 = orc.gcr_solve(A, orc.gcr_param(0, 10, 3000, 1e-10), rhs)
    assert it * 3 < it0


def test_bench_anisotropic_hierarchy_shapes_and_slab_alignment():
    """bench.py's mg3d_aniso: five levels 1024x512x512 -> 1024x512x64 -> 512x256x16 -> 128x64x4 -> 32x16x1, every aggregate
    shape divides its lattice, and at 8 GPUs the slabs are multiples of the product of the aggregate sizes along the
    partitioned dimension of the levels that stay distributed"""
    import bench
    wl = bench.WORKLOADS["mg3d_aniso"]
    lv = bench.scalar_levels(wl["dims"], wl["mg"]["subs"], wl["mg"]["n_eigen"])
    assert [l["site_dims"][1:] for l in lv] == [[1024, 512, 512], [1024, 512, 64], [512, 256, 16], [128, 64, 4]]
    assert [l["n_col"] for l in lv] == [1, 2, 4, 4] and [l["n_eigen"] for l in lv] == [2, 4, 4, 4]
    for l in lv:
        assert all(d % s == 0 for d, s in zip(l["site_dims"], l["sub"]))
    coarsest = [d // s for d, s in zip(lv[-1]["site_dims"][1:], lv[-1]["sub"][1:])]
    assert coarsest == [32, 16, 1]
    align = 1
    for sub in wl["mg"]["subs"]:
        if wl["dims"][0] // (align * sub[0]) >= 8:
            align *= sub[0]
    assert align == 32 and (wl["dims"][0] // 8) % align == 0
    # the CPU sample of the reference arm keeps the aggregate shapes of the first levels
    cl = bench.scalar_levels(wl["cpu_sample"], wl["cpu_mg"]["subs"], wl["cpu_mg"]["n_eigen"])
    assert [l["sub"] for l in cl][:2] == [l["sub"] for l in lv][:2]
