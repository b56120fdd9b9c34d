"""CPU tests of the library's host-side logic that needs no GPU (the C-ABI library loads without a device; compute entry
points are not called)."""
import numpy as np
import pytest

from conftest import GOLD  # noqa: F401


@pytest.fixture(scope="module")
def lib():
    from mgpreconditionedgcr_b200 import capi
    return capi.load()


@pytest.mark.parametrize("seed", [0, 1, 2, 9, 42])
def test_rand_stream_jump_ahead_is_the_glibc_stream(lib, golden, seed):
    """Field::init_rand(seed) slabs: csrc/rand.cu jumps over the prefix with powers of the generator's companion matrix and
    draws chunks on several threads; every slab must be the exact tail of the stream the reference draws with srand/rand
    (tests/golden/c1_apply.npz holds the first 1024 elements from the unmodified reference)"""
    from mgpreconditionedgcr_b200 import capi
    from oracle import pyoracle as orc
    n = 300000
    ref = orc.init_rand(seed, n)
    assert np.array_equal(ref[:1024], golden.c1_apply["rand_seed%d" % seed])
    out = np.empty(n, dtype=np.complex128)
    capi.check(lib.mgcr_rand_stream(seed, 0, n, capi.ptr(out)))
    assert np.array_equal(out, ref)
    for skip in (1, 30, 31, 65535, 65536, 131073, 299999):
        o = np.empty(n - skip, dtype=np.complex128)
        capi.check(lib.mgcr_rand_stream(seed, skip, n - skip, capi.ptr(o)))
        assert np.array_equal(o, ref[skip:]), skip


def test_rand_stream_far_jump_matches_a_chain_of_short_ones(lib):
    """a slab far into the stream (rank 7 of a 512^3 field starts 2 x 117 M draws in) equals the same elements reached through
    a different decomposition of the jump"""
    from mgpreconditionedgcr_b200 import capi
    skip, n = 117440512, 4096
    a = np.empty(n, dtype=np.complex128)
    capi.check(lib.mgcr_rand_stream(0, skip, n, capi.ptr(a)))
    b = np.empty(n + 1000, dtype=np.complex128)
    capi.check(lib.mgcr_rand_stream(0, skip - 1000, n + 1000, capi.ptr(b)))
    assert np.array_equal(a, b[1000:])
    assert np.all(np.abs(a.real) <= 1) and np.all(np.abs(a.imag) <= 1) and len(np.unique(a)) > n // 2


def test_bench_parity_fixture_is_consistent(golden):
    """tests/golden/mg_bench_128.npz (oracle/make_golden_bench.py): converged history, finite envelope, samples of every level's
    near-null vectors"""
    g = golden.mg_bench_128
    h = g["hist"]
    assert h[0] == 1.0 and h[-1] <= 1e-10 and h[-2] > 1e-10 and len(h) == int(g["iters"]) + 1
    assert np.all(np.isfinite(g["env"])) and g["env"].max() < 1e-7 and len(g["env"]) == len(h)
    assert all(("nearnull%d_sample" % l) in g for l in range(3))
    assert abs(np.linalg.norm(g["x_sample"])) > 0 and float(g["x_norm"]) > 0
