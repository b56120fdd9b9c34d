"""Microbenchmark of the block-CSR apply through the C ABI: 7-point block stencil on an n^3 lattice of ne x ne blocks
(random values), assembly-layout kernel vs the sliced image streamed through the bulk-copy ring; results compared bit for
bit.  usage: bcsr_bench.py ne n [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mgpreconditionedgcr_b200 import host  # noqa: E402

ne, n = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
nb = n ** 3
idx = np.arange(nb, dtype=np.int64)
z, y, x = idx // (n * n), (idx // n) % n, idx % n
cols = np.stack([idx - n * n, idx - n, idx - 1, idx, idx + 1, idx + n, idx + n * n], axis=1)
mask = np.stack([z > 0, y > 0, x > 0, idx >= 0, x < n - 1, y < n - 1, z < n - 1], axis=1)
brow = np.zeros(nb + 1, dtype=np.int64)
np.cumsum(mask.sum(axis=1), out=brow[1:])
bcol = cols[mask]
rng = np.random.default_rng(0)
bval = (rng.random((len(bcol), ne, ne)) + 1j * rng.random((len(bcol), ne, ne))).astype(np.complex128)
ctx = host.Context(0)
xv = ctx.init_rand(1, nb * ne)
res = {}
for name, rows in (("assembly", 1 << 62), ("ring", 0)):
    ctx.set_option("blockcsr_ring_rows", rows)
    A = host.HierarchicalSparse(ctx, nb, ne, brow, bcol, bval)
    yv = ctx.field(nb * ne)
    for _ in range(3):
        A(xv, out=yv)
    ctx.sync()
    ctx.set_profile(True)
    for _ in range(reps):
        A(xv, out=yv)
    ctx.sync()
    p = ctx.profile()["blockcsr_apply"]
    ctx.set_profile(False)
    res[name] = yv.numpy()
    print("ne=%d n=%d stages=%s %s: %.1f us %.0f GB/s" % (ne, n, os.environ.get("MGCR_BLOCKCSR_STAGES", "-"), name, 1e3 * p["ms"] / p["calls"],
                                                       p["bytes"] / (p["ms"] * 1e-3) / 1e9), flush=True)
    A.destroy()
print("exact=%s" % bool(np.array_equal(res["assembly"], res["ring"])))
