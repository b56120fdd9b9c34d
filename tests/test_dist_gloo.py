"""N > 1 host logic on CPU (gloo, world_size 2): the slab partition the library hands out (mgcr_slab_range, aligned to the
aggregate size) drives a two-process emulation of the distributed stencil apply and inner product -- halo planes over
gloo send/recv, partial sums all-reduced -- and must reproduce the global oracle result exactly / to rounding.  The CUDA
side of the same path (NCCL) is tests/test_gpu_dist.py."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    from mgpreconditionedgcr_b200 import host
    from oracle import pyoracle as orc
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    dims = [24, 6, 10]
    align = 4
    plane = dims[1] * dims[2]
    V = int(np.prod(dims))
    b, e = host.slab_range(dims[0], align, rank, world)
    ranges = [None] * world
    dist.all_gather_object(ranges, (b, e))
    ok = ranges[0][0] == 0 and ranges[-1][1] == dims[0] and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    ok = ok and all(r[0] % align == 0 and r[1] % align == 0 and r[1] > r[0] for r in ranges)
    k = 1 / 6.01
    f = orc.init_rand(1, V)
    g = orc.init_rand(3, V)
    ref = orc.dirac(orc.hopping(dims), k)(f)
    # local slab with one ghost plane on each side, filled by the neighbours
    x = np.zeros(((e - b) + 2, plane), dtype=np.complex128)
    x[1:-1] = f[b * plane:e * plane].reshape(e - b, plane)
    reqs = []
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(x[1]).view(np.float64)), rank - 1))
    if rank + 1 < world:
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(x[-2]).view(np.float64)), rank + 1))
    if rank > 0:
        t = torch.empty(2 * plane, dtype=torch.float64); dist.recv(t, rank - 1); x[0] = t.numpy().view(np.complex128)
    if rank + 1 < world:
        t = torch.empty(2 * plane, dtype=torch.float64); dist.recv(t, rank + 1); x[-1] = t.numpy().view(np.complex128)
    for r in reqs:
        r.wait()
    # apply the Dirichlet hopping stencil to the padded slab in the neighbour order of the device kernel (z-1, y-1, x-1, x+1, y+1, z+1)
    X = x.reshape(e - b + 2, dims[1], dims[2])
    pad = np.zeros((e - b + 2, dims[1] + 2, dims[2] + 2), dtype=np.complex128)
    pad[:, 1:-1, 1:-1] = X
    s = pad[:-2, 1:-1, 1:-1] + pad[1:-1, :-2, 1:-1]
    s = s + pad[1:-1, 1:-1, :-2]
    s = s + pad[1:-1, 1:-1, 2:]
    s = s + pad[1:-1, 2:, 1:-1]
    s = s + pad[2:, 1:-1, 1:-1]
    y = X[1:-1] - k * s
    ok = ok and np.array_equal(y.reshape(-1), ref[b * plane:e * plane])
    part = np.vdot(f[b * plane:e * plane], g[b * plane:e * plane])
    t = torch.tensor([part.real, part.imag], dtype=torch.float64)
    dist.all_reduce(t)
    full = np.vdot(f, g)
    ok = ok and abs(complex(t[0].item(), t[1].item()) - full) < 1e-12 * abs(full)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_slab_partition_drives_a_two_process_stencil_apply():
    world = 2
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctxm.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_reference_arm_runs_on_rank0_only():
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""
