#!/usr/bin/env python
"""bench.py -- time-to-solution of the GCR / MG-GCR solve path on synthetic operators of BASELINE.json's shapes.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--operator csr|stencil] [--impl reference]

A "step" is one full solve (x0 = 0, relative residual 1e-10) of the named workload with the operator, the right-hand
side and every Krylov vector resident in HBM.  `value` = seconds per solve (device time, CUDA events on the
library's stream, max over ranks); `e2e` = the same solve through the C ABI's host-buffer entry point
(mgcr_gcr_solve_host: pinned host rhs/x0 -> device -> solve -> host).  `roofline` is measured live: every kernel
launch of the timed steps is bracketed by pooled CUDA events inside the library (no synchronisation in the loop) and
the dominant kernel class is reported as algorithmic bytes / device time against MEASURED_PEAKS.json.

--impl reference times the reference's own CPU implementation (oracle/_ref/ref_oracle = the unmodified reference
headers compiled by oracle/Makefile; the C restatement if that binary is absent) on a bounded sample of the same
workload on the host cores, extrapolated by rows x iterations to the same time-to-solution metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> description of the synthetic operator A = I - k H (H = unit hopping, Dirichlet), k = 1/(2 nd + m2)
WORKLOADS = {
    # BASELINE.json configs[1]
    "gcr2d_4096": dict(dims=[4096, 4096], m2=0.01, restart=10, tol=1e-10, max_iter=100000, mg=None,
                       desc="2-D 5-point 4096x4096, I - H/(4+0.01), unpreconditioned restart-10 GCR to 1e-10",
                       cpu_sample=[1024, 1024], cpu_iters=10),
    "gcr3d_256": dict(dims=[256, 256, 256], m2=0.01, restart=10, tol=1e-10, max_iter=100000, mg=None,
                      desc="3-D 7-point 256^3, I - H/(6+0.01), unpreconditioned restart-10 GCR to 1e-10",
                      cpu_sample=[96, 96, 96], cpu_iters=10),
    "gcr3d_512": dict(dims=[512, 512, 512], m2=0.01, restart=10, tol=1e-10, max_iter=100000, mg=None,
                      desc="3-D 7-point 512^3, I - H/(6+0.01), unpreconditioned restart-10 GCR to 1e-10",
                      cpu_sample=[96, 96, 96], cpu_iters=10),
}
# cycle parameters from the B200 sweeps in profiles/r01_mg_param_sweep*.txt: two smoother iterations, at most two coarse GCR
# iterations per level (a short K-cycle), outer restart 3 -- 512^3: 1.43 s against 6.3 s for coarse max_iter 20 / restart 10
MG_DEFAULT = dict(eigen=(0, 10, 10, 1e-8), coarse=(0, 10, 2, 1e-2), smooth=(0, 4, 2, 1e-8))
WORKLOADS.update({
    # BASELINE.json configs[2]: 3-level hierarchy 256^3 -> 64^3 -> 16^3
    "mg3d_256": dict(dims=[256, 256, 256], m2=0.01, restart=3, tol=1e-10, max_iter=1000,
                     mg=dict(subs=[4, 4], n_eigen=[4, 4], **MG_DEFAULT),
                     desc="3-D 7-point 256^3, I - H/(6+0.01), 3-level MG (4^3 aggregates, 4 near-null vectors) preconditioned restart-3 GCR to 1e-10",
                     cpu_sample=[64, 64, 64], cpu_iters=0),
    # BASELINE.json configs[3]: 4-level hierarchy 512^3 -> 128^3 -> 32^3 -> 8^3
    "mg3d_512": dict(dims=[512, 512, 512], m2=0.01, restart=3, tol=1e-10, max_iter=1000,
                     mg=dict(subs=[4, 4, 4], n_eigen=[4, 4, 4], **MG_DEFAULT),
                     desc="3-D 7-point 512^3, I - H/(6+0.01), 4-level MG (4^3 aggregates, 4 near-null vectors) preconditioned restart-3 GCR to 1e-10",
                     cpu_sample=[64, 64, 64], cpu_iters=0),
})


# BASELINE.json configs[4]: anisotropic variable-coefficient operator (generator: host.synthetic_bonds),
# A = diag - H with bonds eps_d * face-averaged exp(0.5 g_d), eps = (1e-4, 1e-2, 1) from the slowest to the fastest dim,
# diag = sum of the site's bonds + 0.01.  Five levels 1024x512x512 -> 1024x512x64 -> 512x256x16 -> 128x64x4 -> 32x16x1: the
# first aggregates are lines along the strongly coupled direction (the y / z oscillations of an isotropic 4^3 aggregate are
# near-null vectors the coarse space would miss: 118 outer iterations against 42 on the CPU oracle at 64^3).
ANISO = dict(eps=(1e-4, 1e-2, 1.0), sigma=0.5, m2=0.01, seed=12345)
WORKLOADS.update({
    "mg3d_aniso": dict(dims=[1024, 512, 512], aniso=ANISO, restart=3, tol=1e-10, max_iter=1000,
                       mg=dict(subs=[(1, 1, 8), (2, 2, 4), (4, 4, 4), (4, 4, 4)], n_eigen=[2, 4, 4, 4], **MG_DEFAULT),
                       desc="3-D 7-point anisotropic variable-coefficient 1024x512x512 (bonds eps_d*exp(0.5 g), eps = 1e-4/1e-2/1, diag = sum + 0.01), "
                            "5-level MG (aggregates 1x1x8, 2x2x4, 4^3, 4^3; 2/4/4/4 near-null vectors) preconditioned restart-3 GCR to 1e-10",
                       cpu_sample=[64, 64, 64], cpu_mg=dict(subs=[(1, 1, 8), (2, 2, 4), (4, 4, 2)], n_eigen=[2, 4, 4]), cpu_iters=0),
    # the same operator and hierarchy shape at 1/8 of the volume (fits the single-GPU profiling passes comfortably)
    "mg3d_aniso_512": dict(dims=[512, 256, 512], aniso=ANISO, restart=3, tol=1e-10, max_iter=1000,
                           mg=dict(subs=[(1, 1, 8), (2, 2, 4), (4, 4, 4), (4, 4, 4)], n_eigen=[2, 4, 4, 4], **MG_DEFAULT),
                           desc="3-D 7-point anisotropic variable-coefficient 512x256x512, 5-level MG preconditioned restart-3 GCR to 1e-10",
                           cpu_sample=[64, 64, 64], cpu_mg=dict(subs=[(1, 1, 8), (2, 2, 4), (4, 4, 2)], n_eigen=[2, 4, 4]), cpu_iters=0),
})


def scalar_levels(dims, subs, n_eigen):
    """MG level configs of a 3-D scalar lattice: site_dims = [1, nz, ny, nx]; the dof of a coarse level is the number of
    near-null vectors of the level above (coarse index = block * ne + e, reference src/MG.h:359,378).  An entry of subs is
    the aggregate edge, or a (z, y, x) tuple."""
    lv, cur, ncol = [], list(dims), 1
    for sub, ne in zip(subs, n_eigen):
        sub3 = [sub] * 3 if isinstance(sub, int) else list(sub)
        lv.append(dict(site_dims=[1] + cur, sub=[1] + sub3, n_spin=1, n_col=ncol, n_eigen=ne))
        cur = [d // q for d, q in zip(cur, sub3)]
        ncol = ne
    return lv


def device_synthetic_bonds(torch, dims, z0, z1, eps, sigma, m2, seed, device):
    """host.synthetic_bonds for planes [z0, z1) computed on the GPU with torch integer ops (input generation only: the
    hash is a pure function of the global site index, so every rank builds its slab independently).  Returns
    (faces[3], diag) as flat float64 CUDA tensors."""
    nzg, ny, nx = dims
    M = (1 << 64) - 1

    def s64(v):   # two's-complement image of a 64-bit constant
        v &= M
        return v - (1 << 64) if v >= (1 << 63) else v

    def hash_unit(site, d):
        z = site * 3 + d + s64(seed * 0x9E3779B97F4A7C15)
        z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * s64(0xBF58476D1CE4E5B9)
        z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * s64(0x94D049BB133111EB)
        z = z ^ ((z >> 31) & ((1 << 33) - 1))
        return ((z >> 11) & ((1 << 53) - 1)).to(torch.float64) * (2.0 / (1 << 53)) - 1.0

    zz = torch.arange(z0, z1, dtype=torch.int64, device=device)[:, None, None]
    site = (zz * ny + torch.arange(ny, dtype=torch.int64, device=device)[None, :, None]) * nx \
        + torch.arange(nx, dtype=torch.int64, device=device)[None, None, :]
    stride = (ny * nx, nx, 1)

    def bond(st, d):
        f = eps[d] * ((torch.exp(sigma * hash_unit(st, d)) + torch.exp(sigma * hash_unit(st + stride[d], d))) * 0.5)
        return f

    faces = []
    for d in range(3):
        f = bond(site, d)
        if d == 0:
            f[(zz == nzg - 1).expand_as(f)] = 0.0
        elif d == 1:
            f[:, ny - 1, :] = 0.0
        else:
            f[:, :, nx - 1] = 0.0
        faces.append(f)
    fz, fy, fx = faces
    diag = torch.full_like(fz, m2)
    diag += fz
    diag[1:] += fz[:-1]
    if z0 > 0:   # bond to the lower slab neighbour's last plane
        lo = bond(site[0:1] - stride[0], 0)
        diag[0:1] += lo
    diag += fy
    diag[:, 1:, :] += fy[:, :-1, :]
    diag += fx
    diag[:, :, 1:] += fx[:, :, :-1]
    return [f.reshape(-1) for f in faces], diag.reshape(-1)


# iteration counts measured on B200 (parity-checked against the CPU oracle at reduced size); used by the reference
# arm, which cannot afford the full CPU solve, to extrapolate time-to-solution.  Updated from BENCH logs.
KNOWN_ITERS = {}
KNOWN_ITERS_FILE = os.path.join(ROOT, "profiles", "bench_iterations.json")
if os.path.exists(KNOWN_ITERS_FILE):
    KNOWN_ITERS.update(json.load(open(KNOWN_ITERS_FILE)))

METRIC = "MG-GCR solve time to 1e-10 & SpMV HBM GB/s at 1/2/4/8 B200 vs CPU ref"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def reference_cpu_sample(wl, host=None, ctx=None):
    """times the reference's CPU solve on a bounded sample of the workload; returns dict with seconds per row-iteration.
    MG workloads with a GPU context at hand (host, ctx): the same oracle run is also the parity reference of the GPU path at
    that size (oracle/parity.py), returned under "parity"."""
    dims = wl["cpu_sample"]
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_oracle")
    V = 1
    for d in dims:
        V *= d
    if wl.get("mg"):
        # The reference's MG is two-level, 6-D only and returns an uninitialised buffer: it cannot
        # run this workload.  The CPU arm is the C restatement of the same algorithm (oracle/mgcr_oracle.c), 1 thread.
        from oracle import parity
        m = dict(wl["mg"], **wl.get("cpu_mg", {}))   # a hierarchy of the same shape that fits the sample lattice
        nlev = len(m["subs"]) + 1
        par = None
        if ctx is not None:
            par, r = parity.bench_parity(host, ctx, wl, dims, m["subs"], m["n_eigen"])
        else:
            _, Ao = parity.operators(None, None, dims, m2=wl.get("m2", 0.01), aniso=wl.get("aniso")) if not wl.get("aniso") else parity.operators(_host_mirror(), None, dims, aniso=wl["aniso"])
            r = parity.oracle_solve(Ao, parity.levels(dims, m["subs"], m["n_eigen"]), wl["mg"], wl["restart"], wl["max_iter"], wl["tol"])
        it = r["iters"]
        return dict(kind="port", cores=1, V=V, iters=it, sec_per_iter=r["solve_s"] / max(it, 1), spmv_seconds=None, setup_seconds=r["arnoldi_s"] + r["setup_s"],
                    parity=par,
                    sample="C restatement (oracle/mgcr_oracle.c, 1 thread; the reference's own MG cannot run 3-D multi-level problems) of the "
                           "same %d-level MG-GCR on %s: %d outer iterations to %.1e in %.1f s (+ %.1f s set-up)"
                           % (nlev, "x".join(map(str, dims)), it, r["hist"][-1], r["solve_s"], r["arnoldi_s"] + r["setup_s"]))
    if os.path.exists(ref):
        cwd = os.path.join(ROOT, "oracle", "_ref", "work", "run", "a")
        os.makedirs(cwd, exist_ok=True)
        os.makedirs(os.path.join(ROOT, "oracle", "_ref", "work", "data", "out_data"), exist_ok=True)
        cmd = [ref, "bench", str(len(dims))] + [str(d) for d in dims] + [str(wl["m2"]), str(wl["restart"]), str(wl["cpu_iters"])]
        out = subprocess.run(cmd, cwd=cwd, check=True, capture_output=True, text=True).stdout
        js = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
        return dict(kind="reference", cores=1, V=V, iters=js["iters"], sec_per_iter=js["seconds_per_iter"], spmv_seconds=js["spmv_seconds"],
                    sample="unmodified reference GCR (1 thread, as in src/GCR.h) on %s, restart %d, %d iterations" % ("x".join(map(str, dims)), wl["restart"], js["iters"]))
    # restatement fallback (port)
    from oracle import pyoracle as orc
    H = orc.hopping(dims)
    A = orc.dirac(H, 1.0 / (2 * len(dims) + wl["m2"]))
    rhs = orc.init_rand(0, H.n)
    t0 = time.perf_counter()
    _, _, it = orc.gcr_solve(A, orc.gcr_param(0, wl["restart"], wl["cpu_iters"], 1e-10), rhs)
    dt = time.perf_counter() - t0
    return dict(kind="port", cores=1, V=V, iters=it, sec_per_iter=dt / max(it, 1), spmv_seconds=None,
                sample="C restatement of the reference GCR (1 thread) on %s, restart %d, %d iterations" % ("x".join(map(str, dims)), wl["restart"], it))


def _host_mirror():
    from mgpreconditionedgcr_b200 import host
    return host


def c1_sample_reference():
    """BASELINE configs[0]: the shipped 4^4 sample operator, DiracOp(k = 0.15292), GCR_Param(0,5,4000,1e-13), rhs = init_rand(0),
    x0 = 0 (src/main.cpp:834-875 on the 4^4 data) solved by the UNMODIFIED reference (oracle/_ref/ref_oracle gcr-file): the one
    configuration where the reference itself runs the whole solve -- a like-for-like CPU number, no extrapolation."""
    import tempfile
    import numpy as np
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_oracle")
    if not os.path.exists(ref):
        return None
    m = np.load(os.path.join(ROOT, "tests", "golden", "c1_matrix.npz"))
    from oracle import pyoracle as orc
    d = tempfile.mkdtemp(prefix="c1_ref_")
    m["row"].astype(np.int64).tofile(os.path.join(d, "row.bin"))
    m["col"].astype(np.int64).tofile(os.path.join(d, "col.bin"))
    m["val"].astype(np.complex128).tofile(os.path.join(d, "val.bin"))
    orc.init_rand(0, 3072).tofile(os.path.join(d, "rhs.bin"))
    np.zeros(3072, dtype=np.complex128).tofile(os.path.join(d, "x0.bin"))
    k = 0.05 + 8 * ((0.17865 - 0.05) / 10.)
    best = None
    for _ in range(3):
        out = subprocess.run([ref, "gcr-file", d, "3072", str(m["val"].size), repr(k), "0", "0", "5", "4000", "1e-13"], check=True, capture_output=True, text=True).stdout
        js = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
        best = js if best is None or js["solve_seconds"] < best["solve_seconds"] else best
    hist = np.fromfile(os.path.join(d, "hist.bin"), dtype=np.float64)
    x = np.fromfile(os.path.join(d, "x.bin"), dtype=np.complex128)
    return dict(seconds=best["solve_seconds"], iters=best["iters"], hist=hist, x=x, k=k)


def extrapolate(sample, wl, iterations):
    V = 1
    for d in wl["dims"]:
        V *= d
    return sample["sec_per_iter"] * (V / sample["V"]) * iterations


def run_reference(args, wl, rank):
    if rank != 0:
        return
    its = KNOWN_ITERS.get(args.workload)
    times = []
    sample = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        sample = reference_cpu_sample(wl)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    its_used = its if its else sample["iters"]
    value = extrapolate(sample, wl, its_used)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "c128 (f64)",
        "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"], "operator": "csr (Sparse<long>, the reference's only format)"},
        "cpu_baseline": {"value": value, "unit": "s", "cores": sample["cores"], "kind": sample["kind"],
                         "sample": sample["sample"] + "; time-to-solution extrapolated as sec/iteration x (rows/sample rows) x %s iterations%s"
                         % (its_used, "" if its else " (full-size iteration count unknown: per-sample count used)"),
                         "sec_per_iter_sample": sample["sec_per_iter"]},
        "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"nproc": os.cpu_count()},
    }
    c1 = c1_sample_reference()
    if c1 is not None:   # BASELINE configs[0]: the one problem the unmodified reference solves whole -- measured, not extrapolated
        line["other_workloads"] = {"c1_sample": {"value": c1["seconds"], "unit": "s", "kind": "reference", "cores": 1, "iterations": c1["iters"],
                                                 "final_rel_residual": float(c1["hist"][-1]),
                                                 "sample": "oracle/_ref/ref_oracle gcr-file: the unmodified reference's GCR (src/GCR.h:158-302) on data/sample_matrix 4x4parsed, whole solve"}}
    print(json.dumps(line))


NOMINAL_HBM_GBS = 8000.0   # B200 data-sheet HBM3e bandwidth (SURVEY.md 8d asks for the fraction against both peaks)


def build_problem(host, torch, ctx, wl, operator, world, rank, local_rank):
    """operator, right-hand side (the reference's Field::init_rand(0) stream, this rank's slab) and solution field"""
    import numpy as np
    dims = wl["dims"]
    nd = len(dims)
    V = int(np.prod(dims))
    if wl.get("aniso"):
        an = wl["aniso"]
        zb, ze = (0, dims[0]) if world == 1 else host.slab_range(dims[0], ctx.slab_align, rank, world)
        tf, td = device_synthetic_bonds(torch, dims, zb, ze, an["eps"], an["sigma"], an["m2"], an["seed"], "cuda:%d" % local_rank)
        torch.cuda.synchronize()
        D = host.Hopping(ctx, dims, faces_dev=[t.data_ptr() for t in tf])
        A = host.DiracOp(ctx, D, 1.0, diag_dev=td.data_ptr())
        ctx.sync()
        del tf, td
        torch.cuda.empty_cache()
    else:
        k = 1.0 / (2 * nd + wl["m2"])
        if operator == "stencil":
            D = host.Hopping(ctx, dims)
        else:
            row, col, val = host.hopping_csr(dims)
            D = host.Sparse(ctx, V, V, row, col, val)
            del row, col, val
        A = host.DiracOp(ctx, D, k)
    n_local = A.get_dim()
    if world == 1:
        rhs = ctx.init_rand(0, V)
    else:
        b, e = host.slab_range(dims[0], ctx.slab_align, rank, world)
        rhs = ctx.init_rand(0, n_local, skip=b * (V // dims[0]))
    return A, rhs, ctx.field(n_local), V


def build_mg(host, ctx, A, wl):
    """MG hierarchy of the workload + its set-up time, split by stage"""
    m = wl["mg"]
    t0 = time.perf_counter()
    mg = host.MG(ctx, A, scalar_levels(wl["dims"], m["subs"], m["n_eigen"]), host.GCR_Param(*m["eigen"]), host.GCR_Param(*m["coarse"]),
                 host.GCR_Param(*m["smooth"]))
    ctx.sync()
    wall = time.perf_counter() - t0
    st = mg.setup_profile()
    device_stages = sum(v for k, v in st.items() if k not in ("total", "rand"))
    setup = {"total_s": wall, "arnoldi_s": st.get("near_null", 0.) - st.get("rand", 0.), "rand_s": st.get("rand", 0.), "galerkin_s": st.get("galerkin", 0.),
             "aggregate_s": st.get("aggregate", 0.), "project_orthonormalise_s": st.get("project_orthonormalise", 0.),
             "ghost_prolongator_s": st.get("ghost_prolongator", 0.), "coarse_halo_gather_s": st.get("coarse_halo_gather", 0.),
             "streaming_image_s": st.get("streaming_image", 0.), "host_s": max(0., wall - device_stages),
             "how": "wall clock per stage inside mgcr_mg_create (stream synchronised at the stage boundaries), summed over the levels; host_s = what is left of the call"}
    return mg, setup


def kernel_classes(prof):
    host_side = {kname: prof.pop(kname) for kname in list(prof) if kname.startswith("host_")}
    total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    classes = {kname: {"ms_per_launch": v["ms"] / max(v["calls"], 1), "launches": v["calls"], "share": v["ms"] / total_ms,
                       "GBps": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None} for kname, v in prof.items()}
    dom = max((kname for kname in prof if not kname.startswith(("nccl_", "p2p_"))), key=lambda kname: prof[kname]["ms"])   # exchange time is reported, not rooflined
    return classes, dom, host_side


def secondary_workload(host, torch, ctx, name, operator, local_rank, peak, world=1, rank=0, dist=None):
    """one more of BASELINE.json's configurations, one warm-up + one timed solve + one profiled solve (on every rank when the run
    is distributed: device time is the maximum over the ranks)"""
    wl = dict(WORKLOADS[name])
    A, rhs, x, V = build_problem(host, torch, ctx, wl, operator, world, rank, local_rank)
    mg, setup = (None, None)
    if wl.get("mg"):
        mg, setup = build_mg(host, ctx, A, wl)
    gcr = host.GCR(ctx, A, host.GCR_Param(0, wl["restart"], wl["max_iter"], wl["tol"], False, None, mg))
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    x.set_zero(); gcr.solve(rhs, x, hist_cap=2)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.set_zero()
    e0.record(stream)
    it, _ = gcr.solve(rhs, x, hist_cap=2)
    e1.record(stream)
    ctx.sync(); torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3
    if dist is not None:
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    r = rhs - A(x)
    final_rel = r.norm() / rhs.norm()
    ctx.set_profile(True)
    x.set_zero(); gcr.solve(rhs, x, hist_cap=2)
    ctx.sync()
    classes, dom, _ = kernel_classes(ctx.profile())
    ctx.set_profile(False)
    out = {"desc": wl["desc"], "operator": operator, "rows": V, "n_gpus": world, "value": sec, "unit": "s", "iterations": it, "final_true_rel_residual": final_rel,
           "dominant_kernel": dom, "dominant_GBps": classes[dom]["GBps"], "dominant_frac": classes[dom]["GBps"] / peak,
           "kernels": {k: {"share": round(v["share"], 4), "GBps": v["GBps"]} for k, v in classes.items()}}
    op = [k for k in classes if k.startswith(("sell", "hopping"))]
    if op:
        out["spmv"] = {"kernel": op[0], "GBps": classes[op[0]]["GBps"], "frac": classes[op[0]]["GBps"] / peak, "frac_nominal": classes[op[0]]["GBps"] / NOMINAL_HBM_GBS}
    if setup:
        out["mg_setup"] = {k: v for k, v in setup.items() if k != "how"}
    if mg is not None:
        mg.destroy()
    for o in (gcr, A):
        o.destroy()
    return out


def c1_sample_workload(host, torch, ctx, local_rank):
    """BASELINE configs[0] on the GPU: the shipped 4^4 sample operator through Sparse (the arrays read_data parses) + DiracOp,
    GCR_Param(0,5,4000,1e-13); beside it the UNMODIFIED reference's own solve of the same problem on this box's CPU."""
    import ctypes as C
    import numpy as np
    from mgpreconditionedgcr_b200 import capi
    m = np.load(os.path.join(ROOT, "tests", "golden", "c1_matrix.npz"))
    k = 0.05 + 8 * ((0.17865 - 0.05) / 10.)
    A = host.DiracOp(ctx, host.Sparse(ctx, 3072, 3072, m["row"].astype(np.int64), m["col"].astype(np.int64), m["val"]), k)
    rhs = ctx.init_rand(0, 3072)
    x = ctx.field(3072)
    prm = host.GCR_Param(0, 5, 4000, 1e-13, False, None, None)
    gcr = host.GCR(ctx, A, prm)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    for _ in range(3):
        x.set_zero(); it, hist = gcr.solve(rhs, x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync()
    e0.record(stream)
    for _ in range(10):
        x.set_zero(); it, hist = gcr.solve(rhs, x, hist_cap=2)
    e1.record(stream)
    ctx.sync(); torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e4
    x.set_zero(); it, hist = gcr.solve(rhs, x)
    # end to end through host buffers (mgcr_gcr_solve_host), wall clock, best of 5
    h_rhs, h_x = rhs.numpy(), np.zeros(3072, dtype=np.complex128)
    hh, itc, e2e = np.zeros(2), C.c_int(), []
    for _ in range(5):
        h_x[:] = 0
        t0 = time.perf_counter()
        capi.check(ctx.lib.mgcr_gcr_solve_host(ctx.h, A.h, C.byref(prm), None, None, capi.ptr(h_rhs), capi.ptr(h_x), capi.ptr(hh), 2, C.byref(itc)))
        e2e.append(time.perf_counter() - t0)
    out = {"desc": "data/sample_matrix 4x4parsed (3072 rows, 39/row), DiracOp(k=0.15292), GCR_Param(0,5,4000,1e-13), rhs = init_rand(0), x0 = 0 "
                   "(src/main.cpp:834-875 on the 4^4 data); whole solve in one persistent kernel",
           "rows": 3072, "value": sec, "unit": "s", "e2e": min(e2e), "iterations": it, "final_rel_residual": float(hist[-1])}
    ref = c1_sample_reference()
    if ref is not None:
        mm = min(len(hist), len(ref["hist"]))
        out["cpu_reference"] = {"value": ref["seconds"], "unit": "s", "kind": "reference", "cores": 1, "iterations": ref["iters"],
                                "sample": "the UNMODIFIED reference (oracle/_ref/ref_oracle gcr-file: src/GCR.h + src/Operator.h compiled in place) solving the same problem, whole solve, no extrapolation"}
        out["parity"] = {"iters_gpu": it, "iters_reference": ref["iters"], "max_hist_rel_first_40": float(np.max(np.abs(hist[:40] - ref["hist"][:40]) / ref["hist"][:40])),
                         "max_hist_rel": float(np.max(np.abs(hist[:mm] - ref["hist"][:mm]) / ref["hist"][:mm])),
                         "x_rel": float(np.linalg.norm(x.numpy() - ref["x"]) / np.linalg.norm(ref["x"]))}
        out["speedup_vs_reference_e2e"] = ref["seconds"] / min(e2e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--operator", default="csr", choices=["csr", "stencil"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the secondary workloads of the default run")
    ap.add_argument("--max-iter", type=int, default=None, help="cap GCR iterations (profiling runs)")
    ap.add_argument("--mg", default=None, help='JSON overriding the workload\'s MG parameters, e.g. \'{"coarse": [0,10,4,0.1], "n_eigen": [8,8]}\'')
    ap.add_argument("--restart", type=int, default=None, help="override the outer GCR restart length")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    default_run = args.workload is None
    if args.workload is None:
        # BASELINE.json's metric is the MG-GCR time-to-solution; its target configuration is the 512^3 4-level solve (configs[3]),
        # which fits one B200 matrix-free, so the same workload runs at every N (strong scaling).  The default run then adds one
        # solve each of configs[0] / [1] / [2] under "other_workloads" ([4] at 8 GPUs).
        args.workload = "mg3d_512"
    wl = dict(WORKLOADS[args.workload])
    if args.max_iter:
        wl["max_iter"] = args.max_iter
    if args.restart:
        wl["restart"] = args.restart
    if args.mg and wl.get("mg"):
        wl["mg"] = dict(wl["mg"], **json.loads(args.mg))
    if args.impl == "reference":
        return run_reference(args, wl, rank)

    import numpy as np
    import torch
    from mgpreconditionedgcr_b200 import host

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    ctx = host.Context(local_rank)

    def set_align(w):
        # no aggregate may straddle two GPUs: slabs are multiples of the product of the aggregate sizes (as far as every
        # rank still gets a slab); deeper levels are gathered
        align = 1
        if w.get("mg"):
            for sub in w["mg"]["subs"]:
                sub0 = sub if isinstance(sub, int) else sub[0]
                if w["dims"][0] // (align * sub0) >= world:
                    align *= sub0
        ctx.set_slab_align(align)

    if world > 1:
        ids = [host.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.init_dist(rank, world, ids[0])
        set_align(wl)
    if wl.get("aniso"):
        args.operator = "stencil"   # configs[4] is defined matrix-free (its CSR would be 45 GB)
    if wl.get("mg") and args.operator == "csr" and args.workload == "mg3d_512":
        args.operator = "stencil"   # the stored 512^3 operator (22.5 GB) plus the hierarchy is built matrix-free by default
    if world > 1 and args.operator == "csr":
        args.operator = "stencil"   # the distributed benchmark path is matrix-free
    A, rhs, x, V = build_problem(host, torch, ctx, wl, args.operator, world, rank, local_rank)
    n_local = A.get_dim()
    mg, setup = None, None
    if wl.get("mg"):
        mg, setup = build_mg(host, ctx, A, wl)
    param = host.GCR_Param(0, wl["restart"], wl["max_iter"], wl["tol"], False, None, mg)
    gcr = host.GCR(ctx, A, param)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    def step():
        x.set_zero()
        return gcr.solve(rhs, x, hist_cap=2)

    def sync_all():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    it = 0
    for _ in range(args.warmup):
        it, _ = step()
    sync_all()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for _ in range(args.steps):
        it, _ = step()
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    # one more step of the same solve with every launch bracketed by pooled CUDA events on the library's stream (no
    # synchronisation inside the loop): per-kernel-class device time and algorithmic bytes for the roofline.  Kept out
    # of the timed steps because two extra event records per launch slow the launch-bound coarse-level solves down.
    ctx.set_profile(True)
    step()
    sync_all()
    prof = ctx.profile()
    ctx.set_profile(False)
    clk = clocks.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sec_per_solve = ms / args.steps / 1e3
    # true residual of the last solve
    r = rhs - A(x)
    final_rel = r.norm() / rhs.norm()
    del r

    # end to end through host buffers: every rank hands the C ABI's host entry point its slab of rhs / x0 in pinned host
    # memory and gets its slab of x back (mgcr_gcr_solve_host: H2D, solve, D2H inside the timed call); max over ranks
    h_rhs = torch.empty(n_local, dtype=torch.complex128, pin_memory=True)
    h_x = torch.zeros(n_local, dtype=torch.complex128, pin_memory=True)
    h_rhs.numpy()[:] = rhs.numpy()
    import ctypes as C
    from mgpreconditionedgcr_b200 import capi
    hist = np.zeros(2)
    itc = C.c_int()
    times = []
    for i in range(2):
        h_x.zero_()
        sync_all()
        t0 = time.perf_counter()
        capi.check(ctx.lib.mgcr_gcr_solve_host(ctx.h, A.h, C.byref(param), None, mg.h if mg else None, C.c_void_p(h_rhs.data_ptr()), C.c_void_p(h_x.data_ptr()),
                                               capi.ptr(hist), 2, C.byref(itc)))
        times.append(time.perf_counter() - t0)
    e2e_s = min(times)
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": 2 * 16 * V, "d2h_bytes_per_step": 16 * V,
           "how": "mgcr_gcr_solve_host on every rank: pinned host rhs + x0 slab -> HBM, solve, x slab -> host; wall clock around the (synchronous) call, best of 2, max over ranks"}
    del h_rhs, h_x
    if setup is not None and dist is not None:   # the slowest rank's set-up
        keys = [k for k in setup if k != "how"]
        t = torch.tensor([setup[k] for k in keys], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for k, v in zip(keys, t.tolist()):
            setup[k] = v

    peak, peak_src = peaks()
    line = None
    if rank == 0:
        classes, dom, host_side = kernel_classes(prof)
        d = prof[dom]
        achieved = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        # DRAM bytes of one launch of the dominant kernel from the committed ncu --set full capture: only when that capture is of
        # this workload, this kernel class and this GPU count (per-launch sizes differ otherwise)
        traffic, traffic_src, traffic_alg = None, None, None
        tf = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tf):
            ent = json.load(open(tf)).get(args.workload, {})
            if ent.get("n_gpus", 1) == world and dom in ent:
                traffic, traffic_src = ent[dom], ent.get("_source")
                traffic_alg = ent.get("_algorithmic", {}).get(dom)   # algorithmic bytes of the SAME (largest) launch the capture is of
        opname = [n for n in prof if n.startswith(("sell", "hopping"))]
        spmv = None
        if opname:
            o = prof[opname[0]]
            g = o["bytes"] / (o["ms"] * 1e-3) / 1e9
            spmv = {"kernel": opname[0], "GBps": g, "frac": g / peak, "frac_nominal": g / NOMINAL_HBM_GBS, "bytes_per_apply": o["bytes"] / max(o["calls"], 1)}
        line = {
            "metric": METRIC, "value": sec_per_solve, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "c128 (f64)",
            "data": "synthetic",
            "config": {"workload": args.workload, "desc": wl["desc"], "operator": args.operator, "rows": V,
                       "l2": "inputs larger than L2 (every vector is %.0f MB, L2 is 126 MB)" % (16 * V / 1e6)},
            "iterations": it, "final_true_rel_residual": final_rel, "mg_setup_seconds": setup["total_s"] if setup else None, "mg_setup": setup,
            "gpu_launches": launches, "clocks": clk, "e2e": e2e,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_nominal": achieved / NOMINAL_HBM_GBS, "peak_nominal": NOMINAL_HBM_GBS,
                         "traffic": traffic, "traffic_algorithmic": traffic_alg, "traffic_source": traffic_src, "peak_source": peak_src,
                         "how": "algorithmic bytes of every launch of the class / its summed CUDA-event time (events on the library stream) over one more identical step after the timed ones"},
            "spmv": spmv, "kernels": classes,
            "host_side": {kname: {"ms": v["ms"], "calls": v["calls"]} for kname, v in host_side.items()},
            "host": {"nproc": os.cpu_count()},
        }
    # ---- everything below runs outside the timed region, after the headline objects have been given back
    if mg is not None:
        mg.destroy()
    gcr.destroy(); A.destroy()
    del rhs, x
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        s = reference_cpu_sample(wl, host, ctx)
        line["cpu_baseline"] = {"value": extrapolate(s, wl, it), "unit": "s", "cores": s["cores"], "kind": s["kind"], "host_nproc": os.cpu_count(),
                                "sample": s["sample"] + "; extrapolated as sec/iteration x (rows/sample rows) x %d iterations (the count this solve took)" % it,
                                "sec_per_iter_sample": s["sec_per_iter"]}
        if s.get("parity"):
            line["parity"] = s["parity"]
    if default_run and not args.no_others:
        others = {}
        if world == 1:
            for name, op in (("mg3d_256", "stencil"), ("gcr2d_4096", "csr")):
                try:
                    others[name] = secondary_workload(host, torch, ctx, name, op, local_rank, peak)
                except Exception as ex:   # a secondary line must never take the headline down
                    others[name] = {"error": repr(ex)[:300]}
            try:
                others["c1_sample"] = c1_sample_workload(host, torch, ctx, local_rank)
            except Exception as ex:
                others["c1_sample"] = {"error": repr(ex)[:300]}
        if world == 8:   # BASELINE configs[4]: the anisotropic 1024x512x512 five-level solve is quoted at 8 GPUs
            try:
                set_align(WORKLOADS["mg3d_aniso"])
                others["mg3d_aniso"] = secondary_workload(host, torch, ctx, "mg3d_aniso", "stencil", local_rank, peak, world, rank, dist)
            except Exception as ex:
                others["mg3d_aniso"] = {"error": repr(ex)[:300]}
        if rank == 0:
            line["other_workloads"] = others
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
