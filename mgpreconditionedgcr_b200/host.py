"""Python mirror of the reference's host interface above the C ABI (used by tests/ and bench.py).

Names follow the reference headers (src/Fields.h, src/Operator.h, src/GCR.h, src/MG.h, src/SolverParam.h):
Field, Sparse, DiracOp, HierarchicalSparse, GCR_Param, GCR, MG_Param, MG.  Every method is one call into
libmgcr_b200.so; nothing is computed here.  The C++ drop-in headers (include/mgcr/) wrap the same entry points.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import GcrParam, LevelCfg, check


class Context:
    """mgcr_ctx: one GPU, one stream."""

    def __init__(self, device=0):
        self.lib = capi.load()
        h = C.c_void_p()
        check(self.lib.mgcr_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        self.slab_align = 1

    def close(self):
        if self.h:
            self.lib.mgcr_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        check(self.lib.mgcr_ctx_sync(self.h))

    @property
    def stream(self):
        s = C.c_void_p()
        check(self.lib.mgcr_ctx_stream(self.h, C.byref(s)))
        return s.value or 0

    @property
    def launches(self):
        n = C.c_int64()
        check(self.lib.mgcr_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def set_profile(self, on):
        check(self.lib.mgcr_ctx_set_profile(self.h, int(on)))

    def profile(self):
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        calls = (C.c_int64 * cap)()
        nbytes = (C.c_double * cap)()
        n = C.c_int()
        check(self.lib.mgcr_ctx_get_profile(self.h, cap, names, ms, calls, nbytes, C.byref(n)))
        return {names[i].decode(): dict(ms=ms[i], calls=calls[i], bytes=nbytes[i]) for i in range(min(n.value, cap))}

    def init_dist(self, rank, nranks, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        check(self.lib.mgcr_ctx_init_dist(self.h, rank, nranks, buf))

    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        check(capi.load().mgcr_nccl_unique_id(buf))
        return bytes(buf.raw)

    def set_option(self, key, value):
        check(self.lib.mgcr_ctx_set_option(self.h, key.encode(), int(value)))

    def set_slab_align(self, align):
        """slab boundaries of distributed operators created afterwards are multiples of `align` planes"""
        check(self.lib.mgcr_ctx_set_slab_align(self.h, int(align)))
        self.slab_align = int(align)

    def rank(self):
        r, n = C.c_int(), C.c_int()
        check(self.lib.mgcr_ctx_rank(self.h, C.byref(r), C.byref(n)))
        return r.value, n.value

    # ---- Field factory helpers
    def field(self, n):
        return Field(self, n)

    def from_numpy(self, a):
        a = capi.c128(a).reshape(-1)
        f = Field(self, a.size)
        check(self.lib.mgcr_vec_upload(self.h, f.ptr, capi.ptr(a), a.size))
        return f

    def init_rand(self, seed, n, skip=0):
        f = Field(self, n)
        check(self.lib.mgcr_vec_init_rand_slab(self.h, seed, skip, n, f.ptr))
        return f

    def blocking(self, dims, sub4, mask=None):
        dims = capi.i64(dims)
        if mask is None:
            mask = [1, 1, 1, 1] + [0] * (len(dims) - 4)
        nsite = int(np.prod([d for d, m in zip(dims, mask) if m]))
        bm = np.empty(nsite, dtype=np.int64)
        bd = np.empty(4, dtype=np.int64)
        nb = C.c_int64()
        m = (C.c_uint8 * len(mask))(*mask)
        sub = capi.i64(sub4)
        check(self.lib.mgcr_blocking_build(self.h, len(dims), dims.ctypes.data_as(C.POINTER(C.c_int64)),
                                           sub.ctypes.data_as(C.POINTER(C.c_int64)), m,
                                           bm.ctypes.data_as(C.POINTER(C.c_int64)), bd.ctypes.data_as(C.POINTER(C.c_int64)),
                                           C.byref(nb)))
        return bm.reshape(nb.value, -1), bd


def slab_range(n, align, rank, nranks):
    b, e = C.c_int64(), C.c_int64()
    check(capi.load().mgcr_slab_range(n, align, rank, nranks, C.byref(b), C.byref(e)))
    return b.value, e.value


class Field:
    """Field<num_type> (src/Fields.h:29-71): an owning device array of complex128."""

    def __init__(self, ctx, n, ptr=None):
        self.ctx = ctx
        self.n = int(n)
        self.owned = ptr is None
        if ptr is None:
            p = C.c_void_p()
            check(ctx.lib.mgcr_vec_alloc(ctx.h, self.n, C.byref(p)))
            ptr = p
        self.ptr = ptr

    def free(self):
        if self.owned and self.ptr is not None and self.ctx.h:
            self.ctx.lib.mgcr_vec_free(self.ctx.h, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def numpy(self):
        out = np.empty(self.n, dtype=np.complex128)
        check(self.ctx.lib.mgcr_vec_download(self.ctx.h, capi.ptr(out), self.ptr, self.n))
        return out

    def upload(self, a):
        a = capi.c128(a).reshape(-1)
        assert a.size == self.n
        check(self.ctx.lib.mgcr_vec_upload(self.ctx.h, self.ptr, capi.ptr(a), self.n))
        return self

    def view(self, offset, n):
        return Field(self.ctx, n, ptr=C.c_void_p(self.ptr.value + 16 * offset))

    def copy(self):
        f = Field(self.ctx, self.n)
        check(self.ctx.lib.mgcr_vec_copy(self.ctx.h, self.n, self.ptr, f.ptr))
        return f

    def set_zero(self):
        check(self.ctx.lib.mgcr_vec_set_constant(self.ctx.h, self.n, 0., 0., self.ptr))
        return self

    def set_constant(self, c):
        c = complex(c)
        check(self.ctx.lib.mgcr_vec_set_constant(self.ctx.h, self.n, c.real, c.imag, self.ptr))
        return self

    def _axpy(self, s, b, out):
        s = complex(s)
        check(self.ctx.lib.mgcr_vec_axpy(self.ctx.h, self.n, s.real, s.imag, b.ptr, self.ptr, out.ptr))
        return out

    def __add__(self, b):
        return self._axpy(1., b, Field(self.ctx, self.n))

    def __sub__(self, b):
        return self._axpy(-1., b, Field(self.ctx, self.n))

    def __iadd__(self, b):
        return self._axpy(1., b, self)

    def __isub__(self, b):
        return self._axpy(-1., b, self)

    def __mul__(self, s):
        s = complex(s)
        out = Field(self.ctx, self.n)
        check(self.ctx.lib.mgcr_vec_scale(self.ctx.h, self.n, s.real, s.imag, self.ptr, out.ptr))
        return out

    def dot(self, b):
        out = (C.c_double * 2)()
        check(self.ctx.lib.mgcr_vec_dot(self.ctx.h, self.n, self.ptr, b.ptr, out))
        return complex(out[0], out[1])

    def squarednorm(self):
        out = C.c_double()
        check(self.ctx.lib.mgcr_vec_squarednorm(self.ctx.h, self.n, self.ptr, C.byref(out)))
        return out.value

    def norm(self):
        return float(np.sqrt(self.squarednorm()))

    def normalise(self):
        check(self.ctx.lib.mgcr_vec_normalise(self.ctx.h, self.n, self.ptr))
        return self

    def gamma5(self, dims, axis):
        dims = capi.i64(dims)
        out = Field(self.ctx, self.n)
        check(self.ctx.lib.mgcr_vec_gamma5(self.ctx.h, len(dims), dims.ctypes.data_as(C.POINTER(C.c_int64)), axis, self.ptr, out.ptr))
        return out


class Operator:
    """Operator<num_type> (src/Operator.h:16-29)."""

    def __init__(self, ctx, h, keep=()):
        self.ctx = ctx
        self.h = h
        self.keep = keep
        self.owned = True

    def get_dim(self):
        nl, ng = C.c_int64(), C.c_int64()
        check(self.ctx.lib.mgcr_op_dim(self.h, C.byref(nl), C.byref(ng)))
        return nl.value

    @property
    def n(self):
        return self.get_dim()

    def global_dim(self):
        nl, ng = C.c_int64(), C.c_int64()
        check(self.ctx.lib.mgcr_op_dim(self.h, C.byref(nl), C.byref(ng)))
        return ng.value

    def apply_bytes(self):
        b = C.c_double()
        check(self.ctx.lib.mgcr_op_apply_bytes(self.h, C.byref(b)))
        return b.value

    def __call__(self, f, out=None):
        if isinstance(f, np.ndarray):
            fin = self.ctx.from_numpy(f)
            res = Field(self.ctx, self.get_dim())
            check(self.ctx.lib.mgcr_op_apply(self.ctx.h, self.h, fin.ptr, res.ptr))
            return res.numpy()
        if out is None:
            out = Field(self.ctx, self.get_dim())
        check(self.ctx.lib.mgcr_op_apply(self.ctx.h, self.h, f.ptr, out.ptr))
        return out

    def residual(self, x, b, out=None):
        """b - A x in one pass (numpy in -> numpy out, Field in -> Field out)"""
        if isinstance(x, np.ndarray):
            fx, fb, res = self.ctx.from_numpy(x), self.ctx.from_numpy(b), Field(self.ctx, self.get_dim())   # all three alive during the call
            check(self.ctx.lib.mgcr_op_residual(self.ctx.h, self.h, fx.ptr, fb.ptr, res.ptr))
            return res.numpy()
        if out is None:
            out = Field(self.ctx, self.get_dim())
        check(self.ctx.lib.mgcr_op_residual(self.ctx.h, self.h, x.ptr, b.ptr, out.ptr))
        return out

    def destroy(self):
        if self.owned and self.h is not None and self.ctx.h:
            self.ctx.lib.mgcr_op_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Sparse(Operator):
    """Sparse<num_type>(rows, cols, ROW, COL, VAL) (src/Operator.h:64)."""

    def __init__(self, ctx, nrow, ncol, row, col, val, row_range=None, nrow_global=None):
        row, col, val = capi.i64(row), capi.i64(col), capi.c128(val)
        h = C.c_void_p()
        if row_range is None:
            check(ctx.lib.mgcr_csr_create(ctx.h, nrow, ncol, capi.ptr(row), capi.ptr(col), capi.ptr(val), C.byref(h)))
        else:
            check(ctx.lib.mgcr_csr_create_dist(ctx.h, nrow_global, row_range[0], row_range[1], capi.ptr(row), capi.ptr(col),
                                               capi.ptr(val), C.byref(h)))
        super().__init__(ctx, h)
        self.nnz = int(row[-1])


def _ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    return arr


class Hopping(Operator):
    """Matrix-free hopping operator of a Dirichlet lattice (what make_hopping + Sparse give, never stored).
    faces: optional list of len(dims) host arrays of n_local doubles, faces[d][i] = coefficient of the bond between site i
    and its +1 neighbour in dim d (unit hopping when omitted).  faces_dev: the same as DEVICE pointers (ints)."""

    def __init__(self, ctx, dims, faces=None, faces_dev=None):
        dims = capi.i64(dims)
        h = C.c_void_p()
        pd = dims.ctypes.data_as(C.POINTER(C.c_int64))
        if faces_dev is not None:
            assert len(faces_dev) == len(dims)
            check(ctx.lib.mgcr_hopping_create_dev(ctx.h, len(dims), pd, _ptr_array([int(p) for p in faces_dev]), C.byref(h)))
        elif faces is not None:
            assert len(faces) == len(dims)
            fs = [np.ascontiguousarray(f, dtype=np.float64).reshape(-1) for f in faces]
            check(ctx.lib.mgcr_hopping_create(ctx.h, len(dims), pd, _ptr_array([f.ctypes.data for f in fs]), C.byref(h)))
        else:
            check(ctx.lib.mgcr_hopping_create(ctx.h, len(dims), pd, None, C.byref(h)))
        super().__init__(ctx, h)
        self.dims = [int(d) for d in dims]


class DiracOp(Operator):
    """DiracOp<num_type>(D, k) = 1 - k D (src/Operator.h:105-122); diag (host array) / diag_dev (device pointer) generalise
    the identity to a real diagonal."""

    def __init__(self, ctx, D, k, diag=None, diag_dev=None):
        k = complex(k)
        h = C.c_void_p()
        if diag_dev is not None:
            check(ctx.lib.mgcr_dirac_create_dev(ctx.h, D.h, k.real, k.imag, C.c_void_p(int(diag_dev)), C.byref(h)))
        else:
            d = None if diag is None else np.ascontiguousarray(diag, dtype=np.float64)
            check(ctx.lib.mgcr_dirac_create(ctx.h, D.h, k.real, k.imag, capi.ptr(d), C.byref(h)))
        super().__init__(ctx, h, keep=(D,))

    def set_k(self, k):
        k = complex(k)
        check(self.ctx.lib.mgcr_dirac_set_k(self.h, k.real, k.imag))


class HierarchicalSparse(Operator):
    """HierarchicalSparse<num_type,int> (src/HierarchicalSparse.h:22-48) from sorted block-CSR arrays."""

    def __init__(self, ctx, nb, ne, brow, bcol, bval):
        brow, bcol, bval = capi.i64(brow), capi.i64(bcol), capi.c128(bval)
        h = C.c_void_p()
        check(ctx.lib.mgcr_blockcsr_create(ctx.h, nb, ne, capi.ptr(brow), capi.ptr(bcol), capi.ptr(bval), C.byref(h)))
        super().__init__(ctx, h)


def GCR_Param(trunc=0, re=0, max_it=100, tau=1e-16, verb=False, solver_l=None, solver_r=None, std_conj=False,
              zero_guess=False):
    """GCR_Param<num_type>(trunc, re, max_it, tau, verb, solver_l, solver_r) (src/SolverParam.h:33)."""
    p = GcrParam(trunc, re, max_it, tau, int(verb), int(std_conj), int(zero_guess))
    p.left_precond = solver_l
    p.right_precond = solver_r
    return p


class GCR(Operator):
    """GCR<num_type>(A, &param) (src/GCR.h:18-50): a solver that is itself an Operator."""

    def __init__(self, ctx, A, param):
        self.A = A
        self.param = param
        left = getattr(param, "left_precond", None)
        right = getattr(param, "right_precond", None)
        h = C.c_void_p()
        check(ctx.lib.mgcr_gcr_op_create(ctx.h, A.h, C.byref(param), left.h if left else None, right.h if right else None, C.byref(h)))
        super().__init__(ctx, h, keep=(A, param, left, right))

    def initialise(self, A):
        check(self.ctx.lib.mgcr_gcr_op_retarget(self.h, A.h))
        self.A = A

    def solve(self, rhs, x, hist_cap=None):
        """GCR::solve(rhs, x) (src/GCR.h:158-302).  Returns (iterations, residual history)."""
        p = self.param
        left = getattr(p, "left_precond", None)
        right = getattr(p, "right_precond", None)
        cap = (p.max_iter + 2) if hist_cap is None else hist_cap
        hist = np.zeros(cap)
        it = C.c_int()
        check(self.ctx.lib.mgcr_gcr_solve(self.ctx.h, self.A.h, C.byref(p), left.h if left else None, right.h if right else None,
                                          rhs.ptr, x.ptr, capi.ptr(hist), cap, C.byref(it)))
        return it.value, hist[: it.value + 1].copy()

    def solve_host(self, rhs, x0):
        """the same through host buffers (numpy in, numpy out)"""
        p = self.param
        right = getattr(p, "right_precond", None)
        left = getattr(p, "left_precond", None)
        rhs = capi.c128(rhs)
        x = capi.c128(x0).copy()
        cap = p.max_iter + 2
        hist = np.zeros(cap)
        it = C.c_int()
        check(self.ctx.lib.mgcr_gcr_solve_host(self.ctx.h, self.A.h, C.byref(p), left.h if left else None, right.h if right else None, capi.ptr(rhs),
                                               capi.ptr(x), capi.ptr(hist), cap, C.byref(it)))
        return x, it.value, hist[: it.value + 1].copy()


def arnoldi(ctx, A, eigen_param, n_vec):
    """Arnoldi::solve (src/MG.h:90-122) -> Field of n_vec * n"""
    n = A.get_dim()
    v = Field(ctx, n_vec * n)
    check(ctx.lib.mgcr_arnoldi(ctx.h, A.h, C.byref(eigen_param), n_vec, v.ptr))
    return v


class MG(Operator):
    """MG<num_type> (src/MG.h:20-61): hierarchy + cycle, usable as a preconditioner Operator.

    levels: list of dict(site_dims=[4], sub=[4], n_spin, n_col, n_eigen) -- the reference's MG_Param mesh /
    subblock_dim / n_eigen, per level and per dimension."""

    def __init__(self, ctx, A, levels, eigen, coarse, smooth, neg_bug=False, std_conj=False, nearnull=None):
        cfg = (LevelCfg * len(levels))()
        for i, lv in enumerate(levels):
            cfg[i].site_dims[:] = list(lv["site_dims"])
            cfg[i].sub[:] = list(lv["sub"])
            cfg[i].n_spin, cfg[i].n_col, cfg[i].n_eigen = lv.get("n_spin", 1), lv.get("n_col", 1), lv["n_eigen"]
        flags = (capi.MG_NEG_NEIGHBOUR_BUG if neg_bug else 0) | (capi.MG_STD_CONJ if std_conj else 0)
        # nearnull: the level-0 vectors (array / Field), or a list with one entry per level (None = own inverse iteration)
        per_level = list(nearnull) if isinstance(nearnull, (list, tuple)) else [nearnull] + [None] * (len(levels) - 1)
        assert len(per_level) == len(levels)
        nn = [None if v is None else (v if isinstance(v, Field) else ctx.from_numpy(np.asarray(v).reshape(-1))) for v in per_level]
        ptrs = (C.c_void_p * len(levels))(*[None if v is None else v.ptr.value for v in nn])
        mg = C.c_void_p()
        check(ctx.lib.mgcr_mg_create_nn(ctx.h, A.h, len(levels), cfg, C.byref(eigen), C.byref(coarse), C.byref(smooth), flags,
                                        ptrs, C.byref(mg)))
        del nn
        self.mg = mg
        self.levels = levels
        h = C.c_void_p()
        check(ctx.lib.mgcr_mg_op_create(ctx.h, mg, C.byref(h)))
        super().__init__(ctx, h, keep=(A, eigen, coarse, smooth))

    def setup_profile(self):
        """wall-clock seconds per set-up stage (summed over the levels)"""
        cap = 32
        names = (C.c_char_p * cap)()
        sec = (C.c_double * cap)()
        n = C.c_int()
        check(self.ctx.lib.mgcr_mg_setup_profile(self.mg, cap, names, sec, C.byref(n)))
        return {names[i].decode(): sec[i] for i in range(min(n.value, cap))}

    def info(self, l=0):
        nf, nb, bl = C.c_int64(), C.c_int64(), C.c_int64()
        ne = C.c_int()
        check(self.ctx.lib.mgcr_mg_level_info(self.mg, l, C.byref(nf), C.byref(nb), C.byref(ne), C.byref(bl)))
        return dict(n_fine=nf.value, n_blocks=nb.value, ne=ne.value, block_len=bl.value)

    def block_map(self, l=0):
        i = self.info(l)
        nsite = int(np.prod(self.levels[l]["site_dims"]))
        out = np.empty(nsite, dtype=np.int64)
        check(self.ctx.lib.mgcr_mg_export_block_map(self.mg, l, capi.ptr(out)))
        return out.reshape(i["n_blocks"], -1)

    def prolongator(self, l=0):
        i = self.info(l)
        out = np.empty(i["n_blocks"] * i["ne"] * i["block_len"], dtype=np.complex128)
        check(self.ctx.lib.mgcr_mg_export_prolongator(self.mg, l, capi.ptr(out)))
        return out.reshape(i["n_blocks"], i["ne"], i["block_len"])

    def coarse(self, l=0):
        i = self.info(l)
        nb, ne = i["n_blocks"], i["ne"]
        brow = np.empty(nb + 1, dtype=np.int64)
        bcol = np.empty(9 * nb, dtype=np.int64)
        bval = np.empty(9 * nb * ne * ne, dtype=np.complex128)
        check(self.ctx.lib.mgcr_mg_export_coarse(self.mg, l, capi.ptr(brow), capi.ptr(bcol), capi.ptr(bval)))
        return brow, bcol, bval.reshape(9 * nb, ne, ne)

    def coarse_op(self, l=0):
        h = C.c_void_p()
        check(self.ctx.lib.mgcr_mg_coarse_op(self.mg, l, C.byref(h)))
        op = Operator(self.ctx, h, keep=(self,))
        op.owned = False
        return op

    def restrict(self, xf, l=0):
        i = self.info(l)
        out = Field(self.ctx, i["n_blocks"] * i["ne"])
        check(self.ctx.lib.mgcr_mg_restrict(self.ctx.h, self.mg, l, xf.ptr, out.ptr))
        return out

    def expand(self, xc, l=0):
        i = self.info(l)
        out = Field(self.ctx, i["n_fine"])
        check(self.ctx.lib.mgcr_mg_prolong(self.ctx.h, self.mg, l, xc.ptr, out.ptr))
        return out

    def cycle(self, b, l=0):
        out = Field(self.ctx, b.n)
        check(self.ctx.lib.mgcr_mg_cycle(self.ctx.h, self.mg, l, b.ptr, out.ptr))
        return out

    def destroy(self):
        super().destroy()
        if getattr(self, "mg", None) is not None and self.ctx.h:
            self.ctx.lib.mgcr_mg_destroy(self.mg)
        self.mg = None


def hopping_csr(dims, faces=None):
    """Host CSR of the hopping matrix of a Dirichlet lattice in the reference's layout (ascending columns): the synthetic
    operator of SURVEY.md 8(d), for feeding Sparse(...) exactly as a user of the reference would.  faces (optional): per-dim
    bond coefficients as in Hopping(...); unit hopping when omitted."""
    dims = [int(d) for d in dims]
    nd = len(dims)
    V = int(np.prod(dims))
    idx = np.arange(V, dtype=np.int64)
    stride = [int(np.prod(dims[d + 1:])) for d in range(nd)]
    coords = [(idx // stride[d]) % dims[d] for d in range(nd)]
    cols, mask, vals = [], [], []
    for d in range(nd):
        cols.append(idx - stride[d])
        mask.append(coords[d] > 0)
        if faces is not None:   # bond between (i - stride) and i lives with the lower site
            vals.append(np.asarray(faces[d], dtype=np.float64).reshape(-1)[np.maximum(idx - stride[d], 0)])
    for d in range(nd - 1, -1, -1):
        cols.append(idx + stride[d])
        mask.append(coords[d] < dims[d] - 1)
        if faces is not None:
            vals.append(np.asarray(faces[d], dtype=np.float64).reshape(-1))
    cols = np.stack(cols, axis=1)
    mask = np.stack(mask, axis=1)
    counts = mask.sum(axis=1)
    row = np.zeros(V + 1, dtype=np.int64)
    np.cumsum(counts, out=row[1:])
    col = cols[mask]
    if faces is None:
        val = np.ones(col.size, dtype=np.complex128)
    else:
        val = np.stack(vals, axis=1)[mask].astype(np.complex128)
    return row, col.astype(np.int64), val


# ----------------------------------------------------------------------------------------------------------------------
# Synthetic anisotropic variable-coefficient operator of BASELINE.json configs[4]:
#   a_d(site) = exp(sigma * g_d(site)), g_d in [-1, 1) from a counter-based hash of (global site, d, seed) -- no rand()
#   stream, so any slab of the lattice can be generated independently on the host or on the device;
#   bond between site i and i+e_d: eps_d * (a_d(i) + a_d(i+e_d)) / 2; diagonal: sum of the bonds of the site + m2;
#   A = diag - H  (DiracOp(H, 1, diag)), symmetric positive definite.
# ----------------------------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def hash_unit(site, d, seed):
    """splitmix64 finaliser of (3*site + d) + seed*golden -> double in [-1, 1).  site: int64/uint64 numpy array."""
    with np.errstate(over="ignore"):
        z = site.astype(np.uint64) * np.uint64(3) + np.uint64(d) + np.uint64((seed * 0x9E3779B97F4A7C15) & _M64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0


def synthetic_bonds(dims, eps=(1e-4, 1e-2, 1.0), sigma=0.5, m2=0.01, seed=12345, z_range=None):
    """Bond arrays (dims order: slowest first) and diagonal of the C5 operator for planes z_range of the slowest dim.
    eps is in dims order (the fastest dim carries the strong coupling).  Returns (faces[3], diag), each (nz, ny, nx)."""
    nzg, ny, nx = [int(d) for d in dims]
    z0, z1 = (0, nzg) if z_range is None else z_range
    # planes z0-1 .. z1 (clamped) so that the bonds to the neighbour slabs are known
    za, zb = max(z0 - 1, 0), min(z1 + 1, nzg)
    zz = np.arange(za, zb, dtype=np.int64)[:, None, None]
    site = (zz * ny + np.arange(ny, dtype=np.int64)[None, :, None]) * nx + np.arange(nx, dtype=np.int64)[None, None, :]
    stride = (ny * nx, nx, 1)
    faces_ext = []
    for d in range(3):
        a = np.exp(sigma * hash_unit(site, d, seed))
        # coefficient at the +1 neighbour: recompute from the hash (the neighbour may lie outside the generated planes)
        an = np.exp(sigma * hash_unit(site + stride[d], d, seed))
        f = eps[d] * ((a + an) * 0.5)
        # no bond leaves the lattice
        if d == 0:
            f[zz[:, 0, 0] == nzg - 1, :, :] = 0.0
        elif d == 1:
            f[:, ny - 1, :] = 0.0
        else:
            f[:, :, nx - 1] = 0.0
        faces_ext.append(f)
    lo = z0 - za
    n = z1 - z0
    diag = np.full((n, ny, nx), m2, dtype=np.float64)
    fz, fy, fx = faces_ext
    diag += fz[lo:lo + n]
    if z0 > 0:
        diag += fz[lo - 1:lo - 1 + n]
    else:
        diag[1:] += fz[lo:lo + n - 1]
    diag += fy[lo:lo + n]
    diag[:, 1:, :] += fy[lo:lo + n, :-1, :]
    diag += fx[lo:lo + n]
    diag[:, :, 1:] += fx[lo:lo + n, :, :-1]
    faces = [np.ascontiguousarray(f[lo:lo + n]) for f in faces_ext]
    return faces, diag
