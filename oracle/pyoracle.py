"""ctypes front-end of the CPU restatement (oracle/mgcr_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg -- never from the product
package.  Builds oracle/_build/liboracle.so on first use (gcc, a few seconds).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c128 = np.complex128
_cp = np.ctypeslib.ndpointer(dtype=np.complex128, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class GcrParam(C.Structure):
    _fields_ = [("truncation", C.c_int), ("restart", C.c_int), ("max_iter", C.c_int), ("tol", C.c_double),
                ("std_conj", C.c_int)]


class LevelCfg(C.Structure):
    _fields_ = [("site_dims", C.c_long * 4), ("sub", C.c_long * 4), ("n_spin", C.c_int), ("n_col", C.c_int),
                ("n_eigen", C.c_int)]


def build():
    so = os.path.join(HERE, "_build", "liboracle.so")
    src = os.path.join(HERE, "mgcr_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_init_rand.argtypes = [C.c_int, C.c_long, _cp]
        L.orc_set_sum_order.argtypes = [C.c_int]
        L.orc_dot.argtypes = [C.c_long, _cp, _cp, _dp]
        L.orc_squarednorm.argtypes = [C.c_long, _cp]
        L.orc_squarednorm.restype = C.c_double
        L.orc_gamma5.argtypes = [_lp, C.c_int, C.c_int, _cp, _cp]
        L.orc_blocking.argtypes = [_lp, C.c_int, _lp, C.c_char_p, _lp, _lp]
        L.orc_blocking.restype = C.c_long
        L.orc_csr_new.argtypes = [C.c_long, C.c_long, _lp, _lp, _cp]
        L.orc_csr_new.restype = C.c_void_p
        L.orc_dirac_new.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_dirac_new.restype = C.c_void_p
        L.orc_hopping_new.argtypes = [C.c_int, _lp]
        L.orc_hopping_new.restype = C.c_void_p
        L.orc_dirac_diag_new.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.orc_dirac_diag_new.restype = C.c_void_p
        L.orc_hopping_var_new.argtypes = [C.c_int, _lp, C.c_void_p]
        L.orc_hopping_var_new.restype = C.c_void_p
        L.orc_csr_nnz.argtypes = [C.c_void_p]
        L.orc_csr_nnz.restype = C.c_long
        L.orc_csr_export.argtypes = [C.c_void_p, _lp, _lp, _cp]
        L.orc_blockcsr_new.argtypes = [C.c_long, C.c_int, _lp, _lp, _cp]
        L.orc_blockcsr_new.restype = C.c_void_p
        L.orc_op_free.argtypes = [C.c_void_p]
        L.orc_op_dim.argtypes = [C.c_void_p]
        L.orc_op_dim.restype = C.c_long
        L.orc_op_apply.argtypes = [C.c_void_p, _cp, _cp]
        L.orc_gcr_solve.argtypes = [C.c_void_p, C.POINTER(GcrParam), C.c_void_p, _cp, _cp, C.c_void_p, C.c_int]
        L.orc_gcr_solve.restype = C.c_int
        L.orc_gcr_solve_lr.argtypes = [C.c_void_p, C.POINTER(GcrParam), C.c_void_p, C.c_void_p, _cp, _cp, C.c_void_p, C.c_int]
        L.orc_gcr_solve_lr.restype = C.c_int
        L.orc_gcr_op_new.argtypes = [C.c_void_p, C.POINTER(GcrParam), C.c_void_p, C.c_int]
        L.orc_gcr_op_new.restype = C.c_void_p
        L.orc_arnoldi.argtypes = [C.c_void_p, C.POINTER(GcrParam), C.c_int, _cp]
        L.orc_mg_new.argtypes = [C.c_void_p, C.c_int, C.POINTER(LevelCfg), C.POINTER(GcrParam), C.POINTER(GcrParam),
                                 C.POINTER(GcrParam), C.c_int, C.c_int, C.c_void_p]
        L.orc_mg_new.restype = C.c_void_p
        L.orc_mg_new_nn.argtypes = [C.c_void_p, C.c_int, C.POINTER(LevelCfg), C.POINTER(GcrParam), C.POINTER(GcrParam),
                                    C.POINTER(GcrParam), C.c_int, C.c_int, C.c_void_p]
        L.orc_mg_new_nn.restype = C.c_void_p
        L.orc_mg_export_nearnull.argtypes = [C.c_void_p, C.c_int, _cp]
        L.orc_mg_free.argtypes = [C.c_void_p]
        for f in ("orc_mg_nblocks", "orc_mg_block_len"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
            getattr(L, f).restype = C.c_long
        L.orc_mg_ne.argtypes = [C.c_void_p, C.c_int]
        L.orc_mg_ne.restype = C.c_int
        L.orc_mg_export_block_map.argtypes = [C.c_void_p, C.c_int, _lp]
        L.orc_mg_export_prolongator.argtypes = [C.c_void_p, C.c_int, _cp]
        L.orc_mg_export_coarse.argtypes = [C.c_void_p, C.c_int, _lp, _lp, _cp]
        L.orc_mg_coarse_op.argtypes = [C.c_void_p, C.c_int]
        L.orc_mg_coarse_op.restype = C.c_void_p
        L.orc_mg_restrict.argtypes = [C.c_void_p, C.c_int, _cp, _cp]
        L.orc_mg_expand.argtypes = [C.c_void_p, C.c_int, _cp, _cp]
        L.orc_mg_cycle.argtypes = [C.c_void_p, C.c_int, _cp, _cp]
        L.orc_mg_op_new.argtypes = [C.c_void_p]
        L.orc_mg_op_new.restype = C.c_void_p
        _LIB = L
    return _LIB


def _c(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def _l(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def set_sum_order(mode):
    """0 = the reference's left-to-right inner products (default); 1 = right to left; 2 = blocked -- see mgcr_oracle.c"""
    lib().orc_set_sum_order(int(mode))


def init_rand(seed, n):
    out = np.empty(n, dtype=c128)
    lib().orc_init_rand(seed, n, out)
    return out


def dot(a, b):
    out = np.empty(2)
    lib().orc_dot(len(a), _c(a), _c(b), out)
    return complex(out[0], out[1])


def squarednorm(a):
    return lib().orc_squarednorm(len(a), _c(a))


def gamma5(dims, axis, x):
    out = np.empty(len(x), dtype=c128)
    lib().orc_gamma5(_l(dims), len(dims), axis, _c(x), out)
    return out


def blocking(dims, sub4, mask=None):
    dims = _l(dims)
    if mask is None:
        mask = [1, 1, 1, 1] + [0] * (len(dims) - 4)
    nsite = int(np.prod([d for d, m in zip(dims, mask) if m]))
    bm = np.empty(nsite, dtype=np.int64)
    bd = np.empty(4, dtype=np.int64)
    nb = lib().orc_blocking(dims, len(dims), _l(sub4), bytes(bytearray(mask)), bm, bd)
    if nb < 0:
        raise ValueError("dimension not divisible by block size")
    return bm.reshape(nb, -1), bd


class Op:
    """Handle to an oracle operator."""

    def __init__(self, h, keep=()):
        self.h = h
        self.keep = keep
        self.n = lib().orc_op_dim(h)

    def __call__(self, x):
        y = np.empty(self.n, dtype=c128)
        lib().orc_op_apply(self.h, _c(x), y)
        return y


def csr(nrow, ncol, row, col, val):
    return Op(lib().orc_csr_new(nrow, ncol, _l(row), _l(col), _c(val)))


def dirac(D, k, diag=None):
    k = complex(k)
    if diag is not None:
        d = np.ascontiguousarray(diag, dtype=np.float64).reshape(-1)
        return Op(lib().orc_dirac_diag_new(D.h, k.real, k.imag, d.ctypes.data_as(C.c_void_p)), keep=(D,))
    return Op(lib().orc_dirac_new(D.h, k.real, k.imag), keep=(D,))


def hopping(dims, faces=None):
    if faces is not None:
        fs = [np.ascontiguousarray(f, dtype=np.float64).reshape(-1) for f in faces]
        arr = (C.c_void_p * len(fs))(*[f.ctypes.data for f in fs])
        return Op(lib().orc_hopping_var_new(len(dims), _l(dims), arr))
    return Op(lib().orc_hopping_new(len(dims), _l(dims)))


def csr_export(op):
    nnz = lib().orc_csr_nnz(op.h)
    row = np.empty(op.n + 1, dtype=np.int64)
    col = np.empty(nnz, dtype=np.int64)
    val = np.empty(nnz, dtype=c128)
    lib().orc_csr_export(op.h, row, col, val)
    return row, col, val


def blockcsr(nb, ne, brow, bcol, bval):
    return Op(lib().orc_blockcsr_new(nb, ne, _l(brow), _l(bcol), _c(bval)))


def gcr_param(truncation=0, restart=0, max_iter=100, tol=1e-16, std_conj=0):
    return GcrParam(truncation, restart, max_iter, tol, std_conj)


def gcr_solve(A, param, rhs, x0=None, precond=None, alias=False, left=None):
    """Returns (x, hist, iters).  alias=True solves with rhs and x the same buffer (src/MG.h:102).  left: the reference's
    left preconditioner (src/GCR.h:201-204, 245-247)."""
    n = A.n
    cap = param.max_iter + 2
    hist = np.zeros(cap)
    if alias:
        x = _c(rhs).copy()
        rhs_buf = x
    else:
        rhs_buf = _c(rhs)
        x = np.zeros(n, dtype=c128) if x0 is None else _c(x0).copy()
    if left is not None:
        it = lib().orc_gcr_solve_lr(A.h, C.byref(param), left.h, precond.h if precond is not None else None, rhs_buf, x,
                                    hist.ctypes.data_as(C.c_void_p), cap)
        return x, hist[: it + 1].copy(), it
    it = lib().orc_gcr_solve(A.h, C.byref(param), precond.h if precond is not None else None, rhs_buf, x,
                             hist.ctypes.data_as(C.c_void_p), cap)
    return x, hist[: it + 1].copy(), it


def gcr_op(A, param, precond=None, zero_guess=True):
    return Op(lib().orc_gcr_op_new(A.h, C.byref(param), precond.h if precond is not None else None, int(zero_guess)),
              keep=(A, precond, param))


def arnoldi(A, param, n_vec):
    v = np.empty((n_vec, A.n), dtype=c128)
    lib().orc_arnoldi(A.h, C.byref(param), n_vec, v.reshape(-1))
    return v


class MG:
    def __init__(self, A, levels, eigen, coarse, smooth, neg_bug=False, std_conj=False, nearnull=None):
        """levels: list of dict(site_dims=[4], sub=[4], n_spin, n_col, n_eigen)."""
        cfg = (LevelCfg * len(levels))()
        for i, lv in enumerate(levels):
            cfg[i].site_dims[:] = list(lv["site_dims"])
            cfg[i].sub[:] = list(lv["sub"])
            cfg[i].n_spin, cfg[i].n_col, cfg[i].n_eigen = lv.get("n_spin", 1), lv.get("n_col", 1), lv["n_eigen"]
        self.A = A
        self.levels = levels
        self.n_level = len(levels)
        if isinstance(nearnull, (list, tuple)):   # one entry per level (None = the level runs its own inverse iteration)
            self._nn = [None if v is None else _c(np.asarray(v).reshape(-1)) for v in nearnull]
            arr = (C.c_void_p * len(levels))(*[None if v is None else v.ctypes.data for v in self._nn])
            self.h = lib().orc_mg_new_nn(A.h, len(levels), cfg, C.byref(eigen), C.byref(coarse), C.byref(smooth), int(neg_bug), int(std_conj), arr)
            return
        nn = None
        if nearnull is not None:
            self._nn = _c(np.asarray(nearnull).reshape(-1))
            nn = self._nn.ctypes.data_as(C.c_void_p)
        self.h = lib().orc_mg_new(A.h, len(levels), cfg, C.byref(eigen), C.byref(coarse), C.byref(smooth),
                                  int(neg_bug), int(std_conj), nn)

    def nblocks(self, l=0):
        return lib().orc_mg_nblocks(self.h, l)

    def nearnull(self, l=0):
        """the n_eigen near-null vectors level l was built from, shape (n_eigen, n_l)"""
        n = self.A.n if l == 0 else self.nblocks(l - 1) * self.ne(l - 1)
        out = np.empty(self.levels[l]["n_eigen"] * n, dtype=c128)
        lib().orc_mg_export_nearnull(self.h, l, out)
        return out.reshape(self.levels[l]["n_eigen"], n)

    def ne(self, l=0):
        return lib().orc_mg_ne(self.h, l)

    def block_map(self, l=0):
        nb = self.nblocks(l)
        nsite = int(np.prod(self.levels[l]["site_dims"]))
        out = np.empty(nsite, dtype=np.int64)
        lib().orc_mg_export_block_map(self.h, l, out)
        return out.reshape(nb, -1)

    def prolongator(self, l=0):
        nb, ne, bl = self.nblocks(l), self.ne(l), lib().orc_mg_block_len(self.h, l)
        out = np.empty(nb * ne * bl, dtype=c128)
        lib().orc_mg_export_prolongator(self.h, l, out)
        return out.reshape(nb, ne, bl)

    def coarse(self, l=0):
        nb, ne = self.nblocks(l), self.ne(l)
        brow = np.empty(nb + 1, dtype=np.int64)
        bcol = np.empty(9 * nb, dtype=np.int64)
        bval = np.empty(9 * nb * ne * ne, dtype=c128)
        lib().orc_mg_export_coarse(self.h, l, brow, bcol, bval)
        return brow, bcol, bval.reshape(9 * nb, ne, ne)

    def coarse_op(self, l=0):
        return Op(lib().orc_mg_coarse_op(self.h, l), keep=(self,))

    def restrict(self, xf, l=0):
        out = np.empty(self.nblocks(l) * self.ne(l), dtype=c128)
        lib().orc_mg_restrict(self.h, l, _c(xf), out)
        return out

    def expand(self, xc, l=0):
        n = self.A.n if l == 0 else self.nblocks(l - 1) * self.ne(l - 1)
        out = np.empty(n, dtype=c128)
        lib().orc_mg_expand(self.h, l, _c(xc), out)
        return out

    def cycle(self, b, l=0):
        out = np.empty(len(b), dtype=c128)
        lib().orc_mg_cycle(self.h, l, _c(b), out)
        return out

    def as_op(self):
        return Op(lib().orc_mg_op_new(self.h), keep=(self,))
