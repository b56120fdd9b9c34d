"""ctypes declarations of the C ABI in include/mgcr_b200.h (libmgcr_b200.so).

The library is the product: if it has not been built (or cannot be loaded) importing fails loudly -- there is no
Python or CPU fallback for any entry point.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libmgcr_b200.so")

OK, ERR_CUDA, ERR_ARG, ERR_OOM, ERR_NCCL, ERR_UNSUPPORTED = range(6)
MG_NEG_NEIGHBOUR_BUG = 1
MG_STD_CONJ = 2


class MgcrError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("mgcr status %d: %s" % (status, message))
        self.status = status


class GcrParam(C.Structure):
    """mgcr_gcr_param == GCR_Param<num_type> (reference src/SolverParam.h:22-36)"""
    _fields_ = [("truncation", C.c_int), ("restart", C.c_int), ("max_iter", C.c_int), ("tol", C.c_double),
                ("verbose", C.c_int), ("std_conj", C.c_int), ("zero_guess", C.c_int)]


class LevelCfg(C.Structure):
    _fields_ = [("site_dims", C.c_int64 * 4), ("sub", C.c_int64 * 4), ("n_spin", C.c_int), ("n_col", C.c_int),
                ("n_eigen", C.c_int)]


_vp = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_dbl = C.c_double
_pi64 = C.POINTER(C.c_int64)
_pint = C.POINTER(C.c_int)
_pdbl = C.POINTER(C.c_double)
_pvp = C.POINTER(C.c_void_p)
_pgp = C.POINTER(GcrParam)
_plc = C.POINTER(LevelCfg)

# name -> argtypes  (every function returns int status except the two marked)
SIGNATURES = {
    "mgcr_ctx_create": [_int, _pvp],
    "mgcr_ctx_destroy": [_vp],
    "mgcr_ctx_sync": [_vp],
    "mgcr_ctx_stream": [_vp, _pvp],
    "mgcr_ctx_launch_count": [_vp, _pi64],
    "mgcr_ctx_set_profile": [_vp, _int],
    "mgcr_ctx_set_option": [_vp, C.c_char_p, _i64],
    "mgcr_ctx_get_profile": [_vp, _int, C.POINTER(C.c_char_p), _pdbl, _pi64, _pdbl, _pint],
    "mgcr_nccl_unique_id": [_vp],
    "mgcr_ctx_init_dist": [_vp, _int, _int, _vp],
    "mgcr_ctx_rank": [_vp, _pint, _pint],
    "mgcr_ctx_set_slab_align": [_vp, _i64],
    "mgcr_allreduce_sum": [_vp, _vp, _int],
    "mgcr_slab_range": [_i64, _i64, _int, _int, _pi64, _pi64],
    "mgcr_vec_alloc": [_vp, _i64, _pvp],
    "mgcr_vec_free": [_vp, _vp],
    "mgcr_vec_upload": [_vp, _vp, _vp, _i64],
    "mgcr_vec_download": [_vp, _vp, _vp, _i64],
    "mgcr_vec_copy": [_vp, _i64, _vp, _vp],
    "mgcr_vec_set_constant": [_vp, _i64, _dbl, _dbl, _vp],
    "mgcr_vec_axpy": [_vp, _i64, _dbl, _dbl, _vp, _vp, _vp],
    "mgcr_vec_scale": [_vp, _i64, _dbl, _dbl, _vp, _vp],
    "mgcr_vec_dot": [_vp, _i64, _vp, _vp, _pdbl],
    "mgcr_vec_squarednorm": [_vp, _i64, _vp, _pdbl],
    "mgcr_vec_dot_local": [_vp, _i64, _vp, _vp, _pdbl],
    "mgcr_vec_squarednorm_local": [_vp, _i64, _vp, _pdbl],
    "mgcr_vec_normalise": [_vp, _i64, _vp],
    "mgcr_vec_gamma5": [_vp, _int, _pi64, _int, _vp, _vp],
    "mgcr_vec_init_rand": [_vp, _int, _i64, _vp],
    "mgcr_vec_init_rand_slab": [_vp, _int, _i64, _i64, _vp],
    "mgcr_rand_stream": [_int, _i64, _i64, _vp],
    "mgcr_blocking_build": [_vp, _int, _pi64, _pi64, C.POINTER(C.c_uint8), _pi64, _pi64, _pi64],
    "mgcr_csr_create": [_vp, _i64, _i64, _vp, _vp, _vp, _pvp],
    "mgcr_csr_create_dist": [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _pvp],
    "mgcr_hopping_create": [_vp, _int, _pi64, _vp, _pvp],
    "mgcr_hopping_create_dev": [_vp, _int, _pi64, _vp, _pvp],
    "mgcr_dirac_create": [_vp, _vp, _dbl, _dbl, _vp, _pvp],
    "mgcr_dirac_create_dev": [_vp, _vp, _dbl, _dbl, _vp, _pvp],
    "mgcr_dirac_set_k": [_vp, _dbl, _dbl],
    "mgcr_blockcsr_create": [_vp, _i64, _int, _vp, _vp, _vp, _pvp],
    "mgcr_callback_op_create": [_vp, _i64, _vp, _vp, _pvp],
    "mgcr_op_apply": [_vp, _vp, _vp, _vp],
    "mgcr_op_residual": [_vp, _vp, _vp, _vp, _vp],
    "mgcr_op_dim": [_vp, _pi64, _pi64],
    "mgcr_op_apply_bytes": [_vp, _pdbl],
    "mgcr_op_destroy": [_vp],
    "mgcr_gcr_solve": [_vp, _vp, _pgp, _vp, _vp, _vp, _vp, _vp, _int, _pint],
    "mgcr_gcr_solve_host": [_vp, _vp, _pgp, _vp, _vp, _vp, _vp, _vp, _int, _pint],
    "mgcr_gcr_op_create": [_vp, _vp, _pgp, _vp, _vp, _pvp],
    "mgcr_gcr_op_retarget": [_vp, _vp],
    "mgcr_arnoldi": [_vp, _vp, _pgp, _int, _vp],
    "mgcr_mg_create": [_vp, _vp, _int, _plc, _pgp, _pgp, _pgp, _int, _vp, _pvp],
    "mgcr_mg_create_nn": [_vp, _vp, _int, _plc, _pgp, _pgp, _pgp, _int, _vp, _pvp],
    "mgcr_mg_destroy": [_vp],
    "mgcr_mg_setup_profile": [_vp, _int, C.POINTER(C.c_char_p), _pdbl, _pint],
    "mgcr_mg_level_info": [_vp, _int, _pi64, _pi64, _pint, _pi64],
    "mgcr_mg_export_block_map": [_vp, _int, _vp],
    "mgcr_mg_export_prolongator": [_vp, _int, _vp],
    "mgcr_mg_export_coarse": [_vp, _int, _vp, _vp, _vp],
    "mgcr_mg_coarse_op": [_vp, _int, _pvp],
    "mgcr_mg_restrict": [_vp, _vp, _int, _vp, _vp],
    "mgcr_mg_prolong": [_vp, _vp, _int, _vp, _vp],
    "mgcr_mg_cycle": [_vp, _vp, _int, _vp, _vp],
    "mgcr_mg_op_create": [_vp, _vp, _pvp],
}

_LIB = None


def load():
    """Load libmgcr_b200.so; raises (never falls back) when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                          "`make -C mgpreconditionedgcr_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.mgcr_last_error.argtypes = []
    lib.mgcr_last_error.restype = C.c_char_p
    lib.mgcr_abi_version.argtypes = []
    lib.mgcr_abi_version.restype = C.c_int
    _LIB = lib
    return lib


def check(status):
    if status != OK:
        raise MgcrError(status, load().mgcr_last_error().decode("utf-8", "replace"))


def ptr(a):
    """host pointer of a contiguous numpy array (or None)"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)
