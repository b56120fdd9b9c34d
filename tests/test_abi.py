"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/mgcr_b200.h
declares, the ctypes table matches the header, and the product fails loudly (no fallback) without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "mgcr_b200.h")


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "mgpreconditionedgcr_b200", "csrc"), "-j8"],
                          stdout=subprocess.DEVNULL)
    from mgpreconditionedgcr_b200 import capi
    return capi.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(mgcr_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        out[m.group(1)] = args
    return out


def test_header_declares_a_complete_boundary():
    fns = declared_functions()
    for must in ("mgcr_ctx_create", "mgcr_csr_create", "mgcr_dirac_create", "mgcr_op_apply", "mgcr_vec_dot", "mgcr_gcr_solve",
                 "mgcr_gcr_solve_host", "mgcr_blocking_build", "mgcr_mg_create", "mgcr_mg_restrict", "mgcr_mg_prolong",
                 "mgcr_blockcsr_create", "mgcr_ctx_init_dist", "mgcr_allreduce_sum"):
        assert must in fns
    # every entry point cites the reference interface it replaces or states that there is none
    text = open(HEADER).read()
    for cite in ("Operator.h:64", "GCR.h:158-302", "MG.h:131", "Mesh.h:236-298", "Fields.h:216-226", "HierarchicalSparse.h:58-98"):
        assert cite in text


def test_library_exports_every_declared_symbol(lib):
    fns = declared_functions()
    assert len(fns) >= 50
    for name in fns:
        assert hasattr(lib, name), "libmgcr_b200.so does not export %s" % name


def test_ctypes_table_matches_header(lib):
    from mgpreconditionedgcr_b200 import capi
    fns = declared_functions()
    for name, args in fns.items():
        if name in ("mgcr_last_error", "mgcr_abi_version"):
            continue
        assert name in capi.SIGNATURES, "capi.py has no signature for %s" % name
        assert len(capi.SIGNATURES[name]) == len(args), "%s: header has %d args, capi.py %d" % (name, len(args), len(capi.SIGNATURES[name]))
    assert set(capi.SIGNATURES) <= set(fns)
    assert lib.mgcr_abi_version() >= 1


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    st = lib.mgcr_ctx_create(0, C.byref(h))
    assert st == 1 and not h.value                      # MGCR_ERR_CUDA
    assert b"no CPU path" in lib.mgcr_last_error()


def test_slab_range_partition(lib):
    from mgpreconditionedgcr_b200 import host
    for n, align, nr in ((512, 4, 8), (512, 64, 8), (256, 4, 3), (48, 16, 2), (8, 8, 4)):
        prev = 0
        for r in range(nr):
            b, e = host.slab_range(n, align, r, nr)
            assert b == prev and b % align == 0 and e % align == 0 and e >= b
            prev = e
        assert prev == n
    with pytest.raises(Exception):
        host.slab_range(10, 4, 0, 2)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the product package or include/ may reference it."""
    bad = []
    for base in ("mgpreconditionedgcr_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    t = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"\boracle\b|pyoracle|liboracle|ref_oracle", t):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
