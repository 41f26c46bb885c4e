"""CPU: the CUDA shared library is built in-tree, loads, and exports every function that
include/biceps_b200.h declares (no compute calls are made here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_]*\s*\**\s+\**(bcp_[a-z0-9_]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    return g.LIB


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("bcp_create", "bcp_destroy", "bcp_chains_create", "bcp_sample", "bcp_sample_launch",
                 "bcp_chains_fetch", "bcp_neglogp", "bcp_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, "symbols declared in include/*.h but not exported: %s" % missing
    lib.bcp_abi_version.restype = ctypes.c_int
    assert lib.bcp_abi_version() == 1


def test_ctypes_binding_covers_the_header(built):
    from biceps_b200 import _cabi
    assert sorted(_cabi._SIGS) == declared_functions()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under biceps_b200/ may import, include, link or
    dlopen anything from oracle/."""
    pkg = os.path.join(ROOT, "biceps_b200")
    bad = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.\.?oracle)|#include\s*[\"<][^\">]*oracle|liboracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), "%s reaches into oracle/" % os.path.join(dirpath, f)
