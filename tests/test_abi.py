"""The C-ABI library loads on a CPU-only box and exports every symbol include/phylok.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import REPO
from kamphir_b200 import _lib


def declared_symbols():
    text = open(os.path.join(REPO, "include", "phylok.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 30
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), "libphylok.so does not export %s" % s
        assert s in _lib.SIGNATURES, "kamphir_b200/_lib.py does not bind %s" % s
    assert sorted(_lib.SIGNATURES) == syms
    assert _lib.lib().pk_abi_version() == 1


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a device the compute entry points must raise; with one this test is skipped."""
    if _lib.lib().pk_device_count() > 0:
        pytest.skip("a GPU is present")
    from kamphir_b200 import PhyloKernel
    pk = PhyloKernel()
    with pytest.raises(_lib.PhyloKError):
        pk.kernel("(a:1,b:2);", "(a:1,b:2);")


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(REPO, "kamphir_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\S*oracle|libpk_oracle|#include.*oracle", src, re.M), f
