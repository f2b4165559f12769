"""CPU: libhpv.so loads, exports every symbol include/hpv.h declares, the ctypes table covers the header, and
the product fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

import hpv_b200
from hpv_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "hpv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hpv_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_a_nontrivial_surface():
    fns = header_functions()
    assert len(fns) >= 30
    for must in ("hpv_create", "hpv_varloss_forward", "hpv_varloss_backward", "hpv_net_u", "hpv_train_steps"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libhpv.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for fn in header_functions():
        assert hasattr(lib, fn), "libhpv.so does not export %s" % fn


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == header_functions()
    lib = hpv_b200.load()
    assert lib.hpv_abi_version() == 1


def test_no_cpu_fallback_without_a_gpu():
    lib = hpv_b200.load()
    if lib.hpv_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(hpv_b200.HpvError) as ei:
        hpv_b200.Engine(0)
    assert "no CPU path" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hp-vpinns_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), "%s mentions the oracle" % f
