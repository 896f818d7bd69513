"""The C-ABI library must load on a machine without a GPU and export every symbol that
include/km_b200.h declares (no compute is attempted here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from km_b200 import _lib
    return _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "km_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(km_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(built):
    L = ctypes.CDLL(built.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libkm_b200.so does not export %s" % n
    assert sorted(built.EXPORTS) == names


def test_row_layout_matches_header(built):
    assert built.ROW_DTYPE.itemsize == 112
    assert built.ROW_DTYPE.fields["min_cov"][1] == 72 and built.ROW_DTYPE.fields["rvaf"][1] == 80


def test_no_gpu_means_loud_failure(built):
    L = built.lib()
    if L.km_device_count() > 0:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = L.km_table_create(0, 31, 1, 1024, ctypes.byref(h))
    assert rc == -6 and b"no CPU fallback" in L.km_last_error()
    from km_b200 import engine
    with pytest.raises(built.KmError):
        engine.Table.create()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "km_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libkmer_store" not in src and "oracle.store" not in src, f
