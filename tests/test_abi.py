"""The C-ABI shared library loads on a machine without a GPU and exports every symbol include/urso_b200.h declares."""
import ctypes
import os
import re

from ursonet_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "urso_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(urso_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 25
    dll = ctypes.CDLL(lib.LIB_PATH)
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, missing


def test_binding_table_matches_header():
    assert sorted(lib.SIGNATURES) == declared_symbols()


def test_load_checks_struct_layout_and_version():
    l = lib.load()
    assert l.urso_version() >= 100
    assert l.urso_sizeof_convgemm_desc() == ctypes.sizeof(lib.ConvGemmDesc)
    assert l.urso_sizeof_wgrad_desc() == ctypes.sizeof(lib.WgradDesc)


def test_errors_are_reported_not_swallowed():
    import pytest
    with pytest.raises(lib.UrsoError):
        lib.call("urso_dense_fwd", None, None, None, 1, 1, 1, None)      # null pointers -> rc != 0 + message
    assert b"null" in lib.load().urso_last_error()


def test_aug_params_layout_matches_header():
    """ursonet_b200.augment.AUG_DTYPE is the numpy image of struct urso_aug_params (include/urso_b200.h)."""
    from ursonet_b200 import augment, lib
    assert lib.load().urso_sizeof_aug_params() == augment.AUG_DTYPE.itemsize
    augment.check_layout()
