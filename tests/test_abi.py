"""The C-ABI library loads without a GPU and exports every symbol include/dregb200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "dregb200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(drb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(pkg):
    from importlib import import_module
    lib_mod = import_module("dreg-nerf_b200._lib")
    names = _declared()
    assert len(names) >= 40
    raw = ctypes.CDLL(lib_mod.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libdregb200.so does not export %s" % n
        assert n in lib_mod.SIGNATURES, "ctypes binding misses %s" % n
    assert sorted(lib_mod.SIGNATURES) == names, "binding declares symbols the header does not"


def test_library_loads_and_reports_version(pkg):
    lib = pkg.load_library()
    assert lib.drb_abi_version() == 3
    assert lib.drb_ngp_table_entries() == 6299960
    assert lib.drb_downsample_workspace_bytes(1000, 260) > 1000 * 260 * 4


def test_argument_validation_without_gpu(pkg):
    """Entry points reject bad arguments before touching the device (error code + message)."""
    lib = pkg.load_library()
    assert lib.drb_conv3d_igemm(None, None) == -1
    assert b"null" in lib.drb_last_error()
    assert lib.drb_split_planes(None, None, None, 5, None) == -1
    assert lib.drb_engine_num_params(None) == 0
