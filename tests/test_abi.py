"""The C-ABI library loads without a GPU and exports every symbol include/dregb200.h declares."""
import ctypes
import os
import re

import pytest

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
    assert lib.drb_abi_version() == 4
    assert lib.drb_ngp_table_entries() == 6299960
    assert lib.drb_downsample_workspace_bytes(1000, 260) > 1000 * 260 * 4


def test_argument_validation_without_gpu(pkg):
    """Entry points reject bad arguments before touching the device (error code + message)."""
    lib = pkg.load_library()
    assert lib.drb_conv3d_igemm(None, None) == -1
    assert b"null" in lib.drb_last_error()
    assert lib.drb_split_planes(None, None, None, 5, None) == -1
    assert lib.drb_engine_num_params(None) == 0


def test_integration_stub_matches_the_binding(pkg):
    """INTEGRATION.md section 3 shows a maintainer the ctypes mirror of drb_conv3d_desc: it must list the fields of
    the struct the library reads (a short struct would be read past its end - VERDICT r1 found it stale)."""
    from importlib import import_module
    lib_mod = import_module("dreg-nerf_b200._lib")
    with open(os.path.join(ROOT, "INTEGRATION.md")) as fh:
        text = fh.read()
    block = text[text.index("class Conv3dDesc"):text.index("assert lib.drb_abi_version")]
    stub = re.findall(r'\("(\w+)", C\.', block)
    assert stub == [f[0] for f in lib_mod.Conv3dDesc._fields_]
    assert "drb_abi_version() == %d" % pkg.load_library().drb_abi_version() in text
    # and the header's struct has the same member names in the same order
    with open(os.path.join(ROOT, "include", "dregb200.h")) as fh:
        header = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    body = header[header.index("typedef struct drb_conv3d_desc {"):header.index("} drb_conv3d_desc;")]
    members = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if " " in decl else decl
        for part in names.split(","):
            m = re.search(r"(\w+)\s*(\[\d+\])?$", part.strip())
            if m:
                members.append(m.group(1))
    assert members == stub, (members, stub)


def _header_members(struct):
    with open(os.path.join(ROOT, "include", "dregb200.h")) as fh:
        header = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    body = header[header.index("typedef struct %s {" % struct):header.index("} %s;" % struct)]
    members = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if " " in decl else decl
        for part in names.split(","):
            m = re.search(r"(\w+)\s*(\[\d+\])?$", part.strip())
            if m:
                members.append(m.group(1))
    return members


@pytest.mark.parametrize("struct,cls", [("drb_wgrad_desc", "WgradDesc"), ("drb_conv3d_desc", "Conv3dDesc"),
                                        ("drb_extract_desc", "ExtractDesc"), ("drb_engine_config", "EngineConfig")])
def test_ctypes_structs_mirror_the_header(pkg, struct, cls):
    """Every descriptor the binding fills has the header's members, in the header's order (ABI 4 appended
    `stage` / `stage_elems` to drb_wgrad_desc: a binding built for ABI 3 would pass a short struct)."""
    from importlib import import_module
    lib_mod = import_module("dreg-nerf_b200._lib")
    assert [f[0] for f in getattr(lib_mod, cls)._fields_] == _header_members(struct)
