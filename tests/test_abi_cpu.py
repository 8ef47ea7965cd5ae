"""CPU-side checks of the drop-in boundary: the C ABI library loads, exports every symbol that
include/flamegpu2_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "flamegpu2_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fgb_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    from flamegpu2_b200 import _capi

    lib = _capi.load_library()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # and the Python signature table covers the whole header
    assert set(names) == set(_capi.SIGNATURES), set(names) ^ set(_capi.SIGNATURES)
    assert lib.fgb_version() == 100


def test_no_cpu_fallback():
    import torch

    from flamegpu2_b200 import _capi

    lib = _capi.load_library()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.fgb_ctx_create(0, C.byref(h))
    assert rc == _capi.FGB_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.fgb_error_string(rc)


def test_struct_layouts_match_reference():
    from flamegpu2_b200 import _capi

    # CUDAScatter::ScatterData {size_t typeLen; char *in; char *out;} (CUDAScatter.cuh:58-62)
    assert C.sizeof(_capi.fgb_var) == 24
    # MessageSpatial3D::MetaData (MessageSpatial3D.h:38-68): 3f,3f,f,(pad)ptr,3u,3f,bool -> 72 bytes
    md = _capi.fgb_spatial_metadata
    assert C.sizeof(md) == 72
    assert md.PBM.offset == 32 and md.grid_dim.offset == 40 and md.environment_width.offset == 52
    assert md.wrap_compatible.offset == 64


def test_description_validation_host_side():
    # the reference's DescriptionValidation / reserved_name host tests (test_bucket.cu:23-52, test_spatial_3d.cu),
    # restated in C++ inside libfgb_models.so (fgbm_selftest_descriptions); pure host code, no GPU
    from flamegpu2_b200 import sim

    assert sim.lib().fgbm_selftest_descriptions() == 0
