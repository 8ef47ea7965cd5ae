"""ctypes binding of the C ABI in include/flamegpu2_b200.h.

The shared library holds hand-written sm_100a kernels only; there is no CPU fallback.  Importing
this module without the built library raises ImportError, and every entry point returns
FGB_ERR_NO_DEVICE on a machine without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FGB_KERNELS_LIB") or os.path.join(_HERE, "lib", "libflamegpu2_b200.so")  # override: A/B builds

FGB_MAX_VARS = 32
FGB_BUILD_DEFAULT = 0
FGB_BUILD_STABLE = 1
FGB_BUILD_TILE_LOCAL = 2
FGB_BUILD_KEYS_READY = 4
FGB_BUILD_EXPECT_GROUPED = 8
FGB_REDUCE_SUM, FGB_REDUCE_MIN, FGB_REDUCE_MAX = 0, 1, 2
FGB_F32, FGB_F64, FGB_I32, FGB_U32, FGB_I64, FGB_U64 = range(6)
FGB_ERR_NO_DEVICE = -3


class fgb_var(C.Structure):
    """Mirror of fgb_var == CUDAScatter::ScatterData (reference CUDAScatter.cuh:58-62)."""

    _fields_ = [("type_len", C.c_size_t), ("in_", C.c_void_p), ("out", C.c_void_p)]


class fgb_spatial_metadata(C.Structure):
    """Mirror of MessageSpatial3D::MetaData (reference MessageSpatial3D.h:38-68)."""

    _fields_ = [
        ("min", C.c_float * 3),
        ("max", C.c_float * 3),
        ("radius", C.c_float),
        ("PBM", C.c_void_p),
        ("grid_dim", C.c_uint * 3),
        ("environment_width", C.c_float * 3),
        ("wrap_compatible", C.c_bool),
    ]


# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "fgb_version": (C.c_int, []),
    "fgb_error_string": (C.c_char_p, [C.c_int]),
    "fgb_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "fgb_ctx_destroy": (C.c_int, [C.c_void_p]),
    "fgb_launch_count": (C.c_ulonglong, [C.c_void_p]),
    "fgb_alloc_generation": (C.c_ulonglong, [C.c_void_p]),
    "fgb_spatial_create": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float,
                                     C.POINTER(C.c_void_p)]),
    "fgb_spatial_create_window": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float,
                                            C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "fgb_spatial_get_window": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "fgb_plane_flags": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgb_array_reorder": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.POINTER(fgb_var), C.c_uint, C.c_uint, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "fgb_scatter_new_agents": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint, C.POINTER(fgb_var), C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p]),
    "fgb_histogram_even": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_double, C.c_double, C.c_void_p,
                                     C.c_void_p]),
    "fgb_slab_signal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgb_slab_wait": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint,
                                C.c_void_p]),
    "fgb_slab_reserve": (C.c_int, [C.c_void_p, C.c_uint, C.c_uint]),
    "fgb_slab_migrate_out": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                       C.c_uint, C.POINTER(fgb_var), C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "fgb_slab_pack_planes": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                       C.c_uint, C.POINTER(fgb_var), C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgb_slab_check_bound": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]),
    "fgb_slab_allreduce": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_uint, C.c_void_p]),
    "fgb_spatial_destroy": (C.c_int, [C.c_void_p]),
    "fgb_spatial_use_pbm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fgb_spatial_get_metadata": (C.c_int, [C.c_void_p, C.POINTER(fgb_spatial_metadata), C.POINTER(C.c_uint)]),
    "fgb_spatial_metadata_device_ptr": (C.c_void_p, [C.c_void_p]),
    "fgb_spatial_bin_count": (C.c_uint, [C.c_void_p]),
    "fgb_spatial_read_pbm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgb_spatial_reserve": (C.c_int, [C.c_void_p, C.c_uint]),
    "fgb_ctx_reserve": (C.c_int, [C.c_void_p, C.c_uint, C.c_uint, C.c_int]),
    "fgb_spatial_writer_args": (C.c_int, [C.c_void_p, C.c_uint, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "fgb_spatial_clear_histogram": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fgb_build_index_ex": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(fgb_var), C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgb_build_index": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.POINTER(fgb_var), C.c_uint, C.c_uint, C.c_void_p]),
    "fgb_sort_spatial": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                   C.POINTER(C.c_uint), C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.POINTER(fgb_var), C.c_uint,
                                   C.c_void_p, C.c_void_p]),
    "fgb_reduce": (C.c_int, [C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgb_transform_reduce": (C.c_int, [C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "fgb_bucket_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "fgb_bucket_get_bounds": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "fgb_build_index_keys": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.POINTER(fgb_var), C.c_uint, C.c_uint,
                                       C.c_void_p]),
    "fgb_bin_permutation": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_uint, C.c_void_p]),
    "fgb_exclusive_scan_u32": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "fgb_compact": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_int, C.c_uint, C.c_void_p, C.c_uint, C.c_uint,
                              C.c_void_p, C.POINTER(fgb_var), C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgb_compact_limited": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_int, C.c_uint, C.c_void_p, C.c_uint, C.c_uint,
                                      C.c_void_p, C.c_uint, C.POINTER(fgb_var), C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fgb_scatter_all": (C.c_int, [C.c_void_p, C.POINTER(fgb_var), C.c_uint, C.c_uint, C.c_void_p, C.c_uint,
                                  C.c_void_p, C.c_void_p]),
    "fgb_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(fgb_var), C.c_uint, C.c_uint, C.c_void_p, C.c_void_p]),
    "fgb_broadcast_init": (C.c_int, [C.c_void_p, C.POINTER(fgb_var), C.c_uint, C.c_uint, C.c_uint, C.c_void_p]),
    "fgb_sort_keys": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float),
                                C.POINTER(C.c_float), C.POINTER(C.c_uint), C.c_uint, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "fgb_sort_by_key": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_int, C.c_uint, C.c_void_p,
                                  C.POINTER(fgb_var), C.c_uint, C.c_void_p, C.c_void_p]),
}


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "flamegpu2_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    return lib


class FgbError(RuntimeError):
    def __init__(self, lib, status: int, where: str):
        self.status = status
        msg = lib.fgb_error_string(status).decode()
        super().__init__(f"{where} failed: {msg} (status {status})")
