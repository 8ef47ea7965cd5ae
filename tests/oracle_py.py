"""numpy/ctypes face of the CPU oracle (oracle/fgb_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_ORACLE_DIR, "libfgb_oracle.so")


class orc_grid(C.Structure):
    _fields_ = [
        ("dims", C.c_int),
        ("min", C.c_float * 3),
        ("max", C.c_float * 3),
        ("radius", C.c_float),
        ("grid_dim", C.c_uint32 * 3),
        ("env_width", C.c_float * 3),
        ("wrap_compatible", C.c_int),
        ("bin_count", C.c_uint32),
    ]


_lib = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_ORACLE_DIR, "fgb_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            build_oracle()
        _lib = C.CDLL(_SO)
        _lib.orc_hash.restype = C.c_uint32
        _lib.orc_filter.restype = C.c_uint32
        _lib.orc_wrap_filter.restype = C.c_uint32
        _lib.orc_compact.restype = C.c_uint32
        _lib.orc_virtual.restype = C.c_float
        _lib.orc_virtual.argtypes = [C.c_float, C.c_float, C.c_float]
        _lib.orc_hash32.restype = C.c_uint32
        _lib.orc_hash32.argtypes = [C.c_uint32, C.c_uint32]
        _lib.orc_sort_max_bit.restype = C.c_int
        _lib.orc_sort_max_bit.argtypes = [C.c_void_p, C.c_int]
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_bucket_range.restype = C.c_uint32
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class Grid:
    def __init__(self, dims, mn, mx, radius):
        self.g = orc_grid()
        mn3 = (C.c_float * 3)(*[float(v) for v in (list(mn) + [0.0] * 3)[:3]])
        mx3 = (C.c_float * 3)(*[float(v) for v in (list(mx) + [0.0] * 3)[:3]])
        lib().orc_grid_init(C.byref(self.g), int(dims), mn3, mx3, C.c_float(radius))
        self.dims = dims
        self.grid_dim = tuple(int(v) for v in self.g.grid_dim)
        self.bin_count = int(self.g.bin_count)
        self.wrap_compatible = bool(self.g.wrap_compatible)
        self.radius = float(self.g.radius)
        self.min = tuple(float(v) for v in self.g.min)
        self.max = tuple(float(v) for v in self.g.max)
        self.env_width = tuple(float(v) for v in self.g.env_width)

    @property
    def ref(self):
        return C.byref(self.g)

    def bin_keys(self, x, y, z=None):
        x, y = f32(x), f32(y)
        z = f32(z) if z is not None else np.zeros_like(x)
        keys = np.empty(len(x), dtype=np.uint32)
        lib().orc_bin_keys(self.ref, C.c_uint32(len(x)), _p(x), _p(y), _p(z), _p(keys))
        return keys

    def build_index(self, x, y, z=None):
        """returns (pbm[bin_count+1], perm[n]) -- stable (source order inside a bin)"""
        x, y = f32(x), f32(y)
        z = f32(z) if z is not None else np.zeros_like(x)
        pbm = np.empty(self.bin_count + 1, dtype=np.uint32)
        perm = np.empty(len(x), dtype=np.uint32)
        lib().orc_build_index(self.ref, C.c_uint32(len(x)), _p(x), _p(y), _p(z), _p(pbm), _p(perm))
        return pbm, perm

    def filter(self, pbm, x, y, z=0.0, cap=1 << 16):
        out = np.empty(cap, dtype=np.uint32)
        n = lib().orc_filter(self.ref, _p(pbm), C.c_float(x), C.c_float(y), C.c_float(z), _p(out), C.c_uint32(cap))
        assert n <= cap
        return out[:n].copy()

    def wrap_filter(self, pbm, x, y, z=0.0, cap=1 << 16):
        out = np.empty(cap, dtype=np.uint32)
        n = lib().orc_wrap_filter(self.ref, _p(pbm), C.c_float(x), C.c_float(y), C.c_float(z), _p(out), C.c_uint32(cap))
        assert n <= cap
        return out[:n].copy()

    def sort_keys(self, x, y, z=None, true3d=False):
        x, y = f32(x), f32(y)
        zz = f32(z) if z is not None else None
        keys = np.empty(len(x), dtype=np.uint32)
        lib().orc_sort_keys(self.ref, C.c_int(int(true3d)), C.c_uint32(len(x)), _p(x), _p(y), _p(zz), _p(keys))
        return keys

    def sort_max_bit(self, true3d=False):
        return int(lib().orc_sort_max_bit(self.ref, C.c_int(int(true3d))))

    def sort_geometry(self, true3d=False):
        """(min[3], width[3], grid_dim[3]) as the reference passes them to calculateSpatialHash
        (CUDASimulation.cu:480-506), including its 3D quirk unless true3d."""
        mn, w, gd = (C.c_float * 3)(), (C.c_float * 3)(), (C.c_uint32 * 3)()
        lib().orc_sort_geometry(self.ref, C.c_int(int(true3d)), mn, w, gd)
        return [float(v) for v in mn], [float(v) for v in w], [int(v) for v in gd]

    def neighbour_count(self, pbm, mid, mx, my, mz, aid, ax, ay, az):
        out = np.empty(len(ax), dtype=np.uint32)
        lib().orc_neighbour_count(self.ref, _p(u32(pbm)), _p(u32(mid)), _p(f32(mx)), _p(f32(my)), _p(f32(mz)),
                                  C.c_uint32(len(ax)), _p(u32(aid)), _p(f32(ax)), _p(f32(ay)), _p(f32(az)), _p(out))
        return out

    def circles_step(self, ids, x, y, z, drift, repulse=0.05, do_sort=True, want_pbm=False):
        ids, x, y, z, drift = u32(ids).copy(), f32(x).copy(), f32(y).copy(), f32(z).copy(), f32(drift).copy()
        pbm = np.empty(self.bin_count + 1, dtype=np.uint32) if want_pbm else None
        lib().orc_circles_step(self.ref, C.c_uint32(len(x)), _p(ids), _p(x), _p(y), _p(z), _p(drift),
                               C.c_float(repulse), C.c_int(int(do_sort)), _p(pbm))
        return ids, x, y, z, drift, pbm


def bucket_build(lower, upper, keys):
    """MessageBucket index: (pbm[upper-lower+2], perm[n]) -- stable"""
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    pbm = np.empty(upper - lower + 2, dtype=np.uint32)
    perm = np.empty(len(keys), dtype=np.uint32)
    lib().orc_bucket_build(C.c_int32(lower), C.c_int32(upper), C.c_uint32(len(keys)), _p(keys), _p(pbm), _p(perm))
    return pbm, perm


def bucket_range(lower, upper, pbm, begin_key, end_key):
    """(first index, count) of MessageBucket::In::Filter(begin_key, end_key); operator()(key) is (key, key + 1)"""
    first = C.c_uint32()
    n = lib().orc_bucket_range(C.c_int32(lower), C.c_int32(upper), _p(pbm), C.c_int32(begin_key), C.c_int32(end_key), C.byref(first))
    return int(first.value), int(n)


def sort_perm(keys, max_bit):
    keys = u32(keys)
    perm = np.empty(len(keys), dtype=np.uint32)
    lib().orc_sort_perm(_p(keys), C.c_uint32(len(keys)), C.c_int(max_bit), _p(perm))
    return perm


def compact(flags, n, invert=False, keep_front=0):
    flags = u32(flags)
    perm = np.empty(max(n, 1), dtype=np.uint32)
    cnt = lib().orc_compact(_p(flags), C.c_int(int(invert)), C.c_uint32(n), C.c_uint32(keep_front), _p(perm))
    return perm[:cnt].copy()


def virtual(x2, x1, w):
    return float(lib().orc_virtual(C.c_float(x2), C.c_float(x1), C.c_float(w)))


def hash32(a, b):
    return int(lib().orc_hash32(C.c_uint32(a & 0xFFFFFFFF), C.c_uint32(b & 0xFFFFFFFF)))


def num_threads():
    return int(lib().orc_num_threads())
