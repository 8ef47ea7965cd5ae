"""Thin Python harness over the C ABI (tests, bench.py, multi-GPU driver).

The product is the sm_100a library plus the C++ API layer under include/flamegpu/; this module
only lends torch's device memory, streams and torch.distributed to it.  Names follow the
reference: a `Spatial` is a MessageSpatial2D/3D CUDAModelHandler, `Context` is the per-simulation
scatter/scan scratch (CUDAScatter + CUDAScanCompaction).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _capi

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _capi.load_library()
    return _lib


def _check(status: int, where: str):
    if status != 0:
        raise _capi.FgbError(lib(), status, where)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(stream: Optional[torch.cuda.Stream] = None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def make_vars(ins: Sequence[torch.Tensor], outs: Sequence[torch.Tensor]):
    """Build an fgb_var[] from matching in/out tensors (one per SoA variable)."""
    assert len(ins) == len(outs)
    arr = (_capi.fgb_var * max(len(ins), 1))()
    for i, (a, b) in enumerate(zip(ins, outs)):
        assert a.is_cuda and b.is_cuda and a.is_contiguous() and b.is_contiguous()
        per_item = a.element_size() * (a.numel() // max(a.shape[0], 1)) if a.dim() > 1 else a.element_size()
        arr[i].type_len = per_item
        arr[i].in_ = a.data_ptr()
        arr[i].out = b.data_ptr()
    return arr, len(ins)


class Context:
    """fgb_ctx: one per simulation."""

    def __init__(self, device: int = 0):
        self.device = device
        h = C.c_void_p()
        _check(lib().fgb_ctx_create(device, C.byref(h)), "fgb_ctx_create")
        self.h = h

    def close(self):
        if self.h:
            lib().fgb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(lib().fgb_launch_count(self.h))

    def reserve(self, n_max: int, max_bit: int = 0, stream_id: int = 0):
        _check(lib().fgb_ctx_reserve(self.h, stream_id, n_max, max_bit), "fgb_ctx_reserve")

    # -- cub::DeviceScan::ExclusiveSum replacement
    def exclusive_scan(self, inp: torch.Tensor, out: torch.Tensor, n: int, stream_id: int = 0):
        _check(lib().fgb_exclusive_scan_u32(self.h, stream_id, _ptr(inp), _ptr(out), n, _stream_ptr()),
               "fgb_exclusive_scan_u32")

    # -- CUDAScatter::scatter replacement (flags -> stable compaction of every variable)
    def compact(self, flags: Optional[torch.Tensor], ins, outs, n: int, *, invert=False, keep_front=0, out_offset=0,
                d_n: Optional[torch.Tensor] = None, d_out_offset: Optional[torch.Tensor] = None,
                d_out_count: Optional[torch.Tensor] = None, d_out_total: Optional[torch.Tensor] = None,
                stream_id: int = 0):
        arr, nv = make_vars(ins, outs)
        _check(lib().fgb_compact(self.h, stream_id, _ptr(flags), int(invert), n, _ptr(d_n), keep_front, out_offset,
                                 _ptr(d_out_offset), arr, nv, _ptr(d_out_count), _ptr(d_out_total), _stream_ptr()),
               "fgb_compact")

    def scatter_all(self, ins, outs, n: int, out_offset: int = 0, d_n=None, d_out_offset=None):
        arr, nv = make_vars(ins, outs)
        _check(lib().fgb_scatter_all(self.h, arr, nv, n, _ptr(d_n), out_offset, _ptr(d_out_offset), _stream_ptr()),
               "fgb_scatter_all")

    def gather(self, position: torch.Tensor, ins, outs, n: int, d_n=None):
        arr, nv = make_vars(ins, outs)
        _check(lib().fgb_gather(self.h, _ptr(position), arr, nv, n, _ptr(d_n), _stream_ptr()), "fgb_gather")

    def broadcast_init(self, defaults, outs, n: int, out_offset: int = 0):
        arr, nv = make_vars(defaults, outs)
        for i, d in enumerate(defaults):
            arr[i].type_len = d.numel() * d.element_size()
        _check(lib().fgb_broadcast_init(self.h, arr, nv, n, out_offset, _stream_ptr()), "fgb_broadcast_init")

    _DTYPES = {torch.float32: _capi.FGB_F32, torch.float64: _capi.FGB_F64, torch.int32: _capi.FGB_I32, torch.int64: _capi.FGB_I64}

    def reduce(self, inp: torch.Tensor, n: int, op: str, *, unsigned=False, d_n=None, stream_id: int = 0):
        """sum / min / max of inp[:n] on the device; returns a Python number (sums: float or exact int)"""
        dt = self._DTYPES[inp.dtype]
        if unsigned:
            dt = {_capi.FGB_I32: _capi.FGB_U32, _capi.FGB_I64: _capi.FGB_U64}[dt]
        opc = {"sum": _capi.FGB_REDUCE_SUM, "min": _capi.FGB_REDUCE_MIN, "max": _capi.FGB_REDUCE_MAX}[op]
        out = torch.zeros(1, dtype=torch.int64, device=inp.device)
        _check(lib().fgb_reduce(self.h, stream_id, opc, dt, _ptr(inp), n, _ptr(d_n), _ptr(out), _stream_ptr()), "fgb_reduce")
        raw = out.cpu().numpy()
        import numpy as np

        if op == "sum":
            if dt in (_capi.FGB_F32, _capi.FGB_F64):
                return float(raw.view(np.float64)[0])
            return int(raw.view(np.uint64)[0]) if unsigned else int(raw[0])
        view = {_capi.FGB_F32: np.float32, _capi.FGB_F64: np.float64, _capi.FGB_I32: np.int32, _capi.FGB_U32: np.uint32,
                _capi.FGB_I64: np.int64, _capi.FGB_U64: np.uint64}[dt]
        return raw.view(view)[0].item()

    # -- CUDAScatter::arrayMessageReorder replacement: out[index[i]] = in[i]; returns nothing, max writes per element -> d_max
    def array_reorder(self, index, array_length: int, ins, outs, n: int, write_count=None, d_max=None, d_n=None, stream_id: int = 0):
        arr, nv = make_vars(ins, outs)
        _check(lib().fgb_array_reorder(self.h, stream_id, _ptr(index), array_length, arr, nv, n, _ptr(d_n), _ptr(write_count), _ptr(d_max),
                                       _stream_ptr()), "fgb_array_reorder")

    # -- CUDAScatter::scatterNewAgents replacement: n structs of agent_size bytes (device) -> SoA columns at out_offset
    def scatter_new_agents(self, aos: torch.Tensor, agent_size: int, offsets, lens, outs, n: int, out_offset: int = 0, d_out_offset=None):
        arr = (_capi.fgb_var * max(len(outs), 1))()
        for i, (off, ln, o) in enumerate(zip(offsets, lens, outs)):
            arr[i].type_len = ln
            arr[i].in_ = aos.data_ptr() + off
            arr[i].out = o.data_ptr()
        _check(lib().fgb_scatter_new_agents(self.h, _ptr(aos), agent_size, arr, len(outs), n, out_offset, _ptr(d_out_offset), _stream_ptr()),
               "fgb_scatter_new_agents")

    def histogram_even(self, inp: torch.Tensor, n: int, bins: int, lower: float, upper: float, *, unsigned=False, d_n=None):
        dt = self._DTYPES[inp.dtype]
        if unsigned:
            dt = {_capi.FGB_I32: _capi.FGB_U32, _capi.FGB_I64: _capi.FGB_U64}[dt]
        out = torch.zeros(bins, dtype=torch.int32, device=inp.device)
        _check(lib().fgb_histogram_even(self.h, dt, _ptr(inp), n, _ptr(d_n), bins, float(lower), float(upper), _ptr(out), _stream_ptr()),
               "fgb_histogram_even")
        return out

    def sort_keys(self, x, y, z, env_min, env_width, grid_dim, n: int, keys_out: torch.Tensor, d_n=None):
        mn = (C.c_float * 3)(*[float(v) for v in (list(env_min) + [0.0] * 3)[:3]])
        w = (C.c_float * 3)(*[float(v) for v in (list(env_width) + [1.0] * 3)[:3]])
        g = (C.c_uint * 3)(*[int(v) for v in (list(grid_dim) + [1] * 3)[:3]])
        _check(lib().fgb_sort_keys(self.h, _ptr(x), _ptr(y), _ptr(z), mn, w, g, n, _ptr(d_n), _ptr(keys_out),
                                   _stream_ptr()), "fgb_sort_keys")

    def sort_spatial(self, x, y, z, env_min, env_width, grid_dim, max_bit: int, keys_out: torch.Tensor, ins, outs, n: int,
                     position_out=None, d_n=None, stream_id: int = 0):
        arr, nv = make_vars(ins, outs)
        mn = (C.c_float * 3)(*[float(v) for v in env_min])
        wd = (C.c_float * 3)(*[float(v) for v in env_width])
        gd = (C.c_uint * 3)(*[int(v) for v in grid_dim])
        _check(lib().fgb_sort_spatial(self.h, stream_id, _ptr(x), _ptr(y), _ptr(z), mn, wd, gd, max_bit, n, _ptr(d_n), _ptr(keys_out),
                                      arr, nv, _ptr(position_out), _stream_ptr()), "fgb_sort_spatial")

    def sort_by_key(self, keys: torch.Tensor, max_bit: int, ins, outs, n: int, position_out=None, d_n=None,
                    stream_id: int = 0):
        arr, nv = make_vars(ins, outs)
        _check(lib().fgb_sort_by_key(self.h, stream_id, _ptr(keys), max_bit, n, _ptr(d_n), arr, nv, _ptr(position_out),
                                     _stream_ptr()), "fgb_sort_by_key")


class Spatial:
    """fgb_spatial: MessageSpatial2D/3D::CUDAModelHandler (PBM owner) for one message list."""

    def __init__(self, ctx: Context, dims: int, env_min, env_max, radius: float):
        self.ctx = ctx
        self.dims = dims
        mn = (C.c_float * 3)(*[float(v) for v in (list(env_min) + [0.0] * 3)[:3]])
        mx = (C.c_float * 3)(*[float(v) for v in (list(env_max) + [0.0] * 3)[:3]])
        h = C.c_void_p()
        _check(lib().fgb_spatial_create(ctx.h, dims, mn, mx, float(radius), C.byref(h)), "fgb_spatial_create")
        self.h = h
        md = _capi.fgb_spatial_metadata()
        bc = C.c_uint()
        _check(lib().fgb_spatial_get_metadata(self.h, C.byref(md), C.byref(bc)), "fgb_spatial_get_metadata")
        self.metadata = md
        self.bin_count = int(bc.value)
        self.grid_dim = tuple(int(v) for v in md.grid_dim)
        self.wrap_compatible = bool(md.wrap_compatible)
        self.radius = float(md.radius)

    def close(self):
        if self.h:
            lib().fgb_spatial_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def metadata_device_ptr(self) -> int:
        return int(lib().fgb_spatial_metadata_device_ptr(self.h) or 0)

    def reserve(self, n_max: int):
        _check(lib().fgb_spatial_reserve(self.h, n_max), "fgb_spatial_reserve")

    def pbm(self):
        """Host copy of the PBM (bin_count + 1 entries) as a numpy uint32 array (synchronises)."""
        import numpy as np

        out = np.empty(self.bin_count + 1, dtype=np.uint32)
        _check(lib().fgb_spatial_read_pbm(self.h, out.ctypes.data_as(C.c_void_p), _stream_ptr()),
               "fgb_spatial_read_pbm")
        return out

    def bin_permutation(self, x, y, z, perm_out, n: int, *, stable=False, tile_local=False, d_n=None):
        flags = _capi.FGB_BUILD_STABLE if stable else _capi.FGB_BUILD_DEFAULT
        if tile_local:
            flags |= _capi.FGB_BUILD_TILE_LOCAL
        _check(lib().fgb_bin_permutation(self.h, n, _ptr(d_n), _ptr(x), _ptr(y), _ptr(z), _ptr(perm_out), flags, _stream_ptr()),
               "fgb_bin_permutation")

    def build_index(self, x, y, z, ins, outs, n: int, *, stable=False, d_n=None, expect_grouped=False):
        arr, nv = make_vars(ins, outs)
        flags = _capi.FGB_BUILD_STABLE if stable else _capi.FGB_BUILD_DEFAULT
        if expect_grouped:
            flags |= _capi.FGB_BUILD_EXPECT_GROUPED
        _check(lib().fgb_build_index(self.h, n, _ptr(d_n), _ptr(x), _ptr(y), _ptr(z), arr, nv, flags, _stream_ptr()),
               "fgb_build_index")


class Bucket:
    """Index over an integer key range (MessageBucket): PBM of upper - lower + 2 words."""

    def __init__(self, ctx: Context, lower: int, upper: int):
        self.ctx = ctx
        h = C.c_void_p()
        _check(lib().fgb_bucket_create(ctx.h, int(lower), int(upper), C.byref(h)), "fgb_bucket_create")
        self.h = h
        self.lower, self.upper = int(lower), int(upper)
        self.bin_count = self.upper - self.lower + 1

    def close(self):
        if getattr(self, "h", None):
            lib().fgb_spatial_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bounds(self):
        mn, mx, p = C.c_int(), C.c_int(), C.c_void_p()
        _check(lib().fgb_bucket_get_bounds(self.h, C.byref(mn), C.byref(mx), C.byref(p)), "fgb_bucket_get_bounds")
        return mn.value, mx.value, p.value

    def pbm(self):
        import numpy as np

        out = np.empty(self.bin_count + 1, dtype=np.uint32)
        _check(lib().fgb_spatial_read_pbm(self.h, out.ctypes.data_as(C.c_void_p), _stream_ptr()), "fgb_spatial_read_pbm")
        return out

    def build_index(self, keys, ins, outs, n: int, *, stable=False, d_n=None):
        arr, nv = make_vars(ins, outs)
        flags = _capi.FGB_BUILD_STABLE if stable else _capi.FGB_BUILD_DEFAULT
        _check(lib().fgb_build_index_keys(self.h, n, _ptr(d_n), _ptr(keys), arr, nv, flags, _stream_ptr()), "fgb_build_index_keys")
