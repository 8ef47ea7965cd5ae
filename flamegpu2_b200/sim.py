"""ctypes face of libfgb_models.so: whole CUDASimulation::step() runs of the example models
(examples/*.cuh compiled against include/flamegpu, the C++ API layer of this repo).

Used by tests, bench.py and the multi-GPU driver.  No CPU fallback: creating a Simulation without a
GPU raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FGB_MODELS_LIB") or os.path.join(_HERE, "lib", "libfgb_models.so")  # override: A/B builds

_lib = None

TEST_MODELS = {
    "count3d": 0, "optional3d": 1, "wrap3d": 2, "count2d": 3, "wrap2d": 4, "death": 5,
    "birth_mandatory": 6, "birth_optional": 7, "birth_optional_death": 8, "birth_other_agent": 9,
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
        # the kernel library must be loaded first (the models library links against it by rpath)
        from . import host

        host.lib()
        L = C.CDLL(LIB_PATH)
        L.fgbm_last_error.restype = C.c_char_p
        L.fgbm_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        L.fgbm_destroy.argtypes = [C.c_void_p]
        L.fgbm_set_population.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint, C.c_uint,
                                          C.POINTER(C.c_char_p), C.POINTER(C.c_void_p)]
        L.fgbm_get_count.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_uint)]
        L.fgbm_get_variable.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t]
        L.fgbm_step.argtypes = [C.c_void_p, C.c_uint]
        L.fgbm_sync.argtypes = [C.c_void_p]
        L.fgbm_stream.argtypes = [C.c_void_p]
        L.fgbm_stream.restype = C.c_void_p
        L.fgbm_launch_count.argtypes = [C.c_void_p]
        L.fgbm_launch_count.restype = C.c_ulonglong
        L.fgbm_graph_count.argtypes = [C.c_void_p]
        L.fgbm_graph_count.restype = C.c_uint
        L.fgbm_graph_width.argtypes = [C.c_void_p]
        L.fgbm_graph_width.restype = C.c_uint
        L.fgbm_list_bound.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        L.fgbm_step_counter.argtypes = [C.c_void_p]
        L.fgbm_step_counter.restype = C.c_uint
        L.fgbm_step_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_uint, C.POINTER(C.c_uint)]
        L.fgbm_profile.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.fgbm_message_count.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint)]
        L.fgbm_message_pbm.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_uint)]
        L.fgbm_message_variable.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t]
        L.fgbm_circles_step_host.argtypes = [C.c_void_p, C.c_uint] + [C.c_void_p] * 4 + [C.c_uint] + [C.c_void_p] * 5
        _lib = L
    return _lib


def _check(rc: int, where: str):
    if rc != 0:
        raise RuntimeError(f"{where}: {lib().fgbm_last_error().decode()}")


class Simulation:
    """One CUDASimulation of a named example model ("circles", "boids3d", "boids2d", "stress", "test")."""

    def __init__(self, model: str, device: int = 0, **params):
        self.model = model
        p = ",".join(f"{k}={v}" for k, v in params.items())
        h = C.c_void_p()
        _check(lib().fgbm_create(model.encode(), p.encode(), device, C.byref(h)), "fgbm_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().fgbm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_population(self, agent: str, variables: Dict[str, np.ndarray], state: Optional[str] = None):
        names = list(variables)
        arrs = [np.ascontiguousarray(variables[k]) for k in names]
        n = len(arrs[0]) if arrs else 0
        cn = (C.c_char_p * max(len(names), 1))(*[k.encode() for k in names])
        cp = (C.c_void_p * max(len(names), 1))(*[a.ctypes.data for a in arrs])
        _check(lib().fgbm_set_population(self.h, agent.encode(), state.encode() if state else None, n, len(names), cn, cp),
               "fgbm_set_population")

    def count(self, agent: str, state: Optional[str] = None) -> int:
        n = C.c_uint()
        _check(lib().fgbm_get_count(self.h, agent.encode(), state.encode() if state else None, C.byref(n)), "fgbm_get_count")
        return int(n.value)

    def get(self, agent: str, var: str, dtype, elements: int = 1, state: Optional[str] = None) -> np.ndarray:
        n = self.count(agent, state)
        shape = (n,) if elements == 1 else (n, elements)
        out = np.empty(shape, dtype=dtype)
        _check(lib().fgbm_get_variable(self.h, agent.encode(), state.encode() if state else None, var.encode(),
                                       out.ctypes.data_as(C.c_void_p), out.nbytes), "fgbm_get_variable")
        return out

    def agent_reduce(self, agent: str, var: str, op: str, kind: str, value: float = 0.0) -> float:
        """HostAgentAPI::sum/min/max/count(value)/mean/std of an agent variable (kind: 'f' float, 'i' int, 'u' unsigned int)"""
        out = C.c_double(value)
        lib().fgbm_agent_reduce.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char, C.POINTER(C.c_double)]
        _check(lib().fgbm_agent_reduce(self.h, agent.encode(), var.encode(), {"sum": 0, "min": 1, "max": 2, "count": 3, "mean": 4, "std": 5}[op], kind.encode(), C.byref(out)),
               "fgbm_agent_reduce")
        return float(out.value)

    def agent_histogram(self, agent: str, var: str, kind: str, bins: int, lower: float, upper: float) -> np.ndarray:
        """HostAgentAPI::histogramEven of an agent variable (kind: 'f' float, 'i' int, 'u' unsigned int)"""
        out = np.zeros(bins, np.uint32)
        lib().fgbm_agent_histogram.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char, C.c_uint, C.c_double, C.c_double, C.c_void_p]
        _check(lib().fgbm_agent_histogram(self.h, agent.encode(), var.encode(), kind.encode(), bins, lower, upper, out.ctypes.data_as(C.c_void_p)),
               "fgbm_agent_histogram")
        return out

    def agent_custom_reduce(self, agent: str, var: str, which: int, kind: str) -> float:
        """HostAgentAPI::reduce / transformReduce with user functors (0 custom sum, 1 custom max, 2 count of a <= 0)"""
        out = C.c_double(0.0)
        lib().fgbm_agent_custom_reduce.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char, C.POINTER(C.c_double)]
        _check(lib().fgbm_agent_custom_reduce(self.h, agent.encode(), var.encode(), which, kind.encode(), C.byref(out)), "fgbm_agent_custom_reduce")
        return float(out.value)

    def step(self, steps: int = 1):
        _check(lib().fgbm_step(self.h, steps), "fgbm_step")

    def simulate(self, steps: int):
        """CUDASimulation::simulate(): init functions, `steps` steps, exit functions"""
        lib().fgbm_simulate.argtypes = [C.c_void_p, C.c_uint]
        _check(lib().fgbm_simulate(self.h, steps), "fgbm_simulate")

    def sync(self):
        _check(lib().fgbm_sync(self.h), "fgbm_sync")

    @property
    def stream(self) -> int:
        return int(lib().fgbm_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(lib().fgbm_launch_count(self.h))

    @property
    def graphs(self) -> int:
        return int(lib().fgbm_graph_count(self.h))

    @property
    def graph_width(self) -> int:
        """Kernel nodes on the widest level of the last captured step graph (> 1: parallel branches)."""
        return int(lib().fgbm_graph_width(self.h))

    def list_bound(self, agent: str, state: Optional[str] = None):
        """(host-side launch bound, capacity) of a state list."""
        b, c = C.c_uint(), C.c_uint()
        _check(lib().fgbm_list_bound(self.h, agent.encode(), state.encode() if state else None, C.byref(b), C.byref(c)), "fgbm_list_bound")
        return int(b.value), int(c.value)

    @property
    def step_counter(self) -> int:
        return int(lib().fgbm_step_counter(self.h))

    def step_times(self) -> np.ndarray:
        """Device seconds of the steps run since the previous call (model created with timing=1).
        CUDASimulation::getElapsedTimeSteps() keeps the whole history like the reference; this returns the new part."""
        n = C.c_uint()
        _check(lib().fgbm_step_times(self.h, None, 0, C.byref(n)), "fgbm_step_times")
        total = int(n.value)
        buf = (C.c_double * max(total, 1))()
        _check(lib().fgbm_step_times(self.h, buf, total, C.byref(n)), "fgbm_step_times")
        seen = getattr(self, "_times_seen", 0)
        self._times_seen = total
        return np.array(buf[seen:total])

    def profile(self) -> dict:
        """{phase: (total_ms, calls)} since the last call (model created with profile=1)."""
        import json

        buf = C.create_string_buffer(1 << 16)
        _check(lib().fgbm_profile(self.h, buf, len(buf)), "fgbm_profile")
        return {k: (v[0], v[1]) for k, v in json.loads(buf.value.decode()).items()}

    def message_count(self, message: str) -> int:
        n = C.c_uint()
        _check(lib().fgbm_message_count(self.h, message.encode(), C.byref(n)), "fgbm_message_count")
        return int(n.value)

    def message_pbm(self, message: str) -> np.ndarray:
        bins = C.c_uint()
        _check(lib().fgbm_message_pbm(self.h, message.encode(), None, C.byref(bins)), "fgbm_message_pbm")
        out = np.empty(bins.value + 1, dtype=np.uint32)
        _check(lib().fgbm_message_pbm(self.h, message.encode(), out.ctypes.data_as(C.c_void_p), C.byref(bins)), "fgbm_message_pbm")
        return out

    def message_variable(self, message: str, var: str, dtype, n: int) -> np.ndarray:
        out = np.empty(n, dtype=dtype)
        _check(lib().fgbm_message_variable(self.h, message.encode(), var.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes),
               "fgbm_message_variable")
        return out

    def circles_step_host(self, x, y, z, drift, steps, out):
        """bench.py e2e path: host buffers in, host buffers out (copies inside the call)."""
        n = len(x)
        _check(lib().fgbm_circles_step_host(self.h, n, x.ctypes.data, y.ctypes.data, z.ctypes.data, drift.ctypes.data, steps,
                                            out["x"].ctypes.data, out["y"].ctypes.data, out["z"].ctypes.data,
                                            out["drift"].ctypes.data, out["id"].ctypes.data), "fgbm_circles_step_host")
