"""Multi-GPU z-slab decomposition of a spatial FLAME GPU model (no reference
counterpart: a reference simulation never spans GPUs).

One process per GPU (torch.distributed, NCCL over NVLink).  Rank r owns the bin planes
[z0, z1) of the slowest grid axis and stores one ghost plane on either side.  Per step:

  layer(s) that OUTPUT the spatial list          (local agents only)
  halo:      messages of plane z0 -> rank r-1, of plane z1-1 -> rank r+1 (fgb_plane_flags + fgb_compact
             pack them; one fixed-capacity NCCL send per neighbour, counts travel as device words),
             received ghosts are appended to the local list with a device-side count
  layer(s) that READ the list                    (PBM build over own + ghost planes, bin arithmetic
                                                  identical to the single-GPU build)
  migration: agents whose new plane left [z0, z1) are packed, removed, sent to the neighbour and appended

Nothing returns to the host inside a step; the launch bounds are tightened from the counts of the previous
step, read back asynchronously (CUDASimulation::endStepPipelined).
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


# ---- host logic (pure; covered by CPU tests) -----------------------------------------------------
def slab_planes(grid_planes: int, world: int, rank: int) -> Tuple[int, int]:
    """Planes [z0, z1) owned by `rank`: contiguous, as even as possible, every rank non-empty."""
    if world > grid_planes:
        raise ValueError(f"{world} ranks but only {grid_planes} bin planes")
    z0 = (grid_planes * rank) // world
    z1 = (grid_planes * (rank + 1)) // world
    return z0, z1


def slab_window(grid_planes: int, world: int, rank: int) -> Tuple[int, int]:
    """(first plane stored, number of planes stored) = own planes plus one ghost plane per side."""
    z0, z1 = slab_planes(grid_planes, world, rank)
    w0, w1 = max(z0 - 1, 0), min(z1 + 1, grid_planes)
    return w0, w1 - w0


def owner_of_plane(grid_planes: int, world: int, plane: int) -> int:
    for r in range(world):
        z0, z1 = slab_planes(grid_planes, world, r)
        if z0 <= plane < z1:
            return r
    raise ValueError(plane)


def exchange_with_neighbours(send_lo: Optional[Sequence[torch.Tensor]], send_hi: Optional[Sequence[torch.Tensor]],
                             recv_lo: Optional[Sequence[torch.Tensor]], recv_hi: Optional[Sequence[torch.Tensor]],
                             rank: int, world: int, group=None):
    """Pairwise exchange with rank-1 ("lo") and rank+1 ("hi"): one batched isend/irecv group
    (ncclGroupStart/End underneath).  Works on CUDA tensors (NCCL) and CPU tensors (gloo)."""
    ops = []
    if rank > 0:
        for t in send_lo or []:
            ops.append(dist.P2POp(dist.isend, t, rank - 1, group))
        for t in recv_lo or []:
            ops.append(dist.P2POp(dist.irecv, t, rank - 1, group))
    if rank < world - 1:
        for t in send_hi or []:
            ops.append(dist.P2POp(dist.isend, t, rank + 1, group))
        for t in recv_hi or []:
            ops.append(dist.P2POp(dist.irecv, t, rank + 1, group))
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()


# ---- device driver ----------------------------------------------------------------------------------
class _Buffers:
    """Fixed-capacity device staging for one list.  Per side ONE flat byte buffer: the variables back to back
    (capacity items each, 16-byte aligned) followed by the count word, so that an exchange is a single send and a
    single receive per neighbour (NCCL's per-operation cost, not bytes, is what an exchange of this size pays)."""

    def __init__(self, layout: List[Tuple[str, int]], capacity: int, device):
        self.layout = layout
        self.capacity = capacity
        offs, o = [], 0
        for _, b in layout:
            offs.append(o)
            o += (capacity * b + 15) // 16 * 16
        self.count_off = o
        self.nbytes = o + 16

        def mk():
            flat = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
            views = [flat[off:off + capacity * b] for off, (_, b) in zip(offs, layout)]
            count = flat[self.count_off:self.count_off + 4].view(torch.int32)
            return flat, views, count

        self.send_flat, self.send, self.send_count = {}, {}, {}
        self.recv_flat, self.recv, self.recv_counts = {}, {}, {}
        for side in ("lo", "hi"):
            self.send_flat[side], self.send[side], self.send_count[side] = mk()
            self.recv_flat[side], self.recv[side], self.recv_counts[side] = mk()

    def ptrs(self, tensors):
        return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class SlabSimulation:
    """A spatial two-phase model (one output layer, one reading layer; Circles, Boids) on `world` GPUs."""

    def __init__(self, model: str, agent: str, message: str, rank: int, world: int, device: int, grid_planes: int,
                 halo_capacity: int, migrate_capacity: int, output_layers=(0, 1), read_layers=(1, 2), **params):
        from . import sim as fsim

        self.rank, self.world, self.device = rank, world, device
        self.agent, self.message = agent, message
        self.grid_planes = grid_planes
        self.z0, self.z1 = slab_planes(grid_planes, world, rank)
        w0, wc = slab_window(grid_planes, world, rank)
        self.sim = fsim.Simulation(model, device=device, win_begin=w0, win_count=wc, graphs=0, **params)
        self.lib = fsim.lib()
        self._declare()
        self.output_layers, self.read_layers = output_layers, read_layers
        self.dev = torch.device(f"cuda:{device}")
        self.msg_buf = _Buffers(self._layout(True, message), halo_capacity, self.dev)
        self.agent_buf = _Buffers(self._layout(False, agent), migrate_capacity, self.dev)
        self.stream = torch.cuda.ExternalStream(self.sim.stream, device=self.dev)
        # the halo exchange (pack, NCCL, append, PBM build) runs on the simulation's exchange stream while the main
        # stream already sorts the agents of the reading layer
        self.xstream = torch.cuda.ExternalStream(int(self.lib.fgbm_exchange_stream(self.sim.h)), device=self.dev)
        self.overflow = False
        # FGB_SLAB_PROFILE=1: CUDA events between the phases of every step (phase_report())
        import os
        self._prof = [] if os.environ.get("FGB_SLAB_PROFILE") else None

    def _mark(self, marks, name, stream=None):
        if marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream or self.stream)
            marks.append((name, e))

    def phase_report(self):
        """{phase: mean microseconds} over the profiled steps (FGB_SLAB_PROFILE=1)."""
        if not self._prof:
            return {}
        torch.cuda.synchronize()
        acc = {}
        for marks in self._prof[len(self._prof) // 4:]:  # skip the first quarter (NCCL setup, first-touch allocations)
            for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
                acc.setdefault(name, []).append(e0.elapsed_time(e1) * 1e3)
        return {k: float(np.mean(v)) for k, v in acc.items()}

    def _declare(self):
        L = self.lib
        L.fgbm_run_layers.argtypes = [C.c_void_p, C.c_uint, C.c_uint]
        L.fgbm_end_step.argtypes = [C.c_void_p]
        L.fgbm_refresh_bounds.argtypes = [C.c_void_p]
        L.fgbm_end_step_pipelined.argtypes = [C.c_void_p]
        L.fgbm_begin_exchange.argtypes = [C.c_void_p]
        L.fgbm_end_exchange.argtypes = [C.c_void_p, C.c_char_p]
        L.fgbm_exchange_stream.argtypes = [C.c_void_p]
        L.fgbm_exchange_stream.restype = C.c_void_p
        L.fgbm_list_layout.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_size_t]
        L.fgbm_slab_pack.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_uint, C.c_int, C.c_void_p, C.c_void_p]
        L.fgbm_list_append.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_uint, C.c_void_p, C.c_void_p]

    def _check(self, rc, where):
        if rc != 0:
            raise RuntimeError(f"{where}: {self.lib.fgbm_last_error().decode()}")

    def _layout(self, is_message: bool, name: str):
        buf = C.create_string_buffer(1 << 14)
        self._check(self.lib.fgbm_list_layout(self.sim.h, int(is_message), name.encode(), buf, len(buf)), "fgbm_list_layout")
        return [(n, int(b)) for n, b in json.loads(buf.value.decode())]

    # one exchange: pack (lo/hi) -> NCCL -> append
    def _exchange(self, is_message: bool, name: str, bufs: _Buffers, lo: int, hi: int, remove: bool, marks=None, tag="",
                  stream=None):
        stream = stream or self.stream
        has_lo, has_hi = self.rank > 0, self.rank < self.world - 1
        self._check(self.lib.fgbm_slab_pack(self.sim.h, int(is_message), name.encode(), self.message.encode(), lo, hi,
                                            bufs.ptrs(bufs.send["lo"]) if has_lo else None,
                                            bufs.ptrs(bufs.send["hi"]) if has_hi else None, bufs.capacity, int(remove),
                                            C.c_void_p(bufs.send_count["lo"].data_ptr()),
                                            C.c_void_p(bufs.send_count["hi"].data_ptr())), "fgbm_slab_pack")
        self._mark(marks, tag + "_pack", stream)
        with torch.cuda.stream(stream):  # NCCL orders itself after this stream and vice versa
            exchange_with_neighbours([bufs.send_flat["lo"]], [bufs.send_flat["hi"]], [bufs.recv_flat["lo"]], [bufs.recv_flat["hi"]],
                                     self.rank, self.world)
        self._mark(marks, tag + "_nccl", stream)
        for side, present in (("lo", has_lo), ("hi", has_hi)):
            if present:
                self._check(self.lib.fgbm_list_append(self.sim.h, int(is_message), name.encode(), bufs.capacity,
                                                      C.c_void_p(bufs.recv_counts[side].data_ptr()), bufs.ptrs(bufs.recv[side])),
                            "fgbm_list_append")
        self._mark(marks, tag + "_append", stream)

    def step(self):
        s = self.sim
        marks = [] if self._prof is not None else None
        self._mark(marks, "start")
        self._check(self.lib.fgbm_run_layers(s.h, *self.output_layers), "fgbm_run_layers")
        self._mark(marks, "output_layers")
        # halo: plane z0 goes down, plane z1-1 goes up; on the exchange stream, concurrently with the agent sort
        self._check(self.lib.fgbm_begin_exchange(s.h), "fgbm_begin_exchange")
        self._exchange(True, self.message, self.msg_buf, self.z0 + 1, self.z1 - 1, remove=False, marks=marks, tag="halo",
                       stream=self.xstream)
        self._check(self.lib.fgbm_end_exchange(s.h, self.message.encode()), "fgbm_end_exchange")
        self._check(self.lib.fgbm_run_layers(s.h, *self.read_layers), "fgbm_run_layers")
        self._mark(marks, "read_layers")
        # migration: agents now below z0 go down, at or above z1 go up
        self._exchange(False, self.agent, self.agent_buf, self.z0, self.z1, remove=True, marks=marks, tag="migrate")
        # end of step + pipelined count refresh: the host consumes the counts of the PREVIOUS step while the
        # device runs this one, so it never drains the GPU
        self._check(self.lib.fgbm_end_step_pipelined(s.h), "fgbm_end_step_pipelined")
        self._mark(marks, "end_step")
        if marks is not None:
            self._prof.append(marks)

    def check_overflow(self):
        """Counts that exceeded the staging capacity mean lost items: fail loudly."""
        for bufs in (self.msg_buf, self.agent_buf):
            worst = max(int(bufs.send_count["lo"].item()), int(bufs.send_count["hi"].item()))
            if worst > bufs.capacity:
                raise RuntimeError(f"slab staging overflow: {worst} items > capacity {bufs.capacity}")

    def close(self):
        self.sim.close()
