"""Multi-GPU z-slab decomposition of a spatial FLAME GPU model (no reference
counterpart: a reference simulation never spans GPUs).

One process per GPU.  Rank r owns the bin planes [z0, z1) of the slowest grid axis and stores one ghost plane on
either side.  The exchange itself is C++/CUDA behind CUDASimulation::step() (configureSlabs): per step

  halo:      before the list is indexed for its first reader, the messages of plane z0 -> rank r-1 and of plane z1-1
             -> rank r+1: fgb_plane_flags + fgb_compact_limited write them (and their count) straight into the
             neighbour's staging buffer (peer memory over NVLink), a flag word publishes the epoch, the received ghosts
             are appended with a device-side count; the PBM build then covers own + ghost planes with bin arithmetic
             identical to the single-GPU build
  migration: at the end of the step agents whose new plane left [z0, z1) are packed to the neighbour the same way,
             removed here, and the arriving ones appended

This module keeps the pure host logic (which planes a rank owns) and the one-off plumbing: torch.distributed
all-gathers the 64-byte CUDA IPC handles of the staging arenas when a simulation is created.
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


def all_gather_bytes(blob: bytes, world: int, device=None, group=None) -> bytes:
    """Every rank contributes `blob` (same length everywhere); returns the concatenation in rank order.
    torch.distributed is only the plumbing here: CUDA tensors over NCCL, CPU tensors over gloo."""
    backend = dist.get_backend(group)
    dev = device if (backend == "nccl" and device is not None) else torch.device("cpu")
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)


# ---- device driver ----------------------------------------------------------------------------------
class SlabSimulation:
    """A spatial model decomposed into z-slabs over `world` GPUs, one process per GPU.

    The decomposition lives behind CUDASimulation::step() (include/flamegpu/simulation/CUDASimulation.h,
    configureSlabs): halo messages and migrating agents are packed straight into the neighbours' staging buffers
    (peer memory over NVLink, mapped through CUDA IPC), the whole step is captured in CUDA graphs.  This class only
    creates the simulation with its slab parameters and all-gathers the 64-byte staging handles once."""

    def __init__(self, model: str, agent: str, message: str, rank: int, world: int, device: int, grid_planes: int,
                 halo_capacity: int, migrate_capacity: int, group=None, **params):
        from . import sim as fsim

        self.rank, self.world, self.device = rank, world, device
        self.agent, self.message = agent, message
        self.grid_planes = grid_planes
        self.z0, self.z1 = slab_planes(grid_planes, world, rank)
        self.dev = torch.device(f"cuda:{device}")
        self.sim = fsim.Simulation(model, device=device, slab_rank=rank, slab_world=world, slab_message=message,
                                   halo_cap=int(halo_capacity), mig_cap=int(migrate_capacity), **params)
        self.lib = L = fsim.lib()
        L.fgbm_slab_handle_bytes.argtypes = [C.c_void_p]
        L.fgbm_slab_handle_bytes.restype = C.c_uint
        L.fgbm_slab_export.argtypes = [C.c_void_p, C.c_void_p]
        L.fgbm_slab_connect.argtypes = [C.c_void_p, C.c_void_p]
        L.fgbm_slab_planes.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.fgbm_slab_error.argtypes = [C.c_void_p, C.POINTER(C.c_uint)]
        z0, z1 = C.c_int(), C.c_int()
        self._check(L.fgbm_slab_planes(self.sim.h, C.byref(z0), C.byref(z1)), "fgbm_slab_planes")
        assert (z0.value, z1.value) == (self.z0, self.z1), "host and C++ slab arithmetic disagree"
        if world > 1:
            nbytes = int(L.fgbm_slab_handle_bytes(self.sim.h))
            buf = C.create_string_buffer(nbytes)
            self._check(L.fgbm_slab_export(self.sim.h, buf), "fgbm_slab_export")
            everyone = all_gather_bytes(buf.raw, world, self.dev, group)
            self._check(L.fgbm_slab_connect(self.sim.h, everyone), "fgbm_slab_connect")
        self.stream = torch.cuda.ExternalStream(self.sim.stream, device=self.dev)

    def _check(self, rc, where):
        if rc != 0:
            raise RuntimeError(f"{where}: {self.lib.fgbm_last_error().decode()}")

    def step(self, steps: int = 1):
        self.sim.step(steps)

    def error_bits(self) -> int:
        bits = C.c_uint()
        self._check(self.lib.fgbm_slab_error(self.sim.h, C.byref(bits)), "fgbm_slab_error")
        return int(bits.value)

    def check_overflow(self):
        """Device-side error word: 1 neighbour timeout, 2 staging overflow (lost items), 4 list outgrew its launch bound."""
        bits = self.error_bits()
        if bits:
            raise RuntimeError(f"slab exchange error bits {bits} (1 neighbour timeout, 2 staging overflow, 4 launch bound)")

    def close(self):
        self.sim.close()
