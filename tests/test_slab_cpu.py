"""Host-side logic of the multi-GPU slab decomposition, on CPU: partition arithmetic and the pairwise
neighbour exchange over torch.distributed with the gloo backend, world_size 2 and 3."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flamegpu2_b200 import slab


def test_slab_planes_cover_the_grid_exactly():
    for planes in (8, 50, 51, 256):
        for world in (1, 2, 3, 4, 8):
            if world > planes:
                continue
            covered = []
            for r in range(world):
                z0, z1 = slab.slab_planes(planes, world, r)
                assert z1 > z0
                covered += list(range(z0, z1))
                w0, wc = slab.slab_window(planes, world, r)
                assert w0 == max(z0 - 1, 0) and w0 + wc == min(z1 + 1, planes)
                for p in range(z0, z1):
                    assert slab.owner_of_plane(planes, world, p) == r
            assert covered == list(range(planes))
    with pytest.raises(ValueError):
        slab.slab_planes(2, 4, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank sends a payload tagged with its rank and a count word, like the halo exchange does
        send_lo = [torch.full((4,), 10 * rank + 1, dtype=torch.int32), torch.tensor([rank * 100 + 1], dtype=torch.int32)]
        send_hi = [torch.full((4,), 10 * rank + 2, dtype=torch.int32), torch.tensor([rank * 100 + 2], dtype=torch.int32)]
        recv_lo = [torch.zeros(4, dtype=torch.int32), torch.zeros(1, dtype=torch.int32)]
        recv_hi = [torch.zeros(4, dtype=torch.int32), torch.zeros(1, dtype=torch.int32)]
        slab.exchange_with_neighbours(send_lo, send_hi, recv_lo, recv_hi, rank, world)
        ok = True
        if rank > 0:  # from rank-1 we receive what it sent "hi"
            ok &= bool((recv_lo[0] == 10 * (rank - 1) + 2).all()) and int(recv_lo[1]) == (rank - 1) * 100 + 2
        else:
            ok &= not recv_lo[0].any()
        if rank < world - 1:  # from rank+1 we receive what it sent "lo"
            ok &= bool((recv_hi[0] == 10 * (rank + 1) + 1).all()) and int(recv_hi[1]) == (rank + 1) * 100 + 1
        else:
            ok &= not recv_hi[0].any()
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_neighbour_exchange_gloo(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {r: 1 for r in range(world)}


def _handle_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # what SlabSimulation does once at start-up: every rank contributes its 64-byte staging handle and receives all
        # of them in rank order (the C++ side maps rank r's arena from bytes [64 r, 64 r + 64))
        blob = bytes([(17 * rank + k) % 251 for k in range(64)])
        everyone = slab.all_gather_bytes(blob, world)
        ok = len(everyone) == 64 * world
        for r in range(world):
            ok &= everyone[64 * r:64 * (r + 1)] == bytes([(17 * r + k) % 251 for k in range(64)])
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_staging_handles_all_gather_gloo(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_handle_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {r: 1 for r in range(world)}


def test_slab_arithmetic_matches_the_cpp_layer():
    # CUDASimulation::configureSlabs computes z0 = planes * rank / world (integer division); the host helper must agree
    for planes in (8, 50, 51, 256):
        for world in (1, 2, 3, 4, 8):
            for r in range(world):
                assert slab.slab_planes(planes, world, r) == ((planes * r) // world, (planes * (r + 1)) // world)
