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


def _flat_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the staging layout of one list (message {id,x,y,z} plus an odd-sized variable): one flat buffer per side,
        # variables back to back (16-byte aligned), count word behind them; exchanged with ONE send/recv per neighbour
        layout = [("id", 4), ("v3", 12), ("x", 4)]
        cap = 37
        b = slab._Buffers(layout, cap, "cpu")
        assert b.count_off % 16 == 0 and b.nbytes == b.count_off + 16
        for side in ("lo", "hi"):
            offs = [v.data_ptr() - b.send_flat[side].data_ptr() for v in b.send[side]]
            assert offs[0] == 0 and all(o % 16 == 0 for o in offs) and offs == sorted(offs)
            assert [v.numel() for v in b.send[side]] == [cap * bytes_ for _, bytes_ in layout]
            assert b.send_count[side].data_ptr() - b.send_flat[side].data_ptr() == b.count_off
            for k, v in enumerate(b.send[side]):
                v.fill_(16 * rank + (0 if side == "lo" else 8) + k + 1)
            b.send_count[side].fill_(1000 * rank + (1 if side == "lo" else 2))
        slab.exchange_with_neighbours([b.send_flat["lo"]], [b.send_flat["hi"]], [b.recv_flat["lo"]], [b.recv_flat["hi"]], rank, world)
        ok = True
        if rank > 0:
            ok &= int(b.recv_counts["lo"]) == 1000 * (rank - 1) + 2
            ok &= all(bool((v == 16 * (rank - 1) + 8 + k + 1).all()) for k, v in enumerate(b.recv["lo"]))
        if rank < world - 1:
            ok &= int(b.recv_counts["hi"]) == 1000 * (rank + 1) + 1
            ok &= all(bool((v == 16 * (rank + 1) + k + 1).all()) for k, v in enumerate(b.recv["hi"]))
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


def test_flat_staging_buffers_exchange_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_flat_worker, args=(3, _free_port(), out), nprocs=3, join=True)
    assert dict(out) == {r: 1 for r in range(3)}
