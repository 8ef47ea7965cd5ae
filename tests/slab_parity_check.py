"""Parity of the z-slab decomposition against a single-GPU run of the same global Circles domain.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/slab_parity_check.py

run_check() is also what `bench.py --gpus N` runs before its timed region.  ONE teacher-forced step from a common
seeded state (both sides start from identical agents, so nothing is amplified by the dynamics):
  * the slab PBMs, restricted to the planes each rank owns and concatenated, ARE the single-GPU PBM, bit-exactly
    (SURVEY.md 8e), and every rank's ghost planes hold exactly the neighbours' boundary planes;
  * every PBM bin of an owned plane holds the same message ids as the single-GPU list (as a multiset);
  * after migration every agent sits on the rank that owns its plane, no id is lost or duplicated;
  * ids and integer state (_auto_sort_bin_index) are bit-exact per agent, floats agree within rtol 1e-5 / atol 2e-6
    (summation order over ~33 neighbours).
A free-running multi-step variant (SLAB_STEPS > 1) is kept as a smoke test of sustained migration with loose tolerances."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flamegpu2_b200 import sim as fsim  # noqa: E402
from flamegpu2_b200 import slab  # noqa: E402

RTOL, ATOL = 1e-5, 2e-6


def _bin_multisets(pbm, ids):
    bins = np.repeat(np.arange(len(pbm) - 1), np.diff(pbm.astype(np.int64)))
    order = np.lexsort((ids, bins))
    return ids[order]


def run_check(rank, world, local, n_per_rank=100_000, planes_per_rank=23, radius=2.0, repulse=0.05, seed=123, free_steps=0):
    """Returns a dict (rank 0: the verdict, other ranks: {"ok": <same flag>}).  Collective: every rank must call it."""
    L = float(np.floor(np.sqrt(n_per_rank / (planes_per_rank * radius))))  # density 1 agent / unit^3
    planes = planes_per_rank * world
    Lz = float(planes * radius)
    n = n_per_rank * world
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, L, n).astype(np.float32)
    y = rng.uniform(0, L, n).astype(np.float32)
    z = rng.uniform(0, np.nextafter(np.float32(Lz), np.float32(0)), n).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    params = dict(env_max=L, env_max_z=Lz, radius=radius, repulse=repulse)
    gx = int(np.ceil(L / radius))
    gxy = gx * gx
    z0, z1 = slab.slab_planes(planes, world, rank)
    w0, wc = slab.slab_window(planes, world, rank)
    plane0 = np.clip(np.floor(z / np.float32(radius)), 0, planes - 1).astype(np.int64)
    mine = (plane0 >= z0) & (plane0 < z1)
    cap = int(3 * n_per_rank // planes_per_rank + 4096)
    s = slab.SlabSimulation("circles", "Circle", "location", rank, world, local, planes, halo_capacity=cap, migrate_capacity=cap, **params)
    s.sim.set_population("Circle", {"x": x[mine], "y": y[mine], "z": z[mine], "_id": ids[mine]})
    steps = max(1, free_steps)
    s.step(steps)
    s.sim.sync()
    err = s.error_bits()
    got = {k: s.sim.get("Circle", k, np.float32) for k in ("x", "y", "z", "drift")}
    got_id = s.sim.get("Circle", "_id", np.uint32)
    got_key = s.sim.get("Circle", "_auto_sort_bin_index", np.uint32)
    pbm = s.sim.message_pbm("location")
    mids = s.sim.message_variable("location", "id", np.uint32, int(pbm[-1]))
    pl = np.clip(np.floor(got["z"] / np.float32(radius)), 0, planes - 1).astype(np.int64)
    on_owner = bool(np.all((pl >= z0) & (pl < z1)))
    # what the example's Validation step function saw in the last step: the GLOBAL sum of "drift" (all-reduced over the slabs)
    import ctypes as C

    lib = fsim.lib()
    lib.fgbm_circles_validation.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    seen_total = C.c_double()
    lib.fgbm_circles_validation(C.byref(seen_total), None, None)
    mine_out = {"ids": got_id, "key": got_key, "pbm": pbm, "mids": mids, "on_owner": on_owner, "err": err, "z0": z0, "z1": z1, "w0": w0, "wc": wc,
                "validation_total": float(seen_total.value)}
    mine_out.update(got)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine_out)
    s.close()
    verdict = {"ok": True}
    if rank == 0:
        ref = fsim.Simulation("circles", device=local, **params)
        ref.set_population("Circle", {"x": x, "y": y, "z": z, "_id": ids})
        ref.step(steps)
        rid = ref.get("Circle", "_id", np.uint32)
        rkey = ref.get("Circle", "_auto_sort_bin_index", np.uint32)
        rv = {k: ref.get("Circle", k, np.float32) for k in ("x", "y", "z", "drift")}
        rpbm = ref.message_pbm("location")
        rmids = ref.message_variable("location", "id", np.uint32, n)
        ref.close()
        fails = []
        if any(g["err"] for g in gathered):
            fails.append(f"device error bits {[g['err'] for g in gathered]}")
        if not all(g["on_owner"] for g in gathered):
            fails.append("agents outside their rank's slab after migration")
        global_drift = float(sum(np.asarray(g["drift"], np.float64).sum() for g in gathered))
        for r, g in enumerate(gathered):
            if abs(g["validation_total"] - global_drift) > 1e-5 * max(global_drift, 1e-30):
                fails.append(f"rank {r}: the Validation step function saw total drift {g['validation_total']}, the slabs hold {global_drift}")
        all_ids = np.concatenate([g["ids"] for g in gathered])
        if len(all_ids) != n or not np.array_equal(np.sort(all_ids), ids):
            fails.append(f"ids lost or duplicated ({len(all_ids)} of {n})")
        pbm_bits = bins_ok = ghosts_ok = order_ok = ints_ok = None
        worst = {}
        if free_steps == 0 and not fails:
            rcounts = np.diff(rpbm.astype(np.int64))
            # (a) concatenation of the owned planes of every slab PBM == the single-GPU PBM; ghost planes complete
            own = []
            ghosts_ok = True
            for g in gathered:
                c = np.diff(g["pbm"].astype(np.int64))
                own.append(c[(g["z0"] - g["w0"]) * gxy:(g["z1"] - g["w0"]) * gxy])
                ghosts_ok &= bool(np.array_equal(c, rcounts[g["w0"] * gxy:(g["w0"] + g["wc"]) * gxy]))
            cat = np.concatenate([[0], np.cumsum(np.concatenate(own))]).astype(np.uint32)
            pbm_bits = bool(np.array_equal(cat, rpbm))
            # (b) the same message ids in every owned bin
            ref_sorted = _bin_multisets(rpbm, rmids)
            bins_ok = True
            for g in gathered:
                lo, hi = int(rpbm[g["z0"] * gxy]), int(rpbm[g["z1"] * gxy])
                wl = g["pbm"].astype(np.int64)
                a, b = int(wl[(g["z0"] - g["w0"]) * gxy]), int(wl[(g["z1"] - g["w0"]) * gxy])
                mine_sorted = _bin_multisets(g["pbm"], g["mids"])[a:b]
                bins_ok &= bool(np.array_equal(mine_sorted, ref_sorted[lo:hi]))
            # (c) integer state per agent, list order of the agents that stayed, floats within tolerance
            pos_in_ref = np.empty(n + 1, np.int64)
            pos_in_ref[rid] = np.arange(n)
            ints_ok = order_ok = True
            owner_before = np.searchsorted(np.array([g["z1"] for g in gathered]), plane0, side="right")
            for r, g in enumerate(gathered):
                p = pos_in_ref[g["ids"]]
                ints_ok &= bool(np.array_equal(g["key"], rkey[p]))
                stayed = owner_before[g["ids"] - 1] == r
                k = int(stayed.sum())
                order_ok &= bool(stayed[:k].all()) and bool(np.all(np.diff(p[:k]) > 0))
                for v in ("x", "y", "z", "drift"):
                    rt = 1e-3 if v == "drift" else RTOL
                    bad = ~np.isclose(g[v], rv[v][p], rtol=rt, atol=ATOL)
                    worst[v] = max(worst.get(v, 0.0), float(np.max(np.abs(g[v] - rv[v][p]))) if len(p) else 0.0)
                    if bad.any():
                        fails.append(f"rank {r}: {int(bad.sum())} agents differ in {v} beyond rtol {rt} / atol {ATOL}")
            if not pbm_bits:
                fails.append("concatenated slab PBMs differ from the single-GPU PBM")
            if not ghosts_ok:
                fails.append("ghost planes differ from the neighbours' boundary planes")
            if not bins_ok:
                fails.append("per-bin message ids differ")
            if not ints_ok:
                fails.append("_auto_sort_bin_index differs")
            # (the agents that stay keep their relative order except for the few tail agents that fill the holes of the
            #  leavers -- fgb_slab_migrate_out -- so `stayers_keep_list_order` is reported, not required)
        elif not fails:
            # the last step's message lists (written through the execution order that survived the migrations of the
            # earlier steps): over the planes each rank owns, every agent id exactly once
            own_msgs = []
            for g in gathered:
                wl = g["pbm"].astype(np.int64)
                a, b = int(wl[(g["z0"] - g["w0"]) * gxy]), int(wl[(g["z1"] - g["w0"]) * gxy])
                own_msgs.append(g["mids"][a:b])
            own_msgs = np.sort(np.concatenate(own_msgs))
            if len(own_msgs) != n or not np.array_equal(own_msgs, ids):
                fails.append(f"messages of the last step: {len(own_msgs)} in owned planes for {n} agents (lost or duplicated)")
            # free-running: summation-order differences are amplified by the dynamics; almost all agents must agree
            pos_in_ref = np.empty(n + 1, np.int64)
            pos_in_ref[rid] = np.arange(n)
            for g in gathered:
                p = pos_in_ref[g["ids"]]
                for v in ("x", "y", "z"):
                    d = np.abs(g[v] - rv[v][p])
                    worst[v] = max(worst.get(v, 0.0), float(d.max()) if len(p) else 0.0)
                    if (d > 5e-4).mean() > 1e-3 or d.max() > 0.05:
                        fails.append(f"free-running {v}: {(d > 5e-4).sum()} agents off, max {d.max()}")
        moved = int((np.clip(np.floor(rv["z"] / np.float32(radius)), 0, planes - 1)[pos_in_ref[ids]] != plane0).sum())
        verdict = {"ok": not fails, "mode": "one teacher-forced step" if free_steps == 0 else f"{free_steps} free-running steps",
                   "world": world, "agents": n, "planes": planes, "agents_that_changed_plane": moved, "pbm_concatenation_bit_exact": pbm_bits,
                   "ghost_planes_complete": ghosts_ok, "bin_multisets_equal": bins_ok, "integer_state_bit_exact": ints_ok,
                   "stayers_keep_list_order": order_ok, "max_abs_diff": worst, "float_tolerance": {"rtol": RTOL, "atol": ATOL},
                   "failures": fails}
    flag = torch.tensor([1 if verdict["ok"] else 0], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    if rank != 0:
        verdict = {"ok": bool(int(flag.item()))}
    return verdict


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ok = True
    # one teacher-forced step, tight tolerances
    v = run_check(rank, world, local, n_per_rank=int(os.environ.get("SLAB_N", "100000")))
    ok &= v["ok"]
    if rank == 0:
        print("slab parity (teacher-forced):", v, "->", "OK" if v["ok"] else "FAIL", flush=True)
    # sustained migration: a large repulse makes agents really cross slab boundaries for several steps
    steps = int(os.environ.get("SLAB_STEPS", "6"))
    if steps > 1:
        v = run_check(rank, world, local, n_per_rank=24000, planes_per_rank=8, repulse=0.6, free_steps=steps)
        ok &= v["ok"]
        if rank == 0:
            print("slab parity (free-running):", v, "->", "OK" if v["ok"] else "FAIL", flush=True)
    dist.destroy_process_group()
    if rank == 0:
        print("-> OK" if ok else "-> FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
