#!/usr/bin/env python
"""bench.py -- agent-steps/s of the Circles-3D spatial benchmark (BASELINE.json) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (sm_100a hot path)
  python bench.py --impl reference [--gpus N] [--steps K] ...     # CPU rendition of the same step

One "step" is one whole CUDASimulation::step() of the Circles model (output_message -> automatic
agent sort -> PBM buildIndex -> move) over all agents.  N=1 runs BASELINE.json configs[1]
(1 M agents, [0,100)^3, radius 2 => 50^3 bins, 8 agents/bin); N>1 is weak scaling with the same
number of agents per GPU: the global box is the single-GPU box stacked N times along z, decomposed into z-slabs
with halo and migration exchange over NCCL every step (flamegpu2_b200/slab.py).  Prints ONE JSON line on rank 0.

Besides the contract's keys the N=1 line carries: `roofline` (buildIndex, algorithmic bytes / its measured
duration against the measured HBM peak), `phases_us` (per-phase device times from an eager profiled pass), `e2e`
(host buffers in and out through the C ABI), `cpu_baseline` (the OpenMP port on the host cores), `reference_cuda`
(the reference's own CUDA build on the same GPU, when oracle/_ref/ref_sim exists) and `opt_in` (the radius-filtered
iterator, off by default, over the same number of steps).

Timing: the agent state is resident in HBM; every step is timed with CUDA events recorded on the
simulation's own stream (where its graph is launched); the whole working set (~80 MB) fits the
126 MB L2, so a 256 MiB buffer is written between steps, outside the events, to flush it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

RADIUS = 2.0
REPULSE = 0.05
S_AGENT = 24  # _auto_sort_bin_index,_id,drift,x,y,z  
S_MSG = 16    # id,x,y,z


def env_extent(n):
    """fixed density 1 agent / unit^3 (the example: ENV_MAX = floor(cbrt(N)))"""
    return float(np.floor(np.cbrt(float(n)) + 1e-6))


def population(n, L, seed):
    rng = np.random.default_rng(seed)
    return [rng.uniform(0.0, L, n).astype(np.float32) for _ in range(3)]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for k in dir(nv):
            if k.startswith("nvmlClocksThrottleReason") or k.startswith("nvmlClocksEventReason"):
                v = getattr(nv, k)
                if isinstance(v, int) and v:
                    names[v] = k.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", "")
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, nm in names.items():
                    if r & bit and bit & (bit - 1) == 0:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1.0)
        idle = {"GpuIdle", "None", "ApplicationsClocksSetting"}
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "reasons": sorted(r for r in self.reasons if r not in idle)}


def cpu_circles(n, L, steps, seed=0):
    """The CPU rendition (oracle port, OpenMP over all host threads) of the same step."""
    import oracle_py as orc

    g = orc.Grid(3, (0, 0, 0), (L, L, L), RADIUS)
    x, y, z = population(n, L, seed)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    d = np.zeros(n, np.float32)
    ids, x, y, z, d, _ = g.circles_step(ids, x, y, z, d, REPULSE)  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    for _ in range(steps):
        ids, x, y, z, d, _ = g.circles_step(ids, x, y, z, d, REPULSE)
    dt = time.perf_counter() - t0
    return n * steps / dt, dt / steps, orc.num_threads()


def run_reference(args):
    """--impl reference: FLAME GPU has no CPU path, so the reference arm is the OpenMP C restatement
    of the reference's algorithm (oracle/, kind "port") on this box's host cores, full workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.agents_per_gpu
    L = env_extent(n)
    steps = max(1, min(args.steps, 40))
    value, sec, cores = cpu_circles(n, L, steps)
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Circles-3D {n} agents, [0,{L:g})^3, radius {RADIUS:g}, whole step on host cores",
                   "agents": n, "note": "FLAME GPU 2 has no CPU implementation; this is the OpenMP C restatement of its step (oracle/)"},
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps of the full {n}-agent workload"},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def reference_cuda(n, L, steps=30, warmup=5):
    """The reference's own CUDA build (oracle/_ref/ref_sim) on the same GPU and workload, if present."""
    try:
        import tempfile

        import fgbs

        if not fgbs.have_ref():
            return None
        x, y, z = population(n, L, 0)
        with tempfile.TemporaryDirectory() as td:
            inp = os.path.join(td, "in.bin")
            fgbs.write_state(inp, {"x": x, "y": y, "z": z})
            js = fgbs.run_ref("circles", {"env_max": L, "radius": RADIUS}, inp, os.path.join(td, "ref"), steps=steps, warmup=warmup)
        t = np.array(js["step_seconds"])
        return {"value": float(n / np.median(t)), "unit": "agent-steps/s", "ms_per_step": float(np.median(t) * 1e3),
                "what": "FLAME GPU 2 v2.0.0-rc.5 built unmodified for sm_100a (seatbelts off), per-step times from getElapsedTimeSteps()"}
    except Exception as e:  # reported, never fatal
        return {"error": str(e)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / reference_cuda / e2e legs")
    ap.add_argument("--stable", type=int, default=0)
    ap.add_argument("--true3d-sort", type=int, default=0)
    ap.add_argument("--overlap", type=int, default=1, help="1: build the PBM on a second stream while the agents are sorted")
    ap.add_argument("--block", type=int, default=128, help="threads per block of the agent function kernels")
    ap.add_argument("--tile-order", type=int, default=1, help="1: tile-local execution order after the auto sort")
    ap.add_argument("--iter-mode", type=int, default=0, help="0 reference visit order (default), 1 radius-filtered lock-step walk (opt-in)")
    ap.add_argument("--bin-order", type=int, default=1, help="run message-reading functions in bin order (b200 extension)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: flamegpu2_b200 has no CPU fallback"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    from flamegpu2_b200 import sim as fsim

    n = args.agents_per_gpu
    L = env_extent(n)
    bins = int(np.ceil(L / RADIUS)) ** 3
    flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.int32, device=f"cuda:{local}")
    if world == 1:
        x, y, z = population(n, L, seed=rank)
        s = fsim.Simulation("circles", device=local, env_max=L, radius=RADIUS, repulse=REPULSE, timing=1, stable=args.stable,
                            true3d_sort=args.true3d_sort, bin_order=args.bin_order, iter_mode=args.iter_mode, overlap=args.overlap, tile_order=args.tile_order, block=args.block)
        s.set_population("Circle", {"x": x, "y": y, "z": z})
        stream = torch.cuda.ExternalStream(s.stream, device=f"cuda:{local}")
        slab_sim = None

        def one_step():
            with torch.cuda.stream(stream):
                flush.add_(1)  # evicts the step's working set from the 126 MB L2; outside the timed events
            s.step(1)
    else:
        # weak scaling: the global box is [0,L)^2 x [0, L*world): `world` slabs of n agents stacked along z,
        # halo messages and migrating agents exchanged over NCCL every step (flamegpu2_b200/slab.py)
        from flamegpu2_b200 import slab

        planes_per_rank = int(np.ceil(L / RADIUS))
        planes = planes_per_rank * world
        Lz = float(planes * RADIUS)
        halo_cap = int(3 * n // planes_per_rank + 8192)      # a boundary plane holds ~n/planes_per_rank messages (3x: clustering)
        mig_cap = int(n // planes_per_rank // 2 + 4096)       # a few percent of a plane changes slab per step
        slab_sim = slab.SlabSimulation("circles", "Circle", "location", rank, world, local, planes, halo_capacity=halo_cap,
                                       migrate_capacity=mig_cap, env_max=L, env_max_z=Lz, radius=RADIUS, repulse=REPULSE,
                                       stable=args.stable, true3d_sort=args.true3d_sort, bin_order=args.bin_order, iter_mode=args.iter_mode, overlap=args.overlap, tile_order=args.tile_order, block=args.block)
        s = slab_sim.sim
        rng = np.random.default_rng(rank)
        z_lo, z_hi = slab_sim.z0 * RADIUS, slab_sim.z1 * RADIUS
        x = rng.uniform(0.0, L, n).astype(np.float32)
        y = rng.uniform(0.0, L, n).astype(np.float32)
        z = rng.uniform(z_lo, np.nextafter(np.float32(z_hi), np.float32(0)), n).astype(np.float32)
        ids = (np.arange(n, dtype=np.uint32) + np.uint32(rank * n + 1))
        s.set_population("Circle", {"x": x, "y": y, "z": z, "_id": ids})
        stream = torch.cuda.ExternalStream(s.stream, device=f"cuda:{local}")

        def one_step():
            slab_sim.step()

    for _ in range(args.warmup):
        one_step()
    s.sync()
    s.step_times()  # drop warm-up timings
    launches0 = s.launches
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        one_step()
    ev1.record(stream)
    s.sync()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    if dist is not None:
        dist.barrier()
    if world == 1:
        st_ = s.step_times()
        if os.environ.get("FGB_BENCH_DEBUG"):
            print("step_times_ms", np.round(st_ * 1e3, 3).tolist(), file=sys.stderr)
        dev_total = float(st_.sum())  # per-step events: the L2 flush between steps is excluded
    else:
        dev_total = ev0.elapsed_time(ev1) * 1e-3  # device time of the whole K-step region on the simulation stream
        slab_sim.check_overflow()
    launches = s.launches - launches0
    if dist is not None:
        t = torch.tensor([dev_total, wall], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_total, wall = float(t[0]), float(t[1])
    value = n * world * args.steps / dev_total
    line = {
        "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"Circles-3D, {n} agents per GPU, [0,{L:g})^3, radius {RADIUS:g} ({bins} bins, ~8 agents/bin), "
                        f"whole CUDASimulation::step() (output_message, auto agent sort, PBM buildIndex, move)",
            "agents_per_gpu": n, "bins": bins, "graphs": s.graphs, "bin_order_execution": bool(args.bin_order), "iterator_mode": args.iter_mode,
            "overlap_index_build": bool(args.overlap), "tile_local_exec_order": bool(args.tile_order),
            "l2": "flushed between steps (256 MiB write outside the timed events)" if world == 1 else
                  "not flushed (exchange-synchronised steps; per-GPU working set ~80 MB)",
            "timing": "sum of per-step CUDA-event times on the simulation stream" if world == 1 else
                      "CUDA events around the K-step region on the simulation stream (NCCL waits included), max over ranks",
            "wall_ms_per_step": wall / args.steps * 1e3,
            "multi_gpu": (f"z-slab decomposition over {world} GPUs, box [0,{L:g})^2 x [0,{L * world:g}); halo messages + "
                          "migrating agents over NCCL send/recv every step") if world > 1 else "single GPU",
        },
        "gpu_launches": int(launches),
        "clocks": clk,
    }
    if rank == 0 and world > 1:
        if slab_sim is not None and os.environ.get("FGB_SLAB_PROFILE"):
            rep = slab_sim.phase_report()  # includes the warm-up steps; diagnostic only
            line["slab_phases_us"] = rep
        print(json.dumps(line), flush=True)
    if rank == 0 and world == 1:
        peak, peak_src = measured_peak()
        # -- per-phase device times from a profiled (eager, event-bracketed) pass of the same workload
        p = fsim.Simulation("circles", device=local, env_max=L, radius=RADIUS, repulse=REPULSE, profile=1, stable=args.stable,
                            true3d_sort=args.true3d_sort, bin_order=args.bin_order, iter_mode=args.iter_mode, overlap=args.overlap, tile_order=args.tile_order, block=args.block)
        p.set_population("Circle", {"x": x, "y": y, "z": z})
        pstream = torch.cuda.ExternalStream(p.stream, device=f"cuda:{local}")
        for i in range(args.warmup + 30):
            with torch.cuda.stream(pstream):
                flush.add_(1)
            p.step(1)
            if i == args.warmup - 1:
                p.profile()
        prof = p.profile()
        p.close()
        phases = {k: v[0] / v[1] * 1e3 for k, v in prof.items() if v[1]}  # microseconds per call
        bi_us = phases.get("build_index")
        alg = n * 2 * S_MSG + 4 * (bins + 1)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get("build_index", {}).get(str(n))
        except Exception:
            pass
        if bi_us:
            line["roofline"] = {"bound": "hbm", "kernel": "buildIndex (k_bin_hist + k_exclusive_scan + k_bin_scatter)",
                                "achieved": alg / bi_us / 1e3, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                                "frac": alg / bi_us / 1e3 / peak, "traffic": traffic, "algorithmic_bytes": alg,
                                "us_per_launch": bi_us}
        line["phases_us"] = phases
        if not args.no_extras and world == 1:
            # -- end to end through the C ABI with HOST buffers (copies inside the timed region)
            k_e2e = min(args.steps, 30)
            hx, hy, hz, hd = (torch.from_numpy(a.copy()).pin_memory().numpy() for a in (x, y, z, np.zeros(n, np.float32)))
            out = {k: torch.empty(n, dtype=torch.float32).pin_memory().numpy() for k in ("x", "y", "z", "drift")}
            out["id"] = torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
            for _ in range(3):
                s.circles_step_host(hx, hy, hz, hd, 1, out)
            t0 = time.perf_counter()
            for _ in range(k_e2e):
                s.circles_step_host(hx, hy, hz, hd, 1, out)
                hx, hy, hz, hd = out["x"], out["y"], out["z"], out["drift"]
            e2e_t = time.perf_counter() - t0
            line["e2e"] = {"value": n * k_e2e / e2e_t, "unit": "agent-steps/s", "h2d_bytes_per_step": 16 * n,
                           "d2h_bytes_per_step": 20 * n, "steps": k_e2e,
                           "call": "fgbm_circles_step_host (setPopulationData + step + getPopulationData)"}
            # -- CPU baseline: the OpenMP port on this box's host cores, bounded sample
            v1, sec1, cores = cpu_circles(n, L, 1)
            ksteps = int(max(1, min(30, 10.0 / max(sec1, 1e-3))))
            v, sec, cores = cpu_circles(n, L, ksteps)
            line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{ksteps} steps of the same {n}-agent Circles workload"}
            line["reference_cuda"] = reference_cuda(n, L)
            if args.iter_mode == 0:
                # -- opt-in, not the headline: the radius-filtered iterator (CUDAConfig().spatialIterationMode = 1),
                #    bit-identical Circles results (tests/test_sim_gpu.py), same workload and timing method
                f = fsim.Simulation("circles", device=local, env_max=L, radius=RADIUS, repulse=REPULSE, timing=1, stable=args.stable,
                                    true3d_sort=args.true3d_sort, bin_order=args.bin_order, iter_mode=1)
                f.set_population("Circle", {"x": x, "y": y, "z": z})
                fstream = torch.cuda.ExternalStream(f.stream, device=f"cuda:{local}")
                k_f = args.steps  # same horizon as the headline: the gain shrinks as Circles clusters
                for i in range(args.warmup + k_f):
                    with torch.cuda.stream(fstream):
                        flush.add_(1)
                    f.step(1)
                    if i == args.warmup - 1:
                        f.sync()
                        f.step_times()
                f.sync()
                f_times = f.step_times()
                ft = float(f_times.sum())
                f.close()
                ftimes = f_times
                line["opt_in"] = {"radius_filtered_iterator": {"value": n * k_f / ft, "unit": "agent-steps/s",
                                                               "ms_per_step": ft / k_f * 1e3, "steps": k_f,
                                                               "ms_first_30_steps": float(ftimes[:30].mean() * 1e3),
                                                               "ms_last_30_steps": float(ftimes[-30:].mean() * 1e3)}}
        print(json.dumps(line), flush=True)
    s.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
