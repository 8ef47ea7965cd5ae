#!/usr/bin/env python
"""bench.py -- agent-steps/s of the Circles-3D spatial benchmark (BASELINE.json) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (sm_100a hot path)
  python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's own CUDA build (oracle/_ref/ref_sim)

One "step" is one whole CUDASimulation::step() of the Circles example (output_message -> automatic agent sort -> PBM
buildIndex -> move -> the example's Validation step function) over all agents.  N=1 runs BASELINE.json configs[1]
(1 M agents, [0,100)^3, radius 2 => 50^3 bins, 8 agents/bin).  N>1 is weak scaling with the same number of agents per
GPU: the global box is the single-GPU box stacked N times along z and decomposed into z-slabs behind
CUDASimulation::step() (halo messages and migrating agents are packed straight into the neighbour's memory over
NVLink; CUDASimulation::configureSlabs).  Prints ONE JSON line on rank 0.

Besides the contract's keys the line carries
  roofline      buildIndex in-step at the headline size (algorithmic bytes / measured duration vs the measured HBM peak)
  rooflines     the same for buildIndex / agent sort at 16.8 M agents per GPU (the north-star share) and for the
                death compaction kernel; `move` with its own bound (issue slots x lanes per instruction, from ncu)
  phases_us     per-phase device times from an eager profiled pass
  e2e           host buffers in and out through the C ABI every step (copies inside the timed region)
  cpu_baseline  the OpenMP port of the step on this box's host cores (FLAME GPU has no CPU path)
  reference_cuda  the reference's own CUDA build on the same GPU (also what `--impl reference` times)
  modes         the same workload with the strict reference iteration order for every function
  north_star    Circles-3D at 16.8 M agents per GPU (128 M on 8 GPUs: BASELINE configs[4]), weak-scaling share
  strong_scaling  the fixed 128 M-agent box on N GPUs
  parity_check  (N>1) one teacher-forced step of a 100 k-agents-per-GPU domain, slabs vs one GPU, before the timed region

Timing: the agent state is resident in HBM; every step is timed with the CUDA event pair that CUDASimulation::step()
records around it on the simulation's own stream (where its graph is launched; the analogue of the reference's
getElapsedTimeSteps()).  `value` is K steps / (sum of a rank's per-step times, max over ranks) at every N: waits for
neighbours (halo, migration, all-reduce) happen inside a step and count; what lies between two steps does not -- at N=1 the
256 MiB write that flushes the 126 MB L2 (the 1 M-agent working set would otherwise stay cache resident), at every N the
host's turn-around after the step function made it wait for the step.  The bracketed K-step region is printed next to it
(config.region_ms_per_step, config.wall_ms_per_step).
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
PY_LOOP = False                  # --py-loop: N > 1 timed steps driven from Python (A/B against the single C++ call)
NORTH_STAR_PER_GPU = 1 << 24     # 16.8 M agents per GPU
NORTH_STAR_CROSS = 512.0         # 512 x 512 cross-section, 64 deep per GPU: 8 GPUs make the 512^3 box of configs[4]


def env_extent(n):
    """fixed density 1 agent / unit^3 (the example: ENV_MAX = floor(cbrt(N)))"""
    return float(np.floor(np.cbrt(float(n)) + 1e-6))


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


# ---- workloads ---------------------------------------------------------------------------------------------------
class Box:
    """A Circles domain [0,cross)^2 x [0, depth_per_rank * world) at density 1 agent / unit^3, `world` z-slabs."""

    def __init__(self, cross, depth_per_rank, world):
        self.cross, self.world = float(cross), world
        self.planes_per_rank = int(np.ceil(depth_per_rank / RADIUS))
        self.planes = self.planes_per_rank * world
        self.depth = float(self.planes * RADIUS)
        self.n_per_rank = int(round(self.cross * self.cross * self.planes_per_rank * RADIUS))
        self.bins_per_rank = int(np.ceil(self.cross / RADIUS)) ** 2 * self.planes_per_rank

    @staticmethod
    def cube(n_per_gpu, world):
        L = env_extent(n_per_gpu)
        b = Box(L, L, world)
        b.n_per_rank = n_per_gpu
        return b

    def describe(self):
        return (f"[0,{self.cross:g})^2 x [0,{self.depth:g}), radius {RADIUS:g}, {self.bins_per_rank * self.world} bins, "
                f"{self.n_per_rank * self.world} agents (~8 per bin)")

    def population(self, rank, seed=0):
        """This rank's agents: uniform in its own slab; ids are globally unique."""
        rng = np.random.default_rng(seed + rank)
        n = self.n_per_rank
        z_lo = rank * self.planes_per_rank * RADIUS
        z_hi = (rank + 1) * self.planes_per_rank * RADIUS
        x = rng.uniform(0.0, self.cross, n).astype(np.float32)
        y = rng.uniform(0.0, self.cross, n).astype(np.float32)
        z = rng.uniform(z_lo, np.nextafter(np.float32(z_hi), np.float32(0)), n).astype(np.float32)
        ids = np.arange(n, dtype=np.uint32) + np.uint32(rank * n + 1)
        return {"x": x, "y": y, "z": z, "_id": ids}

    def model_params(self):
        p = dict(env_max=self.cross, radius=RADIUS, repulse=REPULSE)
        if self.depth != self.cross:
            p["env_max_z"] = self.depth
        return p


def make_sim(box, rank, world, local, **cfg):
    """(simulation, slab driver or None): one GPU runs the plain simulation, N > 1 the slab-decomposed one."""
    from flamegpu2_b200 import sim as fsim

    if world == 1:
        return fsim.Simulation("circles", device=local, **box.model_params(), **cfg), None
    from flamegpu2_b200 import slab

    per_plane = box.n_per_rank // box.planes_per_rank
    halo_cap = int(3 * per_plane + 8192)   # a boundary plane holds ~n/planes messages (3x: clustering)
    mig_cap = int(min(per_plane // 2 + 4096, 131072))   # a few percent of a plane changes slab per step
    sl = slab.SlabSimulation("circles", "Circle", "location", rank, world, local, box.planes, halo_capacity=halo_cap,
                             migrate_capacity=mig_cap, **box.model_params(), **cfg)
    return sl.sim, sl


def timed_run(box, rank, world, local, steps, warmup, flush, dist, **cfg):
    """Runs warmup + steps steps of `box`; returns (seconds of the timed steps: max over ranks, wall, sim facts)."""
    import torch

    cfg_py_loop = PY_LOOP

    s, sl = make_sim(box, rank, world, local, timing=1, **cfg)
    s.set_population("Circle", box.population(rank))
    stream = torch.cuda.ExternalStream(s.stream, device=f"cuda:{local}")

    def one_step():
        if flush is not None:
            with torch.cuda.stream(stream):
                flush.add_(1)  # evicts the step's working set from the 126 MB L2; outside the timed events
        s.step(1)

    for _ in range(warmup):
        one_step()
    s.sync()
    s.step_times()  # drop warm-up timings
    launches0 = s.launches
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:  # NVML queries take driver locks: one sampling rank is enough, the others would only add jitter to a coupled step
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(stream)
    if flush is None and not cfg_py_loop:
        # one call: the K-step loop runs in C++ (CUDASimulation::step() K times).  The ranks of a slab run are coupled every
        # step, so the slowest host of every step gates all of them; a Python loop adds its jitter (GIL hand-overs with the
        # clock sampler thread) to every step of every rank
        s.step(steps)
    else:
        for _ in range(steps):
            one_step()
    ev1.record(stream)
    s.sync()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    if dist is not None:
        dist.barrier()
    per_step = s.step_times()
    # The same clock at every N: the sum of this rank's per-step CUDA-event times (CUDASimulation::step() records an event
    # pair around every step on the simulation stream, like the reference's getElapsedTimeSteps()), max over ranks.  Waits for
    # neighbours (halo, migration, all-reduce) happen inside a step and are included; what lies BETWEEN two steps -- at N=1 the
    # L2 flush, at every N the host's turn-around after a step function made it wait for the step -- is not.  The bracketed
    # K-step region (everything between the first and the last event) is reported next to it as `region`.
    dev_total = float(per_step.sum())
    region = ev0.elapsed_time(ev1) * 1e-3
    if world > 1:
        sl.check_overflow()
    facts = {"launches": s.launches - launches0, "graphs": s.graphs, "clocks": clk, "per_step": per_step,
             "agents_end": s.count("Circle")}
    if dist is not None:
        t = torch.tensor([dev_total, wall, region], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_total, wall, region = float(t[0]), float(t[1]), float(t[2])
        dist.barrier()
    facts["region"] = region
    s.close()
    return dev_total, wall, facts


def profile_phases(box, local, flush, warmup, steps=30, **cfg):
    """{phase: microseconds per call} from an eager pass with CUDA events around every phase (one GPU)."""
    import torch
    from flamegpu2_b200 import sim as fsim

    p = fsim.Simulation("circles", device=local, profile=1, **box.model_params(), **cfg)
    p.set_population("Circle", box.population(0))
    pstream = torch.cuda.ExternalStream(p.stream, device=f"cuda:{local}")
    for i in range(warmup + steps):
        if flush is not None:
            with torch.cuda.stream(pstream):
                flush.add_(1)
        p.step(1)
        if i == warmup - 1:
            p.profile()
    prof = p.profile()
    p.close()
    return {k: v[0] / v[1] * 1e3 for k, v in prof.items() if v[1]}


def roofline_entry(name, alg_bytes, us, peak, peak_src, **extra):
    e = {"bound": "hbm", "kernel": name, "achieved": alg_bytes / us / 1e3, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
         "frac": alg_bytes / us / 1e3 / peak, "algorithmic_bytes": alg_bytes, "us_per_launch": us}
    e.update(extra)
    return e


def committed_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return {}


# ---- the reference arm ------------------------------------------------------------------------------------------
def cpu_circles(n, L, steps, seed=0):
    """The CPU rendition (oracle port, OpenMP over all host threads) of the same step."""
    import oracle_py as orc

    g = orc.Grid(3, (0, 0, 0), (L, L, L), RADIUS)
    rng = np.random.default_rng(seed)
    x, y, z = [rng.uniform(0.0, L, n).astype(np.float32) for _ in range(3)]
    ids = np.arange(1, n + 1, dtype=np.uint32)
    d = np.zeros(n, np.float32)
    ids, x, y, z, d, _ = g.circles_step(ids, x, y, z, d, REPULSE)  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    for _ in range(steps):
        ids, x, y, z, d, _ = g.circles_step(ids, x, y, z, d, REPULSE)
    dt = time.perf_counter() - t0
    return n * steps / dt, dt / steps, orc.num_threads()


def reference_cuda(box, steps=30, warmup=5):
    """The reference's own CUDA build (oracle/_ref/ref_sim: FLAME GPU 2 compiled unmodified for sm_100a) on ONE GPU, whole
    `box` (the reference cannot span GPUs), same seeded population, same model source incl. the Validation step function."""
    import tempfile

    import fgbs

    if not fgbs.have_ref():
        return None
    pops = [box.population(r) for r in range(box.world)]
    cols = {k: np.concatenate([p[k] for p in pops]) for k in ("x", "y", "z")}
    n = len(cols["x"])
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "in.bin")
        fgbs.write_state(inp, cols)
        params = {"env_max": box.cross, "radius": RADIUS, "repulse": REPULSE}
        if box.depth != box.cross:
            params["env_max_z"] = box.depth
        js = fgbs.run_ref("circles", params, inp, os.path.join(td, "ref"), steps=steps, warmup=warmup, timeout=900)
    t = np.array(js["step_seconds"])
    return {"value": float(n * len(t) / t.sum()), "unit": "agent-steps/s", "ms_per_step": float(t.mean() * 1e3), "agents": n, "steps": len(t),
            "what": "FLAME GPU 2 v2.0.0-rc.5 built unmodified for sm_100a (seatbelts off) by oracle/ref_build/build_ref.sh, one GPU, "
                    "per-step times from getElapsedTimeSteps()"}


def run_reference(args):
    """--impl reference.  FLAME GPU 2 has no CPU implementation of this path: the reference arm is the reference's own CUDA
    build (oracle/_ref/ref_sim) on ONE GPU of this box, on the whole N-GPU workload (it cannot span GPUs).  When the build
    is absent the OpenMP C restatement of the step (oracle/, kind "port") runs instead on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)  # torchrun pins it to 1
    box = Box.cube(args.agents_per_gpu, args.gpus)
    n = box.n_per_rank * box.world
    steps = max(1, min(args.steps, 60))
    ref = None
    try:
        ref = reference_cuda(box, steps=steps, warmup=max(args.warmup, 3))
    except Exception as e:  # reported below
        ref = {"error": str(e)[:300]}
    line = {"impl": "reference", "metric": "agent-steps/sec", "unit": "agent-steps/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": max(args.warmup, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "gpu_launches": 0,
            "config": {"workload": f"Circles-3D, {args.agents_per_gpu} agents per GPU, {box.describe()}, whole CUDASimulation::step() (output_message, "
                                   "auto agent sort, PBM buildIndex, move, the example's Validation step function)",
                       "agents": n, "agents_per_gpu": args.agents_per_gpu}}
    if ref and "value" in ref:
        line.update({"value": ref["value"], "ms_per_step": ref["ms_per_step"]})
        line["cpu_baseline"] = {"value": ref["value"], "unit": "agent-steps/s", "cores": 0, "kind": "ref_cuda",
                                "sample": f"{ref['steps']} steps of the full {n}-agent workload on ONE B200", "what": ref["what"]}
        line["config"]["note"] = ("FLAME GPU 2 has no CPU path: this arm is the reference's own CUDA implementation, compiled here, on one GPU "
                                  "(a reference simulation cannot span GPUs; at N > 1 it runs the whole N-GPU workload on one)")
        line["gpu_launches"] = None
    else:
        L = env_extent(n)
        k = max(1, min(steps, 10))
        value, sec, cores = cpu_circles(n, L, k)
        line.update({"value": value, "ms_per_step": sec * 1e3, "steps": k})
        line["cpu_baseline"] = {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                                "sample": f"{k} steps of the full {n}-agent workload", "ref_cuda_error": (ref or {}).get("error")}
        line["config"]["note"] = "oracle/_ref/ref_sim is not built: OpenMP C restatement of the step (oracle/) on the host cores"
    line["e2e"] = {"value": line["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(line), flush=True)


def kernel_rooflines(local, peak, peak_src, n=1 << 24):
    """Isolated-kernel rooflines through the C ABI at 16.8 M items (L2 flushed between repetitions): death compaction."""
    import torch
    from flamegpu2_b200 import host

    dev = f"cuda:{local}"
    ctx = host.Context(local)
    flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
    vars_in = [torch.rand(n, device=dev) for _ in range(4)] + [torch.arange(n, dtype=torch.int32, device=dev) for _ in range(2)]
    vars_out = [torch.empty_like(a) for a in vars_in]
    flags = (torch.rand(n, device=dev) >= 0.1).to(torch.int32)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    ctx.reserve(n, 0)
    ts = []
    for i in range(13):
        flush.add_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.compact(flags, vars_in, vars_out, n, d_out_count=cnt)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    keep = int(cnt.item())
    us = float(np.median(ts))
    ctx.close()
    return {"compact_death_16m": roofline_entry("fgb_compact (k_compact), 16.8 M agents x 24 B, 10 % deaths, isolated", n * 4 + n * S_AGENT + keep * S_AGENT,
                                                us, peak, peak_src)}


def e2e_host_buffers(box, local, k_e2e, cfg):
    """agent-steps/s through the reference-facing API with HOST buffers: every step uploads the population from pinned host
    memory, runs CUDASimulation::step() and downloads the resulting population."""
    import torch
    from flamegpu2_b200 import sim as fsim

    n = box.n_per_rank
    pop = box.population(0)
    s = fsim.Simulation("circles", device=local, **box.model_params(), **cfg)
    s.set_population("Circle", {k: pop[k] for k in ("x", "y", "z")})
    hx, hy, hz, hd = (torch.from_numpy(a.copy()).pin_memory().numpy() for a in (pop["x"], pop["y"], pop["z"], np.zeros(n, np.float32)))
    outs = [{k: torch.empty(n, dtype=torch.float32).pin_memory().numpy() for k in ("x", "y", "z", "drift")} for _ in range(2)]
    for o in outs:
        o["id"] = torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    for i in range(3):
        s.circles_step_host(hx, hy, hz, hd, 1, outs[i & 1])
    t0 = time.perf_counter()
    for i in range(k_e2e):
        out = outs[i & 1]
        s.circles_step_host(hx, hy, hz, hd, 1, out)
        hx, hy, hz, hd = out["x"], out["y"], out["z"], out["drift"]  # the next step starts from this step's result
    e2e_t = time.perf_counter() - t0
    s.close()
    return {"value": n * k_e2e / e2e_t, "unit": "agent-steps/s", "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 20 * n, "steps": k_e2e,
            "ms_per_step": e2e_t / k_e2e * 1e3,
            "call": "fgbm_circles_step_host: setPopulationDataSoA + CUDASimulation::step() + getPopulationDataSoA, wall clock"}


# ---- this repo ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--no-extras", action="store_true", help="skip every leg beyond the headline line (and the N>1 parity check)")
    ap.add_argument("--no-north-star", action="store_true", help="skip the 16.8 M-agents-per-GPU and 128 M strong-scaling legs")
    ap.add_argument("--stable", type=int, default=0)
    ap.add_argument("--true3d-sort", type=int, default=0)
    ap.add_argument("--overlap", type=int, default=1, help="1: build the PBM on a second stream while the agents are sorted")
    ap.add_argument("--block", type=int, default=128, help="threads per block of the agent function kernels")
    ap.add_argument("--tile-order", type=int, default=1, help="1: tile-local execution order after the auto sort")
    ap.add_argument("--iter-mode", type=int, default=-1,
                    help="-1 per function as the model declares (Circles `move`: radius-filtered), 0 reference visit order everywhere, 1 radius-filtered everywhere")
    ap.add_argument("--bin-order", type=int, default=1, help="run message-reading functions in bin order (b200 extension)")
    ap.add_argument("--fused-index", type=int, default=1)
    ap.add_argument("--ordered-output", type=int, default=1)
    ap.add_argument("--py-loop", type=int, default=0, help="N > 1: drive the timed steps from a Python loop instead of one C++ call (A/B)")
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

    cfg = dict(stable=args.stable, true3d_sort=args.true3d_sort, bin_order=args.bin_order, iter_mode=args.iter_mode, overlap=args.overlap,
               tile_order=args.tile_order, block=args.block, fused_index=args.fused_index, ordered_output=args.ordered_output)
    global PY_LOOP
    PY_LOOP = bool(args.py_loop)
    n = args.agents_per_gpu
    box = Box.cube(n, world)
    flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.int32, device=f"cuda:{local}") if world == 1 else None

    # -- N > 1: parity self-check before anything is timed (one teacher-forced step, slabs vs one GPU)
    parity = None
    if world > 1 and not args.no_extras:
        import slab_parity_check

        parity = slab_parity_check.run_check(rank, world, local)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": "agent-steps/sec", "n_gpus": world, "parity_check": parity, "error": "slab parity check failed"}), flush=True)
            dist.destroy_process_group()
            sys.exit(3)

    dev_total, wall, facts = timed_run(box, rank, world, local, args.steps, args.warmup, flush, dist, **cfg)
    value = box.n_per_rank * world * args.steps / dev_total
    ps = facts["per_step"]
    line = {
        "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"Circles-3D, {n} agents per GPU, {box.describe()}, whole CUDASimulation::step() (output_message, auto agent "
                        "sort, PBM buildIndex, move, the example's Validation step function)",
            "agents_per_gpu": n, "bins": box.bins_per_rank * world, "graphs": facts["graphs"], "bin_order_execution": bool(args.bin_order),
            "iterator_mode": args.iter_mode,
            "iterator": ("per function as the model declares: Circles `move` runs the radius-filtered walk" if args.iter_mode < 0 else
                         ("reference order" if args.iter_mode == 0 else "radius-filtered")),
            "overlap_index_build": bool(args.overlap), "tile_local_exec_order": bool(args.tile_order),
            "fused_index_build": bool(args.fused_index), "bin_ordered_output": bool(args.ordered_output),
            "l2": "flushed between steps (256 MiB write outside the timed events)" if world == 1 else
                  "not flushed (the steps of neighbouring ranks are coupled by the exchange; per-GPU working set ~80 MB)",
            "timing": "sum of per-step CUDA-event times on the simulation stream (waits for neighbours happen inside a step and are included), "
                      "max over ranks; the same clock at every N",
            "region_ms_per_step": (facts["region"] / args.steps * 1e3) if world > 1 else None,
            "region": "CUDA events around the whole K-step loop incl. the host's turn-around between steps, max over ranks (N=1: the loop also "
                      "contains the L2 flushes, see wall_ms_per_step)",
            "wall_ms_per_step": wall / args.steps * 1e3,
            "ms_first_30_steps": float(ps[:30].mean() * 1e3) if len(ps) else None,
            "ms_last_30_steps": float(ps[-30:].mean() * 1e3) if len(ps) else None,
            "multi_gpu": (f"z-slab decomposition over {world} GPUs behind CUDASimulation::step(): halo messages + migrating agents packed "
                          "straight into the neighbour's memory over NVLink (CUDA IPC peer mappings), captured in the step graphs")
                         if world > 1 else "single GPU",
        },
        "gpu_launches": int(facts["launches"]),
        "clocks": facts["clocks"],
    }
    if parity is not None:
        line["parity_check"] = parity
    extras = not args.no_extras
    peak, peak_src = measured_peak()

    # -- the north-star share: 16.8 M agents per GPU (128 M on 8 GPUs), and the fixed 128 M box (strong scaling)
    if extras and not args.no_north_star:
        ns_box = Box(NORTH_STAR_CROSS, NORTH_STAR_PER_GPU / NORTH_STAR_CROSS ** 2, world)
        k_ns = 20
        t_ns, _, f_ns = timed_run(ns_box, rank, world, local, k_ns, 5, None, dist, **cfg)
        line["north_star"] = {"value": ns_box.n_per_rank * world * k_ns / t_ns, "unit": "agent-steps/s", "ms_per_step": t_ns / k_ns * 1e3,
                              "steps": k_ns, "warmup": 5, "agents_per_gpu": ns_box.n_per_rank, "agents": ns_box.n_per_rank * world,
                              "scaling": "weak", "graphs": f_ns["graphs"], "workload": "Circles-3D, " + ns_box.describe()}
        if world < 8:
            total_depth = NORTH_STAR_PER_GPU * 8 / NORTH_STAR_CROSS ** 2  # the 512^3 box of BASELINE configs[4]
            st_box = Box(NORTH_STAR_CROSS, total_depth / world, world)
            k_st = 10 if world == 1 else 20
            t_st, _, f_st = timed_run(st_box, rank, world, local, k_st, 3, None, dist, **cfg)
            line["strong_scaling"] = {"value": st_box.n_per_rank * world * k_st / t_st, "unit": "agent-steps/s", "ms_per_step": t_st / k_st * 1e3,
                                      "steps": k_st, "agents": st_box.n_per_rank * world, "scaling": "strong",
                                      "workload": "Circles-3D, " + st_box.describe()}
        else:
            line["strong_scaling"] = dict(line["north_star"], scaling="strong", note="at 8 GPUs the 128 M box IS the north-star run")

    if rank == 0 and world == 1 and extras:
        # -- per-phase device times from a profiled (eager, event-bracketed) pass of the same workload
        phases = profile_phases(box, local, flush, args.warmup, **cfg)
        bins = box.bins_per_rank
        alg = n * 2 * S_MSG + 4 * (bins + 1)
        traffic = committed_json("ncu_traffic.json")
        bi_us = phases.get("build_index")
        if bi_us:
            line["roofline"] = roofline_entry("buildIndex in-step: k_scan_scatter (scan + scatter of every tile in ONE launch); bin keys and histogram are "
                                              "published by the output function", alg, bi_us, peak, peak_src,
                                              traffic=traffic.get("build_index", {}).get(str(n)))
        line["phases_us"] = phases
        rl = {}
        if phases.get("agent_sort"):
            rl["agent_sort_1m"] = roofline_entry("automatic agent sort in-step (keys + histogram, 2 onesweep passes, gather of every variable)",
                                                 n * (2 * S_AGENT + 12), phases["agent_sort"], peak, peak_src)
        if not args.no_north_star:
            ns1 = Box(NORTH_STAR_CROSS, NORTH_STAR_PER_GPU / NORTH_STAR_CROSS ** 2, 1)
            ph16 = profile_phases(ns1, local, None, 5, steps=10, **cfg)
            n16, bins16 = ns1.n_per_rank, ns1.bins_per_rank
            if ph16.get("build_index"):
                rl["build_index_16m"] = roofline_entry("buildIndex in-step, 16.8 M messages", n16 * 2 * S_MSG + 4 * (bins16 + 1), ph16["build_index"],
                                                       peak, peak_src, traffic=traffic.get("build_index", {}).get(str(n16)))
            if ph16.get("agent_sort"):
                rl["agent_sort_16m"] = roofline_entry("automatic agent sort in-step, 16.8 M agents", n16 * (2 * S_AGENT + 12), ph16["agent_sort"], peak, peak_src)
            line["phases_us_16m"] = ph16
        try:
            rl.update(kernel_rooflines(local, peak, peak_src))
        except Exception as e:  # reported, never fatal
            rl["error"] = str(e)[:200]
        line["rooflines"] = rl
        mv = committed_json("move_bound.json")
        if mv:
            line["move"] = mv

        # -- end to end through the C ABI with HOST buffers (copies inside the timed region)
        line["e2e"] = e2e_host_buffers(box, local, min(args.steps, 60), cfg)
        # -- CPU baseline: the OpenMP port on this box's host cores, bounded sample
        L = box.cross
        _, sec1, cores = cpu_circles(n, L, 1)
        ksteps = int(max(1, min(30, 10.0 / max(sec1, 1e-3))))
        v, sec, cores = cpu_circles(n, L, ksteps)
        line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                                "sample": f"{ksteps} steps of the same {n}-agent Circles workload"}
        try:
            line["reference_cuda"] = reference_cuda(box)
        except Exception as e:
            line["reference_cuda"] = {"error": str(e)[:200]}
        # -- the same workload with the strict reference iteration order for every function (iter_mode 0)
        if args.iter_mode != 0:
            c0 = dict(cfg, iter_mode=0)
            t0_, _, f0 = timed_run(box, rank, 1, local, args.steps, args.warmup, flush, None, **c0)
            line["modes"] = {"reference_iteration_order": {"value": n * args.steps / t0_, "unit": "agent-steps/s", "ms_per_step": t0_ / args.steps * 1e3,
                                                           "ms_first_30_steps": float(f0["per_step"][:30].mean() * 1e3),
                                                           "ms_last_30_steps": float(f0["per_step"][-30:].mean() * 1e3)}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
