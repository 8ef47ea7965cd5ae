"""The other BASELINE.json configurations on ONE GPU (bench.py measures configs[1]):

  circles3d_16m  Circles-3D, 16.8 M agents (the per-GPU share of the 128 M / 8 GPU configuration, configs[4])
  boids2d_16m    Boids spatial2D, 16 M agents (configs[2]: 2D PBM + strip iterator)
  stress_32m     birth/death stress, 32 M agents, 10 % death and 5 % birth per step (configs[3])

For each: whole CUDASimulation::step() device time of this repo (per-step CUDA events, working set >> L2),
a per-phase breakdown from a profiled pass, and the reference's own CUDA build (oracle/_ref/ref_sim) on the same
input when it is present (which is why this script lives under tests/: it runs oracle/_ref).  One JSON line per
configuration on stdout.

  python tests/bench_configs.py [--only NAME] [--steps K] [--no-ref]
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def circles_pop(n, L, seed=0):
    rng = np.random.default_rng(seed)
    return {k: rng.uniform(0.0, L, n).astype(np.float32) for k in ("x", "y", "z")}


def boids2d_pop(n, seed=0):
    rng = np.random.default_rng(seed)
    pop = {k: rng.uniform(-0.5, 0.5, n).astype(np.float32) for k in ("x", "y")}
    v = rng.uniform(-1, 1, (2, n)).astype(np.float32)
    v = v / np.maximum(np.linalg.norm(v, axis=0), 1e-6) * rng.uniform(0.1, 1.0, n).astype(np.float32)
    pop["fx"], pop["fy"] = v[0].astype(np.float32), v[1].astype(np.float32)
    return pop


CONFIGS = {
    "circles3d_16m": dict(model="circles", agent="Circle", n=16_777_216,
                          params=lambda n: {"env_max": 256.0, "radius": 2.0, "repulse": 0.05},
                          pop=lambda n: circles_pop(n, 256.0),
                          what="Circles-3D, [0,256)^3, radius 2 (2,097,152 bins, ~8 agents/bin)"),
    "boids2d_16m": dict(model="boids2d", agent="Boid", n=16_000_000,
                        params=lambda n: {"interaction_radius": 0.0007, "separation_radius": 0.00014},
                        pop=lambda n: boids2d_pop(n),
                        what="Boids spatial2D, [-0.5,0.5]^2, interaction radius 0.0007 (1429^2 bins, ~7.8 boids/bin)"),
    "stress_32m": dict(model="stress", agent="Circle", n=32_000_000,
                       params=lambda n: {"env_max": 318.0, "radius": 2.0, "death_mod": 10, "birth_mod": 20},
                       pop=lambda n: circles_pop(n, 318.0),
                       what="birth/death stress, [0,318)^3, radius 2: neighbour count + 10 % death + 5 % birth (agent_out) per step"),
    # small variant for profiling runs (not a BASELINE configuration)
    "stress_4m": dict(model="stress", agent="Circle", n=4_000_000,
                      params=lambda n: {"env_max": 159.0, "radius": 2.0, "death_mod": 10, "birth_mod": 20},
                      pop=lambda n: circles_pop(n, 159.0),
                      what="birth/death stress, [0,159)^3, radius 2 (profiling size)"),
}


def run_ours(cfg, steps, warmup, iter_mode=0):
    from flamegpu2_b200 import sim as fsim

    n = cfg["n"]
    pop = cfg["pop"](n)
    out = {}
    s = fsim.Simulation(cfg["model"], device=0, timing=1, iter_mode=iter_mode, **cfg["params"](n))
    s.set_population(cfg["agent"], pop)
    s.step(warmup)
    s.sync()
    s.step_times()  # drop the warm-up steps (the first one runs on the unsorted initial population)
    n0 = s.count(cfg["agent"])
    s.step(steps)
    s.sync()
    t = s.step_times()
    n1 = s.count(cfg["agent"])
    s.close()
    agents = 0.5 * (n0 + n1)  # population drifts in the stress model
    out["ms_per_step"] = float(t.mean() * 1e3)
    out["value"] = float(agents / t.mean())
    out["agents_start_end"] = [n0, n1]
    p = fsim.Simulation(cfg["model"], device=0, profile=1, iter_mode=iter_mode, **cfg["params"](n))
    p.set_population(cfg["agent"], pop)
    p.step(warmup)
    p.profile()
    p.step(min(steps, 5))
    prof = p.profile()
    p.close()
    out["phases_us"] = {k: v[0] / v[1] * 1e3 for k, v in prof.items() if v[1]}
    return out, pop


def run_reference(cfg, pop, steps, warmup, agents):
    import fgbs

    if not fgbs.have_ref():
        return None
    try:
        with tempfile.TemporaryDirectory() as td:
            inp = os.path.join(td, "in.bin")
            fgbs.write_state(inp, pop)
            js = fgbs.run_ref(cfg["model"], cfg["params"](cfg["n"]), inp, os.path.join(td, "ref"), steps=steps, warmup=warmup,
                              timeout=900)
        t = np.array(js["step_seconds"])
        return {"ms_per_step": float(t.mean() * 1e3), "value": float(agents / t.mean()),
                "what": "FLAME GPU 2 v2.0.0-rc.5 built unmodified for sm_100a, getElapsedTimeSteps()"}
    except Exception as e:  # reported, never fatal
        return {"error": str(e)[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--iter-mode", type=int, default=-1, help="-1 per function as the model declares (default), 0 reference order everywhere, 1 radius-filtered everywhere")
    args = ap.parse_args()
    for name, cfg in CONFIGS.items():
        if (args.only and name != args.only) or (not args.only and name == "stress_4m"):
            continue
        ours, pop = run_ours(cfg, args.steps, args.warmup, args.iter_mode)
        line = {"config": name, "iterator_mode": args.iter_mode, "workload": cfg["what"], "agents": cfg["n"], "steps": args.steps, "warmup": args.warmup,
                "unit": "agent-steps/s", **ours}
        if not args.no_ref:
            # same population trajectory (the models are deterministic), so the same mean agent count
            line["reference_cuda"] = run_reference(cfg, pop, args.steps, args.warmup, 0.5 * sum(ours["agents_start_end"]))
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
