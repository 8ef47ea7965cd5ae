"""Runs K eager (graphs=0) steps of the Circles model; used under ncu to capture `move` at a chosen step.
  python tools/run_circles.py --n 1000000 --steps 301 --iter-mode 1 [--times]"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flamegpu2_b200 import sim as fsim

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=301)
ap.add_argument("--iter-mode", type=int, default=0)
ap.add_argument("--graphs", type=int, default=0)
ap.add_argument("--times", action="store_true")
ap.add_argument("--extra", default="")
ap.add_argument("--cross", type=float, default=0.0, help="> 0: box [0,cross)^2 x [0,depth) at density 1 instead of the cube of --n agents")
ap.add_argument("--depth", type=float, default=64.0)
a = ap.parse_args()
rng = np.random.default_rng(0)
kw = dict(kv.split("=") for kv in a.extra.split(",") if kv)
if a.cross > 0:
    L = a.cross
    a.n = int(round(a.cross * a.cross * a.depth))
    x, y = [rng.uniform(0.0, L, a.n).astype(np.float32) for _ in range(2)]
    z = rng.uniform(0.0, a.depth, a.n).astype(np.float32)
    kw["env_max_z"] = a.depth
else:
    L = float(np.floor(np.cbrt(float(a.n)) + 1e-6))
    x, y, z = [rng.uniform(0.0, L, a.n).astype(np.float32) for _ in range(3)]
s = fsim.Simulation("circles", env_max=L, radius=2.0, repulse=0.05, graphs=a.graphs, iter_mode=a.iter_mode, timing=1 if a.times else 0, **kw)
s.set_population("Circle", {"x": x, "y": y, "z": z})
s.step(a.steps)
s.sync()
if a.times:
    t = s.step_times() * 1e3
    print("ms/step by 30:", [round(float(t[i:i + 30].mean()), 3) for i in range(0, len(t), 30)])
s.close()
