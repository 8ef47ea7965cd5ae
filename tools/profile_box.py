"""Per-phase device times of a Circles box on one GPU (eager profiled pass):  python tools/profile_box.py --cross 512 --depth 512"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--cross", type=float, default=512.0)
ap.add_argument("--depth", type=float, default=64.0)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--iter-mode", type=int, default=-1)
ap.add_argument("--extra", default="", help="k=v,k=v further simulation parameters")
a = ap.parse_args()
box = bench.Box(a.cross, a.depth, 1)
kw = dict(kv.split("=") for kv in a.extra.split(",") if kv)
ph = bench.profile_phases(box, 0, None, 3, steps=a.steps, iter_mode=a.iter_mode, **kw)
print(json.dumps({"agents": box.n_per_rank, "bins": box.bins_per_rank, "phases_us": ph, "total_ms": sum(ph.values()) / 1e3}))
