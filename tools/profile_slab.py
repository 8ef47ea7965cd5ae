"""Per-phase device times of the slab-decomposed Circles step (eager profiled pass on every rank):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/profile_slab.py --cross 512 --depth 64"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--cross", type=float, default=512.0)
ap.add_argument("--depth", type=float, default=64.0, help="per rank")
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--profile", type=int, default=1)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
box = bench.Box(a.cross, a.depth, world)
s, sl = bench.make_sim(box, rank, world, local, profile=a.profile, timing=1)
s.set_population("Circle", box.population(rank))
for i in range(4 + a.steps):
    s.step(1)
    if i == 3:
        s.sync()
        dist.barrier()
        if a.profile:
            s.profile()
        s.step_times()
s.sync()
t = s.step_times() * 1e3
out = {"rank": rank, "agents": box.n_per_rank, "ms_per_step": float(t.mean()), "series_us": [int(round(v * 1e3)) for v in t[:48]]}
if a.profile:
    prof = s.profile()
    out["phases_us"] = {k: round(v[0] / v[1] * 1e3, 1) for k, v in prof.items() if v[1]}
sl.check_overflow()
print(json.dumps(out), flush=True)
dist.barrier()
s.close()
dist.destroy_process_group()
