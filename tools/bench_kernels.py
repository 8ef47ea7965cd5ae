"""Kernel-level microbenchmarks through the C ABI (CUDA-event timing, L2 flushed between reps).
Writes one JSON object per line to stdout.  Used for profiles/, not for BENCH json."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flamegpu2_b200 import host  # noqa: E402

DEV = "cuda:0"


def timed(fn, reps=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.add_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def circles_positions(n, L, sorted_like, seed=0):
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    pos = torch.rand((3, n), generator=g, device=DEV) * L
    if sorted_like:
        # steady-state message order: agents are bin-sorted every step, then move a little
        cell = torch.floor(pos / 2.0).to(torch.int64)
        gd = int(np.ceil(L / 2.0))
        key = (cell[2] * gd + cell[1]) * gd + cell[0]
        order = torch.argsort(key, stable=True)
        pos = pos[:, order].contiguous()
        pos = (pos + (torch.rand((3, n), generator=g, device=DEV) - 0.5) * 0.1).clamp_(0, L * 0.999999)
    return pos.contiguous()


def boids2d_case(ctx, n, reps, flush, peak, jitter_bins):
    """2D list of the Boids-2D shape: bin-sorted a step ago, every point then moved by up to jitter_bins bins."""
    r = 0.0007
    sp = host.Spatial(ctx, 2, (-0.5, -0.5), (0.5, 0.5), r)
    sp.reserve(n)
    g = torch.Generator(device=DEV)
    g.manual_seed(5)
    pos = torch.rand((2, n), generator=g, device=DEV) - 0.5
    gd = sp.grid_dim[0]
    cell = torch.floor((pos + 0.5) / r).to(torch.int64).clamp_(0, gd - 1)
    order = torch.argsort(cell[1] * gd + cell[0], stable=True)
    pos = pos[:, order].contiguous()
    pos = (pos + (torch.rand((2, n), generator=g, device=DEV) - 0.5) * 2 * jitter_bins * r).clamp_(-0.5, 0.4999999).contiguous()
    ids = torch.arange(n, dtype=torch.int32, device=DEV)
    fx = torch.rand(n, generator=g, device=DEV)
    ins = [fx, fx.clone(), ids, pos[0].contiguous(), pos[1].contiguous()]
    outs = [torch.empty_like(a) for a in ins]
    alg = n * 2 * 20 + 4 * (sp.bin_count + 1)
    med, best = timed(lambda: sp.build_index(ins[3], ins[4], None, ins, outs, n), reps=reps, flush=flush)
    print(json.dumps({"op": "build_index_2d", "n": n, "bins": sp.bin_count, "jitter_bins": jitter_bins, "us_median": med,
                      "us_best": best, "alg_bytes": alg, "GBps": alg / med / 1e3, "frac_of_measured_peak": alg / med / 1e3 / peak}),
          flush=True)
    perm = torch.empty(n, dtype=torch.int32, device=DEV)
    for tl in (False, True):
        med, best = timed(lambda: sp.bin_permutation(ins[3], ins[4], None, perm, n, tile_local=tl), reps=reps, flush=flush)
        print(json.dumps({"op": "bin_permutation_2d", "n": n, "tile_local": tl, "jitter_bins": jitter_bins, "us_median": med,
                          "us_best": best}), flush=True)
    sp.close()


def bucket_case(ctx, n, buckets, reps, flush, peak, order):
    """MessageBucket build: n messages {_key, id, payload}, keys uniform over `buckets` buckets"""
    g = torch.Generator(device=DEV)
    g.manual_seed(9)
    keys = torch.randint(0, buckets, (n,), generator=g, device=DEV, dtype=torch.int32)
    if order == "grouped":  # messages written by agents that are themselves ordered by key, slightly perturbed
        keys = torch.sort(keys).values
        keys = (keys + torch.randint(-1, 2, (n,), generator=g, device=DEV, dtype=torch.int32)).clamp_(0, buckets - 1)
    b = host.Bucket(ctx, 0, buckets - 1)
    ins = [keys, torch.arange(n, dtype=torch.int32, device=DEV), torch.rand(n, generator=g, device=DEV)]
    outs = [torch.empty_like(a) for a in ins]
    alg = n * 2 * 12 + 4 * (buckets + 1)
    med, best = timed(lambda: b.build_index(ins[0], ins, outs, n), reps=reps, flush=flush)
    print(json.dumps({"op": "bucket_build_index", "n": n, "buckets": buckets, "key_order": order, "us_median": med, "us_best": best,
                      "alg_bytes": alg, "GBps": alg / med / 1e3, "frac_of_measured_peak": alg / med / 1e3 / peak}), flush=True)
    b.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1000000,16777216")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--boids2d", action="store_true", help="only the 2D (Boids) build / permutation cases")
    ap.add_argument("--bucket", action="store_true", help="only the MessageBucket build cases")
    ap.add_argument("--jitter", type=float, default=None, help="--boids2d: displacement in bins since the list was sorted")
    args = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ctx = host.Context(0)
    flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.int32, device=DEV)  # 256 MB > 126 MB L2
    if args.bucket:
        for n, buckets in ((1_000_000, 125_000), (16_777_216, 2_097_152)):
            for order in ("grouped", "random"):
                bucket_case(ctx, n, buckets, args.reps, flush, peak, order)
        return
    if args.boids2d:
        for jit in ([args.jitter] if args.jitter is not None else [0.05, 0.7]):
            boids2d_case(ctx, 16_000_000, args.reps, flush, peak, jit)
        return
    for n in [int(s) for s in args.sizes.split(",")]:
        L = float(round((n) ** (1.0 / 3.0)))
        sp = host.Spatial(ctx, 3, (0, 0, 0), (L, L, L), 2.0)
        sp.reserve(n)
        for sorted_like in (True, False):
            pos = circles_positions(n, L, sorted_like)
            ids = torch.arange(n, dtype=torch.int32, device=DEV)
            ins = [ids, pos[0].contiguous(), pos[1].contiguous(), pos[2].contiguous()]
            outs = [torch.empty_like(a) for a in ins]
            alg = n * 2 * 16 + 4 * (sp.bin_count + 1)
            for stable, expect in ((False, False), (False, True), (True, False)):
                med, best = timed(lambda: sp.build_index(ins[1], ins[2], ins[3], ins, outs, n, stable=stable, expect_grouped=expect),
                                  reps=args.reps, flush=flush)
                print(json.dumps({"op": "build_index", "n": n, "bins": sp.bin_count, "stable": stable, "expect_grouped": expect,
                                  "input": "bin-sorted+jitter" if sorted_like else "random", "us_median": med,
                                  "us_best": best, "alg_bytes": alg, "GBps": alg / med / 1e3,
                                  "frac_of_measured_peak": alg / med / 1e3 / peak}), flush=True)
        # death compaction: 24 B/agent, 10% die
        vars_in = [torch.rand(n, device=DEV) for _ in range(4)] + [torch.arange(n, dtype=torch.int32, device=DEV)] * 2
        vars_out = [torch.empty_like(a) for a in vars_in]
        flags = (torch.rand(n, device=DEV) >= 0.1).to(torch.int32)
        cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
        ctx.reserve(n, 0)
        med, best = timed(lambda: ctx.compact(flags, vars_in, vars_out, n, d_out_count=cnt), reps=args.reps, flush=flush)
        keep = int(cnt.item())
        alg = n * 4 + n * 24 + keep * 24
        print(json.dumps({"op": "compact_death", "n": n, "kept": keep, "us_median": med, "us_best": best,
                          "alg_bytes": alg, "GBps": alg / med / 1e3, "frac_of_measured_peak": alg / med / 1e3 / peak}),
              flush=True)
        # agent sort (key + stable sort + gather of 6 x 4-byte variables)
        gd = int(np.ceil(L / 2.0))
        mb = int(np.floor(np.log2(gd ** 3))) + 1
        pos = circles_positions(n, L, True, seed=3)
        keys = torch.empty(n, dtype=torch.int32, device=DEV)
        ctx.reserve(n, mb)

        def sort_all():
            ctx.sort_keys(pos[0], pos[1], pos[2], (0, 0, 0), (L, L, L), (gd, gd, gd), n, keys)
            ctx.sort_by_key(keys, mb, vars_in, vars_out, n)

        med, best = timed(sort_all, reps=args.reps, flush=flush)
        alg = n * (2 * 24 + 12)
        print(json.dumps({"op": "agent_sort", "n": n, "max_bit": mb, "us_median": med, "us_best": best,
                          "alg_bytes": alg, "GBps": alg / med / 1e3, "frac_of_measured_peak": alg / med / 1e3 / peak}),
              flush=True)
        sp.close()


if __name__ == "__main__":
    main()
