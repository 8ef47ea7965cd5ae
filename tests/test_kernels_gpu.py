"""GPU parity tests of the C-ABI kernels against the CPU oracle (bit-exact integer/index work)."""
import numpy as np
import pytest
import torch

import oracle_py as orc

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ctx():
    from flamegpu2_b200 import host

    c = host.Context(0)
    yield c
    c.close()


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def as_u32(tensor):
    return tensor.cpu().numpy().view(np.uint32)


def _positions(n, lo, hi, seed, sorted_like=False, dims=3):
    rng = np.random.default_rng(seed)
    p = [rng.uniform(lo[a], hi[a], n).astype(np.float32) for a in range(dims)]
    return p


def _check_build(ctx, dims, mn, mx, radius, pos, stable, extra_len=0, expect_grouped=False):
    from flamegpu2_b200 import host

    n = len(pos[0])
    g = orc.Grid(dims, mn, mx, radius)
    sp = host.Spatial(ctx, dims, mn, mx, radius)
    assert sp.bin_count == g.bin_count and sp.grid_dim == g.grid_dim and sp.wrap_compatible == g.wrap_compatible
    ids = np.arange(n, dtype=np.uint32)
    names = ["id", "x", "y"] + (["z"] if dims == 3 else [])
    host_vars = {"id": ids, "x": pos[0], "y": pos[1]}
    if dims == 3:
        host_vars["z"] = pos[2]
    if extra_len:  # an array variable (test_spatial_3d.cu:568-616 ArrayVariable)
        arr = np.arange(n * extra_len, dtype=np.uint32).reshape(n, extra_len) * 3
        host_vars["v"] = arr
        names.append("v")
    ins = [t(host_vars[k]) for k in names]
    outs = [torch.zeros_like(a) for a in ins]
    zt = ins[names.index("z")] if dims == 3 else None
    sp.build_index(ins[names.index("x")], ins[names.index("y")], zt, ins, outs, n, stable=stable, expect_grouped=expect_grouped)
    torch.cuda.synchronize()
    pbm = sp.pbm()
    pbm_ref, perm_ref = g.build_index(pos[0], pos[1], pos[2] if dims == 3 else None)
    assert np.array_equal(pbm, pbm_ref), "PBM must be bit-exact"
    got_id = as_u32(outs[0])
    if stable:
        assert np.array_equal(got_id, perm_ref), "stable mode: source order inside every bin"
    else:
        # within-bin order is unspecified (reference: atomicInc arrival order): compare as multisets per bin
        a = np.sort(got_id.reshape(-1))
        assert np.array_equal(a, np.arange(n, dtype=np.uint32)), "every message appears exactly once"
        keys = g.bin_keys(pos[0], pos[1], pos[2] if dims == 3 else None)
        bins_of_sorted = np.repeat(np.arange(g.bin_count, dtype=np.uint32), np.diff(pbm.astype(np.int64)))
        assert np.array_equal(keys[got_id], bins_of_sorted), "every message sits inside its own bin"
    # every variable travelled with its message
    for k, o in zip(names, outs):
        hv = host_vars[k]
        assert np.array_equal(o.cpu().numpy().view(hv.dtype).reshape(hv.shape), hv[got_id]), k
    sp.close()


@pytest.mark.parametrize("dims", [3, 2])
@pytest.mark.parametrize("order", ["random", "bin_sorted", "mixed"])
@pytest.mark.parametrize("n", [1, 2049, 300007])
def test_build_index_expect_grouped_is_correct_for_any_order(ctx, n, order, dims):
    # FGB_BUILD_EXPECT_GROUPED is a performance hint: tiles that are NOT grouped are scattered inside the scan + scatter
    # launch (one atomic per message) instead of by the worklist launch; the result must be the same for any input order
    mn, mx, radius = [0.0] * dims, [30.0, 22.0, 17.0][:dims], 1.5
    pos = _positions(n, mn, mx, seed=7 * n + dims, dims=dims)
    if order != "random":
        g = orc.Grid(dims, mn, mx, radius)
        keys = g.bin_keys(pos[0], pos[1], pos[2] if dims == 3 else None)
        o = np.argsort(keys, kind="stable")
        if order == "mixed":  # half of the list bin-ordered, the rest behind it in random order (like ghosts appended by a halo)
            rest = np.arange(n)[np.isin(np.arange(n), o[: n // 2], invert=True)]
            o = np.concatenate([o[: n // 2], rest])
        pos = [p[o].copy() for p in pos]
    _check_build(ctx, dims, mn, mx, radius, pos, False, expect_grouped=True)
    # and repeated on one handler: the look-back words, counters and histogram come back clean without the worklist launch
    from flamegpu2_b200 import host

    g = orc.Grid(dims, mn, mx, radius)
    sp = host.Spatial(ctx, dims, mn, mx, radius)
    ins = [t(p) for p in pos]
    outs = [torch.zeros_like(a) for a in ins]
    pbm_ref, _ = g.build_index(pos[0], pos[1], pos[2] if dims == 3 else None)
    for k in range(3):
        sp.build_index(ins[0], ins[1], ins[2] if dims == 3 else None, ins, outs, n, expect_grouped=(k != 1))
        assert np.array_equal(sp.pbm(), pbm_ref)
    sp.close()


@pytest.mark.parametrize("stable", [False, True])
@pytest.mark.parametrize("n", [1, 5, 2049, 100003])
def test_build_index_3d(ctx, n, stable):
    pos = _positions(n, (-0.5, -0.5, -0.5), (10.5, 7.5, 5.5), seed=n)  # some points outside -> clamped
    _check_build(ctx, 3, (0, 0, 0), (10, 7, 5), 0.5, pos, stable)


@pytest.mark.parametrize("stable", [False, True])
def test_build_index_3d_nonfactor_and_array_var(ctx, stable):
    pos = _positions(30011, (0, 0, 0), (50.1, 50.1, 50.1), seed=7)
    _check_build(ctx, 3, (0, 0, 0), (50.1, 50.1, 50.1), 10.0, pos, stable, extra_len=3)


@pytest.mark.parametrize("stable", [False, True])
@pytest.mark.parametrize("n", [3, 2049, 250001])
def test_build_index_2d(ctx, n, stable):
    pos = _positions(n, (0, 0), (11, 11), seed=n + 1, dims=2)
    _check_build(ctx, 2, (0, 0), (11, 11), 1.0 if n < 10000 else 0.05, pos, stable)


@pytest.mark.parametrize("stable", [False, True])
def test_build_index_degenerate_single_bin(ctx, stable):
    # every message in one bin (big-bin path of the stable fix-up), and a second crowded bin
    n = 20000
    x = np.full(n, 1.25, np.float32)
    x[::3] = 7.75
    pos = [x, np.full(n, 2.5, np.float32), np.full(n, 0.5, np.float32)]
    _check_build(ctx, 3, (0, 0, 0), (10, 10, 10), 1.0, pos, stable)


def test_build_index_golden_mandatory(ctx, golden_dir):
    import os

    pos = np.fromfile(os.path.join(golden_dir, "mandatory3d_pos.f32"), dtype=np.float32).reshape(3, -1)
    _check_build(ctx, 3, (0, 0, 0), (5, 5, 5), 1.0, [pos[0], pos[1], pos[2]], True)
    _check_build(ctx, 3, (0, 0, 0), (5, 5, 5), 1.0, [pos[0], pos[1], pos[2]], False)


def test_build_index_empty_and_device_count(ctx):
    from flamegpu2_b200 import host

    sp = host.Spatial(ctx, 3, (0, 0, 0), (5, 5, 5), 1.0)
    n = 4096
    pos = _positions(n, (0, 0, 0), (5, 5, 5), seed=2)
    ins = [t(p) for p in pos]
    outs = [torch.zeros_like(a) for a in ins]
    sp.build_index(ins[0], ins[1], ins[2], ins, outs, n)
    assert sp.pbm()[-1] == n
    # ReadEmpty (test_spatial_3d.cu:507-537): empty list -> PBM all zero
    sp.build_index(ins[0], ins[1], ins[2], ins, outs, 0)
    assert not sp.pbm().any()
    # device-resident count smaller than the launch bound
    d_n = torch.tensor([1000], dtype=torch.int32, device=DEV)
    g = orc.Grid(3, (0, 0, 0), (5, 5, 5), 1.0)
    for stable in (False, True):
        sp.build_index(ins[0], ins[1], ins[2], ins, outs, n, d_n=d_n, stable=stable)
        pbm_ref, _ = g.build_index(pos[0][:1000], pos[1][:1000], pos[2][:1000])
        assert np.array_equal(sp.pbm(), pbm_ref)
    d_n.zero_()
    sp.build_index(ins[0], ins[1], ins[2], ins, outs, n, d_n=d_n)
    assert not sp.pbm().any()
    sp.close()


def test_build_index_repeatable_many_calls(ctx):
    # the histogram / look-back words must come back clean after every call
    from flamegpu2_b200 import host

    g = orc.Grid(3, (0, 0, 0), (20, 20, 20), 1.0)
    sp = host.Spatial(ctx, 3, (0, 0, 0), (20, 20, 20), 1.0)
    for it in range(6):
        n = 50000 + 1111 * it
        pos = _positions(n, (0, 0, 0), (20, 20, 20), seed=100 + it)
        ins = [t(p) for p in pos]
        outs = [torch.zeros_like(a) for a in ins]
        sp.build_index(ins[0], ins[1], ins[2], ins, outs, n, stable=bool(it & 1))
        pbm_ref, _ = g.build_index(*pos)
        assert np.array_equal(sp.pbm(), pbm_ref)
    sp.close()


@pytest.mark.parametrize("dims", [3, 2])
@pytest.mark.parametrize("n", [1, 2047, 2048, 6000, 300007])
def test_bin_permutation_global_and_tile_local(ctx, n, dims):
    # fgb_bin_permutation (b200 extension): the execution order of an agent function that reads a spatial list
    from flamegpu2_b200 import host

    mn, mx, radius = [0.0] * dims, [30.0, 22.0, 17.0][:dims], 1.5
    pos = _positions(n, mn, mx, seed=n + dims, dims=dims)
    g = orc.Grid(dims, mn, mx, radius)
    keys = g.bin_keys(pos[0], pos[1], pos[2] if dims == 3 else None)
    sp = host.Spatial(ctx, dims, mn, mx, radius)
    dp = [t(p) for p in pos]
    perm = torch.zeros(n, dtype=torch.int32, device=DEV)
    # global: sorted by bin, PBM = offsets
    sp.bin_permutation(dp[0], dp[1], dp[2] if dims == 3 else None, perm, n)
    torch.cuda.synchronize()
    pm = as_u32(perm)
    assert np.array_equal(np.sort(pm), np.arange(n, dtype=np.uint32))
    assert np.all(np.diff(keys[pm].astype(np.int64)) >= 0)
    pbm_ref, _ = g.build_index(pos[0], pos[1], pos[2] if dims == 3 else None)
    assert np.array_equal(sp.pbm(), pbm_ref)
    # tile local: a permutation inside every 2048-item tile, equal bins contiguous inside the tile
    perm.zero_()
    sp.bin_permutation(dp[0], dp[1], dp[2] if dims == 3 else None, perm, n, tile_local=True)
    torch.cuda.synchronize()
    pm = as_u32(perm)
    assert np.array_equal(np.sort(pm), np.arange(n, dtype=np.uint32))
    for t0 in range(0, n, 2048):
        seg = pm[t0:t0 + 2048]
        assert seg.min() >= t0 and seg.max() < min(t0 + 2048, n)
        k = keys[seg]
        runs = 1 + int(np.count_nonzero(np.diff(k.astype(np.int64))))
        assert runs == len(np.unique(k)), "every bin of the tile forms one contiguous run"
    sp.close()


@pytest.mark.parametrize("stable", [False, True])
@pytest.mark.parametrize("n,lower,upper,dist", [(1, 0, 1, "uniform"), (1024, 12, 12 + 512, "pairs"), (300007, -50, 4000, "uniform"),
                                               (250000, 7, 9, "uniform"), (100000, 0, 99999, "sorted")])
def test_bucket_build_index(ctx, n, lower, upper, dist, stable):
    # fgb_build_index_keys == MessageBucket::CUDAModelHandler::buildIndex: PBM bit-exact, every variable travels
    from flamegpu2_b200 import host

    rng = np.random.default_rng(n + upper)
    if dist == "pairs":
        keys = (12 + np.arange(n) // 2).astype(np.int32)  # the reference test's keys
    elif dist == "sorted":
        keys = np.sort(rng.integers(lower, upper + 1, n)).astype(np.int32)
    else:
        keys = rng.integers(lower, upper + 1, n).astype(np.int32)
    b = host.Bucket(ctx, lower, upper)
    assert b.bounds()[:2] == (lower, upper + 1)
    ids = np.arange(n, dtype=np.uint32)
    arr = (np.arange(n * 3, dtype=np.uint32).reshape(n, 3) * 7)
    ins = [t(keys), t(ids), t(arr)]
    outs = [torch.zeros_like(a) for a in ins]
    b.build_index(ins[0], ins, outs, n, stable=stable)
    torch.cuda.synchronize()
    pbm_ref, perm_ref = orc.bucket_build(lower, upper, keys)
    pbm = b.pbm()
    assert np.array_equal(pbm, pbm_ref), "PBM must be bit-exact"
    got = as_u32(outs[1])
    if stable:
        assert np.array_equal(got, perm_ref)
    else:
        assert np.array_equal(np.sort(got), ids)
        assert np.array_equal(keys[got] - lower, np.repeat(np.arange(upper - lower + 1), np.diff(pbm.astype(np.int64))))
    assert np.array_equal(outs[0].cpu().numpy(), keys[got]) and np.array_equal(as_u32(outs[2]).reshape(n, 3), arr[got])
    # empty list: PBM all zero
    b.build_index(ins[0], ins, outs, 0)
    torch.cuda.synchronize()
    assert not b.pbm().any()
    b.close()


@pytest.mark.parametrize("n", [0, 1, 255, 2048, 100003, 3000001])
def test_reduce_sum_min_max(ctx, n):
    # fgb_reduce == the reductions behind HostAgentAPI::sum/min/max; integer results exact, float sums accumulated
    # in double (tolerance: the rounding of the inputs' own sum order is not reproduced) and identical run to run
    rng = np.random.default_rng(n)
    cases = [(rng.normal(0, 100, n).astype(np.float32), False), (rng.normal(0, 1e6, n).astype(np.float64), False),
             (rng.integers(-2**31, 2**31 - 1, n).astype(np.int32), False), (rng.integers(0, 2**32 - 1, n).astype(np.uint32), True),
             (rng.integers(-2**40, 2**40, n).astype(np.int64), False)]
    for a, unsigned in cases:
        d = t(a.view(np.int32) if a.dtype == np.uint32 else a)
        if n == 0:
            d = torch.zeros(1, dtype=d.dtype, device=DEV)
        got = ctx.reduce(d, n, "sum", unsigned=unsigned)
        if a.dtype.kind == "f":
            ref = float(a.astype(np.float64).sum())
            assert abs(got - ref) <= 1e-9 * max(1.0, float(np.abs(a.astype(np.float64)).sum())), a.dtype
            assert ctx.reduce(d, n, "sum", unsigned=unsigned) == got, "reproducible"
        else:
            assert got == int(a.astype(object).sum()) if n else got == 0, a.dtype
        if n:
            assert ctx.reduce(d, n, "min", unsigned=unsigned) == a.min().item(), a.dtype
            assert ctx.reduce(d, n, "max", unsigned=unsigned) == a.max().item(), a.dtype
    # device-resident count smaller than the bound
    a = rng.integers(0, 1000, 5000).astype(np.int32)
    cnt = torch.tensor([1234], dtype=torch.int32, device=DEV)
    assert ctx.reduce(t(a), 5000, "sum", d_n=cnt) == int(a[:1234].sum())


@pytest.mark.parametrize("n", [1, 4095, 4096, 4097, 125001, 3000000])
def test_exclusive_scan(ctx, n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 50, n).astype(np.uint32)
    inp = t(a)
    out = torch.zeros(n + 1, dtype=torch.int32, device=DEV)
    ctx.exclusive_scan(inp, out, n)
    ref = np.concatenate([[0], np.cumsum(a.astype(np.uint64))]).astype(np.uint32)
    assert np.array_equal(as_u32(out), ref)


@pytest.mark.parametrize("n,frac", [(1, 1.0), (7, 0.5), (2048, 0.9), (2049, 0.0), (100001, 0.9), (1000003, 0.5)])
def test_compact_death(ctx, n, frac):
    # AgentDeath (test_cuda_simulation.cu:406-430) + AgentDeath_array (test_device_api.cu:15-58)
    rng = np.random.default_rng(n)
    flags = (rng.random(n) < frac).astype(np.uint32)
    x = rng.random(n).astype(np.float32)
    ids = np.arange(n, dtype=np.uint32)
    arr = rng.integers(0, 1 << 30, (n, 4)).astype(np.uint32)   # 16-byte array variable
    arr3 = rng.integers(0, 1 << 30, (n, 3)).astype(np.uint32)  # 12-byte array variable
    d64 = rng.random(n).astype(np.float64)
    b8 = rng.integers(0, 255, n).astype(np.uint8)
    ins = [t(x), t(ids), t(arr), t(arr3), t(d64), t(b8)]
    outs = [torch.zeros_like(a) for a in ins]
    cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
    tot = torch.zeros(1, dtype=torch.int32, device=DEV)
    ctx.compact(t(flags), ins, outs, n, d_out_count=cnt, d_out_total=tot)
    perm = orc.compact(flags, n)
    k = len(perm)
    assert int(cnt.item()) == k and int(tot.item()) == k
    for h, o in zip([x, ids, arr, arr3, d64, b8], outs):
        assert np.array_equal(o.cpu().numpy()[:k], h[perm]), "survivor order must be stable and bit-exact"
    # a second call on the same scratch (self-cleaning look-back words)
    ctx.compact(t(flags), ins[:2], outs[:2], n, invert=True, d_out_count=cnt)
    perm = orc.compact(flags, n, invert=True)
    assert int(cnt.item()) == len(perm)
    assert np.array_equal(outs[1].cpu().numpy().view(np.uint32)[: len(perm)], ids[perm])


def test_compact_keep_front_offset_and_device_args(ctx):
    # function-condition style: disabled agents at the front copied unconditionally (CUDAScatter.cu:82-83),
    # birth-append style: output offset and item count read from device words
    rng = np.random.default_rng(5)
    n, keep_front, off = 70001, 1234, 5000
    flags = (rng.random(n - keep_front) < 0.4).astype(np.uint32)
    ids = np.arange(n, dtype=np.uint32)
    out = torch.full((n + off,), -1, dtype=torch.int32, device=DEV)
    cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
    tot = torch.zeros(1, dtype=torch.int32, device=DEV)
    d_off = torch.tensor([off], dtype=torch.int32, device=DEV)
    d_n = torch.tensor([n - 777], dtype=torch.int32, device=DEV)
    ctx.compact(t(flags), [t(ids)], [out], n, keep_front=keep_front, d_out_offset=d_off, d_n=d_n, d_out_count=cnt,
                d_out_total=tot)
    perm = orc.compact(flags, n - 777, keep_front=keep_front)
    assert int(cnt.item()) == len(perm) and int(tot.item()) == off + len(perm)
    got = as_u32(out)
    assert np.array_equal(got[off: off + len(perm)], ids[perm])
    assert np.all(got[:off] == 0xFFFFFFFF) and np.all(got[off + len(perm):] == 0xFFFFFFFF)


def test_scatter_all_gather_broadcast(ctx):
    rng = np.random.default_rng(9)
    n = 33333
    a = rng.integers(0, 1 << 31, n).astype(np.uint32)
    b = rng.random((n, 3)).astype(np.float32)
    oa = torch.zeros(n + 10, dtype=torch.int32, device=DEV)
    ob = torch.zeros((n + 10, 3), dtype=torch.float32, device=DEV)
    ctx.scatter_all([t(a), t(b)], [oa, ob], n, out_offset=10)
    assert np.array_equal(as_u32(oa)[10:], a) and np.array_equal(ob.cpu().numpy()[10:], b)
    perm = rng.permutation(n).astype(np.uint32)
    ga = torch.zeros(n, dtype=torch.int32, device=DEV)
    gb = torch.zeros((n, 3), dtype=torch.float32, device=DEV)
    ctx.gather(t(perm), [t(a), t(b)], [ga, gb], n)
    assert np.array_equal(as_u32(ga), a[perm]) and np.array_equal(gb.cpu().numpy(), b[perm])
    # default-value broadcast for new agents (test_device_agent_creation.cu:604 default values)
    d1 = t(np.array([12.5], np.float32))
    d2 = t(np.array([1, 2, 3], np.uint32))
    o1 = torch.zeros(100, dtype=torch.float32, device=DEV)
    o2 = torch.zeros((100, 3), dtype=torch.int32, device=DEV)
    ctx.broadcast_init([d1, d2], [o1, o2], 60, out_offset=40)
    assert np.all(o1.cpu().numpy()[40:] == 12.5) and np.all(o1.cpu().numpy()[:40] == 0)
    assert np.all(o2.cpu().numpy()[40:] == [1, 2, 3])


@pytest.mark.parametrize("true3d", [False, True])
@pytest.mark.parametrize("n", [4, 1000, 200003])
def test_agent_sort(ctx, n, true3d):
    # auto-sort: key kernel bit-exact (including the reference's collapsed-z quirk for 3D lists,
    # CUDASimulation.cu:487), stable order bit-exact (test_spatial_agent_sort.cu:69-111)
    g = orc.Grid(3, (-5, -5, -5), (5, 5, 5), 0.2)
    rng = np.random.default_rng(n)
    if n == 4:
        p = -np.arange(4, dtype=np.float32)
        pos = [p, p.copy(), p.copy()]
    else:
        pos = [rng.uniform(-5.3, 5.3, n).astype(np.float32) for _ in range(3)]  # out-of-range keys wrap in uint
    keys_ref = g.sort_keys(*pos, true3d=true3d)
    keys = torch.zeros(n, dtype=torch.int32, device=DEV)
    ins = [t(p) for p in pos]
    mn, w, gd = g.sort_geometry(true3d)
    ctx.sort_keys(ins[0], ins[1], ins[2], mn, w, gd, n, keys)
    assert np.array_equal(as_u32(keys), keys_ref)
    mb = g.sort_max_bit(true3d)
    perm_ref = orc.sort_perm(keys_ref, mb)
    order = np.arange(n, dtype=np.uint32)
    ins2 = ins + [t(order)]
    outs = [torch.zeros_like(a) for a in ins2]
    pos_out = torch.zeros(n, dtype=torch.int32, device=DEV)
    ctx.sort_by_key(keys, mb, ins2, outs, n, position_out=pos_out)
    assert np.array_equal(as_u32(pos_out), perm_ref)
    assert np.array_equal(as_u32(outs[3]), order[perm_ref])
    assert np.array_equal(outs[0].cpu().numpy(), pos[0][perm_ref])
    if n == 4:
        assert list(as_u32(outs[3])) == [3, 2, 1, 0]
    # the fused entry point (keys computed inside the histogram pass) gives the same keys and the same order
    keys2 = torch.zeros(n, dtype=torch.int32, device=DEV)
    outs2 = [torch.zeros_like(a) for a in ins2]
    pos2 = torch.zeros(n, dtype=torch.int32, device=DEV)
    ctx.sort_spatial(ins[0], ins[1], ins[2], mn, w, gd, mb, keys2, ins2, outs2, n, position_out=pos2)
    assert np.array_equal(as_u32(keys2), keys_ref) and np.array_equal(as_u32(pos2), perm_ref)
    assert np.array_equal(as_u32(outs2[3]), order[perm_ref])
    # compaction right after a sort on the same scratch slot
    flags = (rng.random(n) < 0.5).astype(np.uint32)
    cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
    o = torch.zeros(n, dtype=torch.int32, device=DEV)
    ctx.compact(t(flags), [t(order)], [o], n, d_out_count=cnt)
    assert int(cnt.item()) == int(flags.sum())


def test_agent_sort_2d_and_many_ties(ctx):
    g = orc.Grid(2, (0, 0), (8, 8), 1.0)
    n = 50000
    rng = np.random.default_rng(77)
    pos = [rng.uniform(0, 8, n).astype(np.float32) for _ in range(2)]
    keys_ref = g.sort_keys(pos[0], pos[1])
    keys = torch.zeros(n, dtype=torch.int32, device=DEV)
    mn, w, gd = g.sort_geometry()
    ctx.sort_keys(t(pos[0]), t(pos[1]), None, mn, w, gd, n, keys)
    assert np.array_equal(as_u32(keys), keys_ref)
    mb = g.sort_max_bit()
    order = np.arange(n, dtype=np.uint32)
    out = torch.zeros(n, dtype=torch.int32, device=DEV)
    ctx.sort_by_key(keys, mb, [t(order)], [out], n)   # ~780 agents per bin: big-bin fix-up path
    assert np.array_equal(as_u32(out), orc.sort_perm(keys_ref, mb))


# ---- the remaining CUDAScatter kernels ------------------------------------------------------------
@pytest.mark.parametrize("n,length", [(1, 1), (1000, 1000), (70001, 100000), (300000, 300000)])
def test_array_message_reorder(ctx, n, length):
    """reorder_array_messages (CUDAScatter.cu:540-566): out[index[i]] = in[i]; unique indices -> max writes 1"""
    rng = np.random.default_rng(n)
    index = rng.permutation(length)[:n].astype(np.int32)
    v4 = rng.integers(0, 2**31, n).astype(np.int32)
    v12 = rng.integers(0, 2**31, (n, 3)).astype(np.int32)
    ins = [torch.from_numpy(a).to(DEV) for a in (v4, v12)]
    outs = [torch.full((length,), -1, dtype=torch.int32, device=DEV), torch.full((length, 3), -1, dtype=torch.int32, device=DEV)]
    wc = torch.zeros(length, dtype=torch.int32, device=DEV)
    mx = torch.zeros(1, dtype=torch.int32, device=DEV)
    ctx.array_reorder(torch.from_numpy(index).to(DEV), length, ins, outs, n, write_count=wc, d_max=mx)
    exp4 = np.full(length, -1, np.int32)
    exp12 = np.full((length, 3), -1, np.int32)
    exp4[index] = v4
    exp12[index] = v12
    assert np.array_equal(outs[0].cpu().numpy(), exp4) and np.array_equal(outs[1].cpu().numpy(), exp12)
    assert int(mx.item()) == 1 and int(wc.sum().item()) == 0, "write counters are folded to their maximum and re-zeroed"


def test_array_message_reorder_conflict_and_out_of_bounds(ctx):
    """two messages for one element -> max writes 2 (the reference raises ArrayMessageWriteConflict, CUDAScatter.cu:629-651);
    an index beyond the array is dropped (:553-554); more messages than elements is rejected (:579-581)"""
    index = torch.tensor([0, 5, 5, 9, 12], dtype=torch.int32, device=DEV)
    val = torch.arange(5, dtype=torch.int32, device=DEV)
    out = torch.full((10,), -1, dtype=torch.int32, device=DEV)
    wc = torch.zeros(10, dtype=torch.int32, device=DEV)
    mx = torch.zeros(1, dtype=torch.int32, device=DEV)
    ctx.array_reorder(index, 10, [val], [out], 5, write_count=wc, d_max=mx)
    o = out.cpu().numpy()
    assert int(mx.item()) == 2 and o[0] == 0 and o[9] == 3 and o[5] in (1, 2) and (o[[1, 2, 3, 4, 6, 7, 8]] == -1).all()
    with pytest.raises(Exception):
        ctx.array_reorder(index, 4, [val], [out], 5)


@pytest.mark.parametrize("n", [1, 255, 256, 1000, 100003])
def test_scatter_new_agents_aos_to_soa(ctx, n):
    """scatter_new_agents (CUDAScatter.cu:348-366): structs {f32 x; u32 id; i32 arr[3]; u8 flag; pad} appended after 7 agents"""
    agent_size = 24
    rng = np.random.default_rng(n)
    raw = rng.integers(0, 256, (n, agent_size)).astype(np.uint8)
    aos = torch.from_numpy(raw).to(DEV)
    offsets, lens = [0, 4, 8, 20], [4, 4, 12, 1]
    outs = [torch.zeros(n + 7, dtype=torch.int32, device=DEV), torch.zeros(n + 7, dtype=torch.int32, device=DEV),
            torch.zeros((n + 7, 3), dtype=torch.int32, device=DEV), torch.zeros(n + 7, dtype=torch.uint8, device=DEV)]
    off = torch.tensor([7], dtype=torch.int32, device=DEV)
    ctx.scatter_new_agents(aos, agent_size, offsets, lens, outs, n, d_out_offset=off)
    for o, ofs, ln in zip(outs, offsets, lens):
        got = o.cpu().numpy().view(np.uint8).reshape(n + 7, ln)
        assert not got[:7].any(), "existing agents are untouched"
        assert np.array_equal(got[7:], raw[:, ofs:ofs + ln])


@pytest.mark.parametrize("n", [0, 1, 1000, 2000003])
def test_histogram_even(ctx, n):
    """cub::DeviceHistogram::HistogramEven semantics as HostAgentAPI::histogramEven uses them (test_histogram_even.cu:50-64)"""
    rng = np.random.default_rng(n + 1)
    f = rng.uniform(-2, 22, n).astype(np.float32)
    h = ctx.histogram_even(torch.from_numpy(f).to(DEV), n, 10, 0.0, 20.0).cpu().numpy()
    inside = f[(f >= 0) & (f < 20)]
    exp = np.bincount(((inside - np.float32(0)) * (np.float32(10) / np.float32(20))).astype(np.int32), minlength=10)[:10]
    assert np.array_equal(h, exp)
    i = rng.integers(-5, 30, n).astype(np.int32)
    h = ctx.histogram_even(torch.from_numpy(i).to(DEV), n, 10, 0, 20).cpu().numpy()
    inside = i[(i >= 0) & (i < 20)]
    assert np.array_equal(h, np.bincount(inside * 10 // 20, minlength=10)[:10])
    hb = ctx.histogram_even(torch.from_numpy(i).to(DEV), n, 5000, -5, 30).cpu().numpy()  # more bins than the shared-memory path holds
    assert hb.sum() == n
