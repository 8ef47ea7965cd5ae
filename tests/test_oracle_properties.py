"""Property tests of the CPU oracle (oracle/fgb_oracle.c) against independent numpy restatements: the oracle is what
the CUDA path is judged against, so it is itself cross-checked from a second, differently written implementation.
CPU only; sizes include empty and ragged inputs."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle_py as orc

SET = dict(max_examples=40, deadline=None)


@settings(**SET)
@given(n=st.integers(0, 3000), dims=st.sampled_from([2, 3]), radius=st.sampled_from([0.5, 1.0, 1.7]), seed=st.integers(0, 2**31 - 1))
def test_build_index_is_a_stable_counting_sort(n, dims, radius, seed):
    rng = np.random.default_rng(seed)
    mn, mx = [-1.0, 0.0, 2.0][:dims], [6.3, 5.0, 7.9][:dims]
    pos = [rng.uniform(mn[a] - 0.5, mx[a] + 0.5, n).astype(np.float32) for a in range(dims)]  # some points outside: clamped
    g = orc.Grid(dims, mn, mx, radius)
    keys = g.bin_keys(*pos)
    # independent restatement of getGridPosition / getHash (MessageSpatial3DDevice.cuh:646-672) in numpy float32
    cell = []
    for a in range(dims):
        c = np.floor((pos[a] - np.float32(mn[a])) / np.float32(radius)).astype(np.int64)
        cell.append(np.clip(c, 0, g.grid_dim[a] - 1))
    ref = cell[0] + g.grid_dim[0] * cell[1] + (g.grid_dim[0] * g.grid_dim[1] * cell[2] if dims == 3 else 0)
    assert np.array_equal(keys, ref.astype(np.uint32))
    pbm, perm = g.build_index(*pos)
    assert np.array_equal(perm, np.argsort(keys, kind="stable").astype(np.uint32))
    assert np.array_equal(pbm, np.concatenate([[0], np.cumsum(np.bincount(keys, minlength=g.bin_count))]).astype(np.uint32))


@settings(**SET)
@given(n=st.integers(0, 5000), frac=st.floats(0.0, 1.0), invert=st.booleans(), keep_front=st.integers(0, 50), seed=st.integers(0, 2**31 - 1))
def test_compaction_is_stable_and_honours_keep_front(n, frac, invert, keep_front, seed):
    rng = np.random.default_rng(seed)
    keep_front = min(keep_front, n)
    flags = (rng.random(max(n - keep_front, 0)) < frac).astype(np.uint32)  # flags cover the items AFTER the kept front
    got = orc.compact(flags, n, invert=invert, keep_front=keep_front)
    kept = (flags == 1) != invert
    ref = np.concatenate([np.arange(keep_front), keep_front + np.nonzero(kept)[0]]).astype(np.uint32)
    assert np.array_equal(got, ref)


@settings(**SET)
@given(n=st.integers(0, 4000), max_bit=st.integers(1, 20), seed=st.integers(0, 2**31 - 1))
def test_sort_perm_is_a_stable_sort_on_the_low_bits(n, max_bit, seed):
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    perm = orc.sort_perm(keys, max_bit)
    masked = keys & np.uint32((1 << max_bit) - 1)
    assert np.array_equal(perm, np.argsort(masked, kind="stable").astype(np.uint32))


@settings(**SET)
@given(n=st.integers(0, 3000), lower=st.integers(-100, 100), span=st.integers(1, 400), seed=st.integers(0, 2**31 - 1))
def test_bucket_build_and_ranges(n, lower, span, seed):
    rng = np.random.default_rng(seed)
    upper = lower + span
    keys = rng.integers(lower, upper + 1, n).astype(np.int32)
    pbm, perm = orc.bucket_build(lower, upper, keys)
    assert np.array_equal(perm, np.argsort(keys, kind="stable").astype(np.uint32))
    assert np.array_equal(pbm, np.concatenate([[0], np.cumsum(np.bincount(keys - lower, minlength=span + 1))]).astype(np.uint32))
    for _ in range(20):
        b = int(rng.integers(lower - 2, upper + 3))
        e = int(rng.integers(lower - 2, upper + 4))
        first, cnt = orc.bucket_range(lower, upper, pbm, b, e)
        if b >= lower and e < upper + 1 and b <= e:  # the reference's acceptance test (MessageBucketDevice.cuh:269)
            assert cnt == int(((keys >= b) & (keys < e)).sum()) and first == int((keys < b).sum())
        else:
            assert cnt == 0


@settings(max_examples=15, deadline=None)
@given(n=st.integers(1, 500), dims=st.sampled_from([2, 3]), seed=st.integers(0, 2**31 - 1))
def test_filter_visits_exactly_the_moore_neighbourhood(n, dims, seed):
    # In::Filter (MessageSpatial3DDevice.cuh:693-719): every message whose cell differs by at most 1 in each axis,
    # each exactly once -- checked against an O(n^2) numpy restatement
    rng = np.random.default_rng(seed)
    mn, mx, radius = [0.0] * dims, [5.0, 4.0, 3.0][:dims], 1.0
    pos = [rng.uniform(mn[a], mx[a], n).astype(np.float32) for a in range(dims)]
    g = orc.Grid(dims, mn, mx, radius)
    pbm, perm = g.build_index(*pos)
    sp = [p[perm] for p in pos]  # the bin-sorted list the iterator walks
    cells = np.stack([np.clip(np.floor(p / np.float32(radius)).astype(np.int64), 0, g.grid_dim[a] - 1) for a, p in enumerate(sp)])
    for i in rng.choice(n, size=min(n, 12), replace=False):
        q = [float(p[i]) for p in pos]
        got = g.filter(pbm, *q)
        c = [int(np.clip(np.floor(np.float32(q[a]) / np.float32(radius)), 0, g.grid_dim[a] - 1)) for a in range(dims)]
        near = np.all(np.abs(cells - np.array(c)[:, None]) <= 1, axis=0)
        assert len(got) == len(set(got.tolist())) and set(got.tolist()) == set(np.nonzero(near)[0].tolist())


@settings(max_examples=10, deadline=None)
@given(n=st.integers(2, 400), seed=st.integers(0, 2**31 - 1))
def test_neighbour_count_and_circles_step_against_brute_force(n, seed):
    rng = np.random.default_rng(seed)
    L, radius, repulse = 6.0, 2.0, 0.05
    x, y, z = (rng.uniform(0, L, n).astype(np.float32) for _ in range(3))
    ids = np.arange(1, n + 1, dtype=np.uint32)
    g = orc.Grid(3, (0, 0, 0), (L, L, L), radius)
    pbm, perm = g.build_index(x, y, z)
    # integer neighbour count (stress model): separately rounded products and sums, strict '<'
    nb = g.neighbour_count(pbm, ids[perm], x[perm], y[perm], z[perm], ids, x, y, z)
    dx, dy, dz = (a[None, :] - a[:, None] for a in (x, y, z))  # [agent, message] = message - agent
    d2 = ((dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)).astype(np.float32) + (dz * dz).astype(np.float32)
    inside = (d2.astype(np.float32) < np.float32(radius) * np.float32(radius)) & ~np.eye(n, dtype=bool)
    assert np.array_equal(nb, inside.sum(axis=1).astype(np.uint32))
    # Circles move (examples/circles_model.cuh, reference main.cu:5-54) without the sort: float32 brute force, the
    # summation order differs from the bin walk, hence a tolerance
    _, x2, y2, z2, drift, _ = g.circles_step(ids, x, y, z, np.zeros(n, np.float32), repulse=repulse, do_sort=False)
    sep = np.sqrt((dx * dx + dy * dy + dz * dz).astype(np.float32))
    act = (sep < radius) & (sep > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where(act, np.sin((sep / np.float32(radius)) * np.float32(3.141) * np.float32(-2)) * np.float32(repulse), 0).astype(np.float32)
        fx, fy, fz = (np.where(act, k * d / sep, 0).sum(axis=1) for d in (dx, dy, dz))
    cnt = np.maximum(act.sum(axis=1), 1)
    fx, fy, fz = fx / cnt, fy / cnt, fz / cnt
    assert np.allclose(x2, x + fx, rtol=1e-5, atol=1e-5) and np.allclose(y2, y + fy, rtol=1e-5, atol=1e-5)
    assert np.allclose(z2, z + fz, rtol=1e-5, atol=1e-5)
    assert np.allclose(drift, np.sqrt(fx * fx + fy * fy + fz * fz), rtol=1e-4, atol=1e-5)
