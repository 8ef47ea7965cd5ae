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
