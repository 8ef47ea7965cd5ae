"""GPU parity tests of whole CUDASimulation::step() runs through the C++ API layer
(include/flamegpu + examples/*.cuh, driven through libfgb_models.so) against the CPU oracle and
against the known answers of the reference's own tests (citations into /root/reference)."""
import os

import numpy as np
import pytest

import oracle_py as orc

pytestmark = pytest.mark.gpu

# float agent state: one step from a common state, differences come only from the summation order
# over ~33 in-radius neighbours and FMA contraction / libm-vs-CUDA sinf (<= 2 ulp):  33 * 2^-24 ~ 2e-6
RTOL, ATOL = 1e-5, 2e-6


def _sim(model, **kw):
    from flamegpu2_b200 import sim as fsim

    return fsim.Simulation(model, **kw)


def _circles_pop(n, L, seed):
    rng = np.random.default_rng(seed)
    return [rng.uniform(0, L, n).astype(np.float32) for _ in range(3)]


@pytest.mark.parametrize("graphs", [1, 0])
@pytest.mark.parametrize("n,L", [(16384, 25.0), (100000, 46.0)])
def test_circles_steps_vs_oracle(n, L, graphs):
    pos = _circles_pop(n, L, seed=n)
    g = orc.Grid(3, (0, 0, 0), (L, L, L), 2.0)
    s = _sim("circles", env_max=L, radius=2.0, graphs=graphs)
    s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
    ids = np.arange(1, n + 1, dtype=np.uint32)
    x, y, z, d = pos[0], pos[1], pos[2], np.zeros(n, np.float32)
    for step in range(4):
        # teacher forcing: the oracle always starts from OUR previous state
        i_ref, x_ref, y_ref, z_ref, d_ref, pbm_ref = g.circles_step(ids, x, y, z, d, want_pbm=True)
        s.step(1)
        ids = s.get("Circle", "_id", np.uint32)
        x, y, z, d = (s.get("Circle", v, np.float32) for v in ("x", "y", "z", "drift"))
        assert np.array_equal(ids, i_ref), f"step {step}: agent order after the auto-sort must be bit-exact"
        assert np.array_equal(s.message_pbm("location"), pbm_ref), f"step {step}: PBM must be bit-exact"
        for a, b, nm in ((x, x_ref, "x"), (y, y_ref, "y"), (z, z_ref, "z")):
            assert np.allclose(a, b, rtol=RTOL, atol=ATOL), f"step {step}: {nm}"
        assert np.allclose(d, d_ref, rtol=1e-3, atol=ATOL), f"step {step}: drift"
        # the sorted message list holds every message exactly once, grouped by bin
        mid = s.message_variable("location", "id", np.uint32, n)
        assert np.array_equal(np.sort(mid), np.arange(1, n + 1, dtype=np.uint32))
    assert s.step_counter == 4
    if graphs:
        assert 1 <= s.graphs <= 3, "double-buffer parity gives at most a couple of distinct step graphs"
    s.close()


def test_circles_stable_order_is_deterministic():
    n, L = 50000, 36.0
    pos = _circles_pop(n, L, seed=5)
    res = []
    for _ in range(2):
        s = _sim("circles", env_max=L, radius=2.0, stable=1)
        s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
        s.step(5)
        res.append([s.get("Circle", v, np.float32) for v in ("x", "y", "z", "drift")] + [s.get("Circle", "_id", np.uint32)])
        s.close()
    for a, b in zip(*res):
        assert np.array_equal(a, b), "stable message order must make runs bit-reproducible"


def test_circles_no_sort_matches_oracle():
    n, L = 20000, 27.0
    pos = _circles_pop(n, L, seed=9)
    g = orc.Grid(3, (0, 0, 0), (L, L, L), 2.0)
    s = _sim("circles", env_max=L, radius=2.0, sort_period=0)  # setSortPeriod(0): SortingDisabled
    s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
    s.step(1)
    ids = s.get("Circle", "_id", np.uint32)
    assert np.array_equal(ids, np.arange(1, n + 1, dtype=np.uint32))
    i_ref, x_ref, *_ = g.circles_step(np.arange(1, n + 1, dtype=np.uint32), pos[0], pos[1], pos[2], np.zeros(n, np.float32),
                                      do_sort=False)
    assert np.allclose(s.get("Circle", "x", np.float32), x_ref, rtol=RTOL, atol=ATOL)
    s.close()


def test_reference_mandatory3d(golden_dir):
    # Spatial3DMessageTest.Mandatory (test_spatial_3d.cu:72-205) on its own seeded population
    pos = np.fromfile(os.path.join(golden_dir, "mandatory3d_pos.f32"), dtype=np.float32).reshape(3, -1)
    expect = np.fromfile(os.path.join(golden_dir, "mandatory3d_expect.u32"), dtype=np.uint32)
    s = _sim("test", which=0, max_x=5, max_y=5, max_z=5, radius=1, sort_period=0)
    s.set_population("agent", {"x": pos[0], "y": pos[1], "z": pos[2]})
    s.step(1)
    assert np.array_equal(s.get("agent", "count", np.uint32), expect)
    assert not s.get("agent", "badCount", np.uint32).any()
    s.close()
    # and with the automatic sort on: same counts per agent id
    s = _sim("test", which=0, max_x=5, max_y=5, max_z=5, radius=1, sort_period=1)
    s.set_population("agent", {"x": pos[0], "y": pos[1], "z": pos[2]})
    s.step(1)
    ids = s.get("agent", "_id", np.uint32)
    assert np.array_equal(s.get("agent", "count", np.uint32), expect[ids - 1])
    s.close()


def test_reference_optional3d(golden_dir):
    # Spatial3DMessageTest.Optional / OptionalNone (test_spatial_3d.cu:207-446): only flagged agents output
    pos = np.fromfile(os.path.join(golden_dir, "mandatory3d_pos.f32"), dtype=np.float32).reshape(3, -1)
    n = pos.shape[1]
    rng = np.random.default_rng(4)
    g = orc.Grid(3, (0, 0, 0), (5, 5, 5), 1.0)
    for frac in (0.8, 0.0):
        do_out = (rng.random(n) < frac).astype(np.int32)
        s = _sim("test", which=1, max_x=5, max_y=5, max_z=5, radius=1, sort_period=0)
        s.set_population("agent", {"x": pos[0], "y": pos[1], "z": pos[2], "do_output": do_out})
        s.step(1)
        sel = np.nonzero(do_out)[0]
        pbm, perm = g.build_index(pos[0][sel], pos[1][sel], pos[2][sel])
        expect = np.array([len(g.filter(pbm, pos[0][i], pos[1][i], pos[2][i])) for i in range(n)], dtype=np.uint32)
        assert np.array_equal(s.get("agent", "count", np.uint32), expect)
        ids_sorted = (sel + 1).astype(np.uint32)[perm]
        idsum = np.array([ids_sorted[g.filter(pbm, pos[0][i], pos[1][i], pos[2][i])].sum(dtype=np.uint32) for i in range(n)],
                         dtype=np.uint32)
        assert np.array_equal(s.get("agent", "idsum", np.uint32), idsum)
        assert np.array_equal(s.message_pbm("location"), pbm)
        s.close()


@pytest.mark.parametrize("off", [(0.0, 0.0, 0.0), (141.0, -540.0, 200.0), (-1401.5, 5640.3, -2008.8)])
def test_reference_wrapped3d(off):
    # Spatial3DMessageTest.Wrapped/2/3 (test_spatial_3d.cu:855-975)
    off = np.array(off, dtype=np.float32)
    mx = (off + np.float32(70)).astype(np.float32)
    ii, jj, kk = np.meshgrid(np.arange(35), np.arange(35), np.arange(35), indexing="ij")
    x = (ii.ravel() * np.float32(2.0) + off[0]).astype(np.float32)
    y = (jj.ravel() * np.float32(2.0) + off[1]).astype(np.float32)
    z = (kk.ravel() * np.float32(2.0) + off[2]).astype(np.float32)
    s = _sim("test", which=2, min_x=float(off[0]), min_y=float(off[1]), min_z=float(off[2]), max_x=float(mx[0]),
             max_y=float(mx[1]), max_z=float(mx[2]), radius=3.5)
    s.set_population("agent", {"x": x, "y": y, "z": z})
    s.step(1)
    assert np.all(s.get("agent", "count", np.uint32) == 27)
    assert np.all(s.get("agent", "badCount", np.uint32) <= 189)
    for v in ("result_x", "result_y", "result_z"):
        assert np.all(s.get("agent", v, np.float32) == 0.0)
    s.close()


def test_reference_2d(golden_dir):
    # Spatial2DMessageTest.Mandatory (test_spatial_2d.cu:66-180) and Wrapped
    pos = np.fromfile(os.path.join(golden_dir, "mandatory2d_pos.f32"), dtype=np.float32).reshape(2, -1)
    expect = np.fromfile(os.path.join(golden_dir, "mandatory2d_expect.u32"), dtype=np.uint32)
    s = _sim("test", which=3, max_x=11, max_y=11, radius=1, sort_period=0)
    s.set_population("agent", {"x": pos[0], "y": pos[1]})
    s.step(1)
    assert np.array_equal(s.get("agent", "count", np.uint32), expect)
    assert not s.get("agent", "badCount", np.uint32).any()
    s.close()
    ii, jj = np.meshgrid(np.arange(35), np.arange(35), indexing="ij")
    s = _sim("test", which=4, max_x=70, max_y=70, radius=3.5)
    s.set_population("agent", {"x": (ii.ravel() * 2.0).astype(np.float32), "y": (jj.ravel() * 2.0).astype(np.float32)})
    s.step(1)
    assert np.all(s.get("agent", "count", np.uint32) == 9)
    assert np.all(s.get("agent", "result_x", np.float32) == 0.0) and np.all(s.get("agent", "result_y", np.float32) == 0.0)
    s.close()


def test_reference_bounds_not_factor_radius():
    # Spatial3DMessageTest.bounds_not_factor_radius (test_spatial_3d.cu:1111-1188)
    h = np.float32(50.1) / np.float32(2)
    pts = np.array([[h, 0.0, h], [h, 18.0, h], [0.0, h, 0.5], [18.0, h, 0.5], [h, 50.0, np.float32(50.1)],
                    [h, 50.0, np.float32(50.1) - np.float32(10.11)]], dtype=np.float32)
    s = _sim("test", which=0, max_x=50.1, max_y=50.1, max_z=50.1, radius=10, sort_period=0)
    s.set_population("agent", {"x": pts[:, 0].copy(), "y": pts[:, 1].copy(), "z": pts[:, 2].copy()})
    s.step(1)
    assert list(s.get("agent", "count", np.uint32)) == [2, 2, 2, 2, 1, 1]
    s.close()


def test_reference_buffer_not_init_and_read_empty():
    # test_spatial_3d.cu:507-537 ReadEmpty / :1066-1097 buffer_not_init: reading a list nobody wrote
    s = _sim("test", which=1, max_x=5, max_y=5, max_z=5, radius=1)
    n = 100
    rng = np.random.default_rng(0)
    s.set_population("agent", {"x": rng.uniform(0, 5, n).astype(np.float32), "y": rng.uniform(0, 5, n).astype(np.float32),
                               "z": rng.uniform(0, 5, n).astype(np.float32), "do_output": np.zeros(n, np.int32)})
    s.step(2)
    assert not s.get("agent", "count", np.uint32).any()
    assert s.message_count("location") == 0 and not s.message_pbm("location").any()
    s.close()


def test_reference_agent_death_order():
    # TestCUDASimulation.AgentDeath (test_cuda_simulation.cu:406-430) + AgentDeath_array (test_device_api.cu:15-58)
    n = 1024
    rng = np.random.default_rng(12)
    x = rng.integers(0, 13, n).astype(np.uint32)
    arr = np.stack([x + 1, x + 2, x + 3], axis=1).astype(np.uint32)
    s = _sim("test", which=5)
    s.set_population("agent", {"x": x, "arr": arr})
    s.step(1)
    keep = x % 2 != 0
    assert s.count("agent") == int(keep.sum())
    assert np.array_equal(s.get("agent", "x", np.uint32), x[keep] + 12), "survivors keep their original order"
    assert np.array_equal(s.get("agent", "arr", np.uint32, 3), arr[keep])
    assert np.array_equal(s.get("agent", "_id", np.uint32), np.arange(1, n + 1, dtype=np.uint32)[keep])
    s.close()


@pytest.mark.parametrize("which,name", [(6, "mandatory"), (7, "optional"), (8, "optional_death"), (9, "other_agent")])
def test_reference_device_agent_creation(which, name):
    # DeviceAgentCreationTest.* (test_device_agent_creation.cu:53-1130): sizes, value multisets, defaults, unique ids
    n = 1024
    ids0 = np.arange(n, dtype=np.uint32)
    s = _sim("test", which=which)
    s.set_population("agent", {"x": (ids0 + 1.0).astype(np.float32), "id": ids0})
    s.step(1)
    uid = ids0 + 1
    if name == "mandatory":
        assert s.count("agent") == 2 * n
        x, idv = s.get("agent", "x", np.float32), s.get("agent", "id", np.uint32)
        assert np.array_equal(idv[:n], ids0) and np.array_equal(x[:n], ids0 + 1.0)       # parents first, in order
        assert np.array_equal(idv[n:], uid) and np.array_equal(x[n:], uid + 12.0)        # children in parent order
        assert np.all(s.get("agent", "untouched", np.float32) == 15.0)                     # defaults of unset variables
        aid = s.get("agent", "_id", np.uint32)
        assert len(np.unique(aid)) == 2 * n and set(aid[n:]) == set(range(n + 1, 2 * n + 1))
    elif name == "optional":
        sel = uid % 2 == 1
        assert s.count("agent") == n + int(sel.sum())
        idv = s.get("agent", "id", np.uint32)
        assert np.array_equal(idv[n:], uid[sel])
        assert np.array_equal(s.get("agent", "x", np.float32)[n:], uid[sel] + 12.0)
    elif name == "optional_death":
        sel = uid % 2 == 1  # these parents survive and give birth; the others die
        assert s.count("agent") == 2 * int(sel.sum())
        idv = s.get("agent", "id", np.uint32)
        k = int(sel.sum())
        assert np.array_equal(idv[:k], ids0[sel]) and np.array_equal(idv[k:], uid[sel])
        assert len(np.unique(s.get("agent", "_id", np.uint32))) == 2 * k
    else:
        assert s.count("agent") == n and s.count("agent2") == n
        assert np.array_equal(s.get("agent2", "id", np.uint32), uid)
        assert np.array_equal(s.get("agent2", "x", np.float32), uid + 12.0)
        assert np.all(s.get("agent2", "untouched", np.float32) == 15.0)
        assert np.array_equal(np.sort(s.get("agent2", "_id", np.uint32)), np.arange(1, n + 1, dtype=np.uint32))
    # a second step keeps working (bounds refreshed, graph re-used or re-captured)
    s.step(1)
    if name == "mandatory":
        assert s.count("agent") == 4 * n
    s.close()


def test_stress_birth_death_vs_oracle():
    # BASELINE config 3 in miniature: death + birth + spatial messages in one function, several steps
    n, L = 30000, 31.0
    pos = _circles_pop(n, L, seed=21)
    g = orc.Grid(3, (0, 0, 0), (L, L, L), 2.0)
    s = _sim("stress", env_max=L, radius=2.0, death_mod=10, birth_mod=20)
    s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
    ids = np.arange(1, n + 1, dtype=np.uint32)
    x, y, z = pos
    next_id = n + 1
    for step in range(3):
        # oracle step from our state
        keys = g.sort_keys(x, y, z)
        perm = orc.sort_perm(keys, g.sort_max_bit())
        pbm, mperm = g.build_index(x, y, z)                  # messages were output before the sort
        sid, sx, sy, sz = ids[perm], x[perm], y[perm], z[perm]
        nb = g.neighbour_count(pbm, ids[mperm], x[mperm], y[mperm], z[mperm], sid, sx, sy, sz)
        die = np.array([orc.hash32(int(i), step) % 10 == 0 for i in sid])
        born = np.array([orc.hash32(int(i) ^ 0x9E3779B9, step) % 20 == 0 for i in sid])
        s.step(1)
        cnt = s.count("Circle")
        k, b = int((~die).sum()), int(born.sum())
        assert cnt == k + b, f"step {step}"
        ids2 = s.get("Circle", "_id", np.uint32)
        x2, y2, z2 = (s.get("Circle", v, np.float32) for v in ("x", "y", "z"))
        assert np.array_equal(ids2[:k], sid[~die]), "survivors: stable, bit-exact order"
        assert np.array_equal(s.get("Circle", "neighbours", np.uint32)[:k], nb[~die]), "integer state bit-exact"
        assert np.array_equal(x2[:k], sx[~die])
        # children: appended after the survivors in parent order; ids are a set of fresh values
        assert np.array_equal(s.get("Circle", "parent", np.uint32)[k:], sid[born])
        assert np.array_equal(x2[k:], sx[born]) and np.array_equal(z2[k:], sz[born])
        assert set(ids2[k:]) == set(range(next_id, next_id + b))
        assert np.array_equal(s.message_pbm("location"), pbm)
        next_id += b
        ids, x, y, z = ids2, x2, y2, z2
    s.close()


def test_boids_run_and_stay_in_bounds():
    # examples/cpp/boids_spatial3D shape (BASELINE config 0): 4096 boids, 100 steps; and the 2D variant
    n = 4096
    rng = np.random.default_rng(12)
    for model, dims in (("boids3d", 3), ("boids2d", 2)):
        pop = {k: rng.uniform(-0.5, 0.5, n).astype(np.float32) for k in ("x", "y", "z")[:dims]}
        v = rng.uniform(-1, 1, (dims, n)).astype(np.float32)
        v = v / np.linalg.norm(v, axis=0) * rng.uniform(0.1, 1.0, n).astype(np.float32)
        for a, k in enumerate(("fx", "fy", "fz")[:dims]):
            pop[k] = v[a].astype(np.float32)
        s = _sim(model)
        s.set_population("Boid", pop)
        s.step(100)
        assert s.count("Boid") == n
        for k in ("x", "y", "z")[:dims]:
            p = s.get("Boid", k, np.float32)
            assert np.all(np.isfinite(p)) and p.min() >= -0.5 and p.max() <= 0.5
        sp = np.sqrt(sum(s.get("Boid", k, np.float32).astype(np.float64) ** 2 for k in ("fx", "fy", "fz")[:dims]))
        assert sp.max() <= 1.0 + 0.05 * np.sqrt(dims) + 1e-4
        assert sorted(s.get("Boid", "_id", np.uint32)) == list(range(1, n + 1))
        s.close()


def test_radius_filtered_iterator_other_models():
    # the opt-in iterator on models with births / deaths (stress) and on Boids 3D / 2D: every in-radius message is
    # presented once in the same relative order, so the results are bit-identical to the reference-order iterator
    n, L = 40000, 34.0
    pos = _circles_pop(n, L, seed=7)
    res = []
    for m in (0, 1):
        s = _sim("stress", env_max=L, radius=2.0, death_mod=10, birth_mod=20, iter_mode=m)
        s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
        s.step(1)  # one step: newborn ids are atomic-order dependent, so later decisions (hashes of ids) differ run to run
        res.append([s.get("Circle", v, np.uint32) for v in ("neighbours", "parent")] + [s.get("Circle", "x", np.float32)] +
                   [np.sort(s.get("Circle", "_id", np.uint32))])  # newborn ids are handed out in atomic order: compare as a set
        s.close()
    assert len(res[0][0]) == len(res[1][0]) and len(res[0][0]) != n
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    rng = np.random.default_rng(3)
    nb = 20000
    for model, dims in (("boids3d", 3), ("boids2d", 2)):
        pop = {k: rng.uniform(-0.5, 0.5, nb).astype(np.float32) for k in ("x", "y", "z")[:dims]}
        for k in ("fx", "fy", "fz")[:dims]:
            pop[k] = rng.uniform(-0.5, 0.5, nb).astype(np.float32)
        res = []
        for m in (0, 1):
            s = _sim(model, iter_mode=m, stable=1)
            s.set_population("Boid", pop)
            s.step(5)
            res.append([s.get("Boid", k, np.float32) for k in ("x", "y", "fx", "fy")] + [s.get("Boid", "_id", np.uint32)])
            s.close()
        for a, b in zip(*res):
            assert np.array_equal(a, b), model


@pytest.mark.parametrize("which", [12, 13, 14])
def test_bucket_messaging_reference_tests(which):
    # BucketMessageTest.Mandatory / Optional / OptionalNone / Mandatory_Range (reference test_bucket.cu:99-316, 612-683)
    n = 1024
    ids = np.arange(n, dtype=np.int32)
    rng = np.random.default_rng(5)
    for variant in ((0, 1, 2) if which == 13 else (0,)):
        do_out = np.ones(n, np.int32)
        if which == 13:
            do_out = (rng.integers(0, 2, n) if variant == 0 else (np.zeros(n) if variant == 1 else np.ones(n))).astype(np.int32)
        s = _sim("test", which=which)
        s.set_population("agent", {"id": ids, "do_output": do_out})
        s.step(1)
        c1, c2, sm = (s.get("agent", v, np.uint32) for v in ("count1", "count2", "sum"))
        sent = do_out.astype(bool)
        bucket_count = np.bincount(ids[sent] // 2, minlength=n // 2)
        bucket_sum = np.bincount(ids[sent] // 2, weights=ids[sent], minlength=n // 2).astype(np.int64)
        m1 = np.where(ids == 0, 0, ids - 1) // 2
        if which == 14:
            m4 = (ids // 8) * 4
            exp_c = sum(bucket_count[m4 + j] for j in range(4))
            exp_s = sum(bucket_sum[m4 + j] for j in range(4))
            assert np.array_equal(c1, exp_c) and np.array_equal(sm, exp_s) and np.array_equal(c2, bucket_count[ids // 2])
        else:
            assert np.array_equal(c1, bucket_count[m1]) and np.array_equal(c2, bucket_count[m1]) and np.array_equal(sm, bucket_sum[m1])
        pbm_ref, _ = orc.bucket_build(12, 12 + n // 2, 12 + ids[sent] // 2)
        assert np.array_equal(s.message_pbm("bucket"), pbm_ref)
        s.close()


def test_host_agent_reductions():
    # HostAgentAPI::sum / min / max (what step functions such as Circles' drift validation call) run on the device
    n, L = 50000, 37.0
    pos = _circles_pop(n, L, seed=2)
    s = _sim("circles", env_max=L, radius=2.0)
    s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
    s.step(2)
    drift = s.get("Circle", "drift", np.float32)
    ids = s.get("Circle", "_id", np.uint32)
    assert abs(s.agent_reduce("Circle", "drift", "sum", "f") - float(drift.astype(np.float64).sum())) <= 1e-6 * float(drift.sum())
    assert s.agent_reduce("Circle", "drift", "min", "f") == float(drift.min())
    assert s.agent_reduce("Circle", "drift", "max", "f") == float(drift.max())
    assert s.agent_reduce("Circle", "_id", "sum", "u") == float(ids.astype(np.uint64).sum() % (1 << 32))  # sum<unsigned> wraps like T
    assert s.agent_reduce("Circle", "_id", "max", "u") == float(ids.max())
    # count(variable, value), mean and POPULATION standard deviation (reference HostAgentAPI.cuh:561-604, 700-718)
    keys = s.get("Circle", "_auto_sort_bin_index", np.uint32)
    v = int(np.bincount(keys).argmax())
    assert s.agent_reduce("Circle", "_auto_sort_bin_index", "count", "u", value=v) == float((keys == v).sum())
    assert s.agent_reduce("Circle", "_id", "count", "u", value=0) == 0.0
    x = s.get("Circle", "x", np.float32).astype(np.float64)
    assert abs(s.agent_reduce("Circle", "x", "mean", "f") - x.mean()) <= 1e-9 * L
    assert abs(s.agent_reduce("Circle", "x", "std", "f") - x.std()) <= 1e-9 * L
    s.close()


@pytest.mark.parametrize("graphs", [1, 0])
def test_step_function_reduction_recorded_behind_the_step(graphs):
    # The Circles example's Validation step function sums "drift" every step.  From the second step on that reduction is
    # recorded behind the step itself (inside its CUDA graph) and its result reaches the step function through mapped
    # host memory: the value must be the sum over the list exactly as that step left it, every step.
    import ctypes as C

    from flamegpu2_b200 import sim as fsim

    L = fsim.lib()
    L.fgbm_circles_validation.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    n, Lx = 60000, 39.0
    pos = _circles_pop(n, Lx, seed=5)
    s = _sim("circles", env_max=Lx, radius=2.0, graphs=graphs, validation=1)
    s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
    tot, d0, i0 = C.c_double(), C.c_uint(), C.c_uint()
    L.fgbm_circles_validation(C.byref(tot), C.byref(d0), C.byref(i0))
    seen = d0.value + i0.value
    for k in range(7):
        s.step(1)
        d, i = C.c_uint(), C.c_uint()
        L.fgbm_circles_validation(C.byref(tot), C.byref(d), C.byref(i))
        assert d.value + i.value == seen + k + 1, "the step function ran once per step"
        drift = s.get("Circle", "drift", np.float32)
        ref = float(drift.astype(np.float64).sum())
        assert abs(tot.value - ref) <= 2e-6 * ref, (k, tot.value, ref)  # the step function keeps the sum as float
    s.close()


def test_true3d_sort_key_extension():
    # b200 extension: the intended x,y,z sort key (the reference's key collapses z, CUDASimulation.cu:487)
    n, L = 30000, 31.0
    pos = _circles_pop(n, L, seed=33)
    g = orc.Grid(3, (0, 0, 0), (L, L, L), 2.0)
    s = _sim("circles", env_max=L, radius=2.0, true3d_sort=1)
    s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
    s.step(1)
    keys = g.sort_keys(*pos, true3d=True)
    perm = orc.sort_perm(keys, g.sort_max_bit(True))
    assert np.array_equal(s.get("Circle", "_id", np.uint32), perm + 1)
    assert np.array_equal(s.get("Circle", "_auto_sort_bin_index", np.uint32), keys[perm])
    s.close()


def test_bin_order_execution_does_not_change_results():
    # the execution-order permutation only changes which thread runs which agent
    n, L = 60000, 39.0
    pos = _circles_pop(n, L, seed=44)
    res = []
    for bin_order in (0, 1):
        s = _sim("circles", env_max=L, radius=2.0, stable=1, bin_order=bin_order)
        s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
        s.step(3)
        res.append([s.get("Circle", v, np.float32) for v in ("x", "y", "z", "drift")] + [s.get("Circle", "_id", np.uint32)])
        s.close()
    for a, b in zip(*res):
        assert np.array_equal(a, b)


def test_host_buffer_step_matches_population_path():
    # the bulk SoA exchange (bench.py e2e path) gives the same step as setPopulationData/getPopulationData
    n, L = 50000, 36.0
    pos = _circles_pop(n, L, seed=8)
    a = _sim("circles", env_max=L, radius=2.0, stable=1)
    a.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
    a.step(1)
    b = _sim("circles", env_max=L, radius=2.0, stable=1)
    out = {k: np.empty(n, np.float32) for k in ("x", "y", "z", "drift")}
    out["id"] = np.empty(n, np.uint32)
    for _ in range(2):  # twice: the second call re-uploads over a used simulation
        b.circles_step_host(pos[0], pos[1], pos[2], np.zeros(n, np.float32), 1, out)
        assert np.array_equal(out["id"], a.get("Circle", "_id", np.uint32))
        for v in ("x", "y", "z", "drift"):
            assert np.array_equal(out[v], a.get("Circle", v, np.float32)), v
    a.close()
    b.close()


@pytest.mark.parametrize("bin_order", [1, 0])
def test_iterator_radius_filtered_mode(bin_order, golden_dir):
    # b200 iterator mode 1: the messages within the radius are presented (lock-step walk, per-lane queue).
    # Circles filters by radius itself, so the step must be bit-identical (same in-radius messages in the same
    # relative order), with and without bin-order execution (lanes of a warp in different bins).
    mode = 1
    n, L = 60001, 39.0  # not a multiple of the warp size: partial last warp
    pos = _circles_pop(n, L, seed=91)
    res = []
    for m in (0, mode):
        s = _sim("circles", env_max=L, radius=2.0, stable=1, iter_mode=m, bin_order=bin_order)
        s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
        s.step(3)
        res.append([s.get("Circle", v, np.float32) for v in ("x", "y", "z", "drift")] + [s.get("Circle", "_id", np.uint32)])
        s.close()
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    # dense case: more accepted messages than one queue holds (forces intermediate drains)
    n2, L2 = 20000, 8.0
    pos2 = _circles_pop(n2, L2, seed=17)
    res = []
    for m in (0, mode):
        s = _sim("circles", env_max=L2, radius=2.0, stable=1, iter_mode=m, bin_order=bin_order)
        s.set_population("Circle", {"x": pos2[0], "y": pos2[1], "z": pos2[2]})
        s.step(2)
        res.append([s.get("Circle", v, np.float32) for v in ("x", "y", "z", "drift")])
        s.close()
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    p3 = np.fromfile(os.path.join(golden_dir, "mandatory3d_pos.f32"), dtype=np.float32).reshape(3, -1)
    expect = np.fromfile(os.path.join(golden_dir, "mandatory3d_expect.u32"), dtype=np.uint32)
    s = _sim("test", which=0, max_x=5, max_y=5, max_z=5, radius=1, sort_period=0, iter_mode=mode, bin_order=bin_order)
    s.set_population("agent", {"x": p3[0], "y": p3[1], "z": p3[2]})
    s.step(1)
    cnt = s.get("agent", "count", np.uint32)
    d = np.sqrt(((p3[:, :, None].astype(np.float64) - p3[:, None, :]) ** 2).sum(axis=0))
    # this model counts every message it is shown: all messages within the radius, plus the padding messages (at
    # infinity) a lane receives while other lanes of its warp still have accepted messages queued
    assert np.all(cnt >= (d < 0.9999).sum(axis=1)) and cnt.sum() < expect.sum()
    s.close()
    # 2D twin
    p2 = np.fromfile(os.path.join(golden_dir, "mandatory2d_pos.f32"), dtype=np.float32).reshape(2, -1)
    e2 = np.fromfile(os.path.join(golden_dir, "mandatory2d_expect.u32"), dtype=np.uint32)
    s = _sim("test", which=3, max_x=11, max_y=11, radius=1, sort_period=0, iter_mode=mode)
    s.set_population("agent", {"x": p2[0], "y": p2[1]})
    s.step(1)
    c2 = s.get("agent", "count", np.uint32)
    assert c2.sum() > 0 and c2.sum() < e2.sum()
    s.close()


def test_reference_function_condition_split():
    # TestAgentFunctionConditions.SplitAgents (test_agent_function_conditions.cu:48-105)
    n = 1000
    x = (np.arange(2 * n) % 2).astype(np.int32)
    y = np.tile(np.array([13, 14, 15, 16], np.int32), (2 * n, 1))
    s = _sim("test", which=10)
    s.set_population("agent", {"x": x, "y": y}, state="Start")
    s.step(1)
    assert s.count("agent", "Start") == 0 and s.count("agent", "End") == n and s.count("agent", "End2") == n
    assert np.all(s.get("agent", "x", np.int32, state="End") == 2)
    assert np.all(s.get("agent", "y", np.int32, 4, state="End") == [3, 4, 5, 6])
    assert np.all(s.get("agent", "x", np.int32, state="End2") == -1)
    assert np.all(s.get("agent", "y", np.int32, 4, state="End2") == [23, 24, 25, 26])
    # ids keep the original relative order inside each destination state
    ids = np.arange(1, 2 * n + 1, dtype=np.uint32)
    assert np.array_equal(s.get("agent", "_id", np.uint32, state="End"), ids[x == 1])
    assert np.array_equal(s.get("agent", "_id", np.uint32, state="End2"), ids[x != 1])
    s.close()


def test_function_condition_with_death_same_state():
    n = 5000
    rng = np.random.default_rng(6)
    x = rng.integers(0, 1000, n).astype(np.int32)
    s = _sim("test", which=11)
    s.set_population("agent", {"x": x})
    s.step(1)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    passed = x % 3 == 0
    keep = passed & (x % 2 != 0)
    exp_ids = np.concatenate([ids[~passed], ids[keep]])       # disabled agents first, then the executing survivors
    exp_x = np.concatenate([x[~passed], x[keep] + 1000])
    assert np.array_equal(s.get("agent", "_id", np.uint32), exp_ids)
    assert np.array_equal(s.get("agent", "x", np.int32), exp_x)
    # AllDisabled (:107-140): a second step where nothing passes must leave the list intact
    s.step(1)
    assert s.count("agent") > 0
    s.close()


def test_host_agent_histogram_and_custom_reductions():
    """HostAgentAPI::histogramEven / reduce / transformReduce with user functors (reference
    tests/test_cases/runtime/agent/host_reduction/test_histogram_even.cu, test_reduce.cu, test_transform_reduce.cu)"""
    n = 50000
    rng = np.random.default_rng(4)
    x = rng.integers(-100, 100, n).astype(np.int32)
    s = _sim("test", which=11)  # agent with one int variable "x"
    s.set_population("agent", {"x": x})
    h = s.agent_histogram("agent", "x", "i", 10, -100, 100)
    assert np.array_equal(h, np.bincount((x.astype(np.int64) + 100) * 10 // 200, minlength=10))
    assert s.agent_custom_reduce("agent", "x", 0, "i") == float(x.sum())
    assert s.agent_custom_reduce("agent", "x", 1, "i") == float(x.max())
    assert s.agent_custom_reduce("agent", "x", 2, "i") == float((x <= 0).sum())
    s.close()
