"""Parity against the reference's own CUDA build (oracle/_ref/ref_sim) at BASELINE.json sizes: one step from a common
seeded state (teacher-forced: both sides start every compared step from the SAME state), so every integer quantity
must be bit-exact -- PBM, per-bin message multisets, agent order, ids, counts -- and floats hold within the tolerance of
tests/test_ref_parity_gpu.py (summation order over the neighbourhood; CUDA vs libm sinf <= 2 ulp)."""
import numpy as np
import pytest

import fgbs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not fgbs.have_ref(), reason="oracle/_ref/ref_sim not built")]

RTOL, ATOL = 1e-5, 2e-6
F32 = {k: np.float32 for k in ("x", "y", "z", "drift", "fx", "fy", "fz")}


def _sim(model, **kw):
    from flamegpu2_b200 import sim as fsim

    return fsim.Simulation(model, **kw)


def _bins_multiset_equal(pbm, a, b):
    bins = np.repeat(np.arange(len(pbm) - 1), np.diff(pbm.astype(np.int64)))
    return np.array_equal(a[np.lexsort((a, bins))], b[np.lexsort((b, bins))])


def test_circles_1m_vs_reference(tmp_path):
    """BASELINE configs[1]: Circles-3D, 1 M agents, [0,100)^3, radius 2 (50^3 bins)."""
    n, L = 1_000_000, 100.0
    rng = np.random.default_rng(0)
    pos = [rng.uniform(0, L, n).astype(np.float32) for _ in range(3)]
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, {"x": pos[0], "y": pos[1], "z": pos[2]})
    fgbs.run_ref("circles", {"env_max": L, "radius": 2.0}, inp, str(tmp_path / "ref"), steps=1, dump_messages="location")
    ref = fgbs.read_state(str(tmp_path / "ref.Circle.bin"), F32)
    ref_pbm = fgbs.read_state(str(tmp_path / "ref.pbm.location.bin"))["_pbm"]
    ref_msg = fgbs.read_state(str(tmp_path / "ref.msg.location.bin"), F32)
    for mode in (-1, 0):  # the default (radius-filtered `move`, as the model declares) and the strict reference order
        s = _sim("circles", env_max=L, radius=2.0, iter_mode=mode)
        s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
        s.step(1)
        assert np.array_equal(s.get("Circle", "_id", np.uint32), ref["_id"]), "agent order"
        assert np.array_equal(s.get("Circle", "_auto_sort_bin_index", np.uint32), ref["_auto_sort_bin_index"])
        pbm = s.message_pbm("location")
        assert np.array_equal(pbm, ref_pbm), "PBM"
        assert _bins_multiset_equal(pbm, s.message_variable("location", "id", np.uint32, n), ref_msg["id"])
        for v in ("x", "y", "z"):
            assert np.allclose(s.get("Circle", v, np.float32), ref[v], rtol=RTOL, atol=ATOL), (v, mode)
        assert np.allclose(s.get("Circle", "drift", np.float32), ref["drift"], rtol=1e-3, atol=ATOL)
        s.close()


def test_boids3d_4096_100_steps_teacher_forced(tmp_path):
    """BASELINE configs[0]: boids_spatial3D, 4096 agents, 100 steps.  The reference runs 100 free steps and dumps the
    state after each; our step k starts from the reference's state k-1 and must reproduce state k: ids and list order
    bit-exact at EVERY step, floats within tolerance."""
    n, steps = 4096, 100
    rng = np.random.default_rng(12)
    pop = {k: rng.uniform(-0.5, 0.5, n).astype(np.float32) for k in ("x", "y", "z")}
    v = rng.uniform(-1, 1, (3, n)).astype(np.float32)
    v = (v / np.linalg.norm(v, axis=0) * rng.uniform(0.1, 1.0, n)).astype(np.float32)
    pop.update({"fx": v[0].copy(), "fy": v[1].copy(), "fz": v[2].copy()})
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, pop)
    fgbs.run_ref("boids3d", {}, inp, str(tmp_path / "ref"), steps=steps, dump_steps=True)
    s = _sim("boids3d")
    prev = dict(pop)
    prev["_id"] = np.arange(1, n + 1, dtype=np.uint32)
    worst = 0.0
    for k in range(1, steps + 1):
        ref = fgbs.read_state(str(tmp_path / f"ref.s{k}.Boid.bin"), F32)
        s.set_population("Boid", {c: prev[c] for c in ("x", "y", "z", "fx", "fy", "fz", "_id")})
        s.step(1)
        assert np.array_equal(s.get("Boid", "_id", np.uint32), ref["_id"]), f"agent order at step {k}"
        for c in ("x", "y", "z", "fx", "fy", "fz"):
            mine = s.get("Boid", c, np.float32)
            assert np.allclose(mine, ref[c], rtol=1e-4, atol=1e-5), (c, k)
            worst = max(worst, float(np.max(np.abs(mine - ref[c]))))
        prev = ref
    s.close()
    assert worst < 1e-4


def test_stress_4m_vs_reference(tmp_path):
    """BASELINE configs[3] at 4 M agents: 10 % deaths, 5 % births, Spatial3D messages; all results are integers or copies."""
    n = 4_000_000
    L = float(np.floor(np.cbrt(n)))
    rng = np.random.default_rng(77)
    pos = [rng.uniform(0, L, n).astype(np.float32) for _ in range(3)]
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, {"x": pos[0], "y": pos[1], "z": pos[2]})
    params = {"env_max": L, "radius": 2.0, "death_mod": 10, "birth_mod": 20}
    fgbs.run_ref("stress", params, inp, str(tmp_path / "ref"), dump_messages="location")
    ref = fgbs.read_state(str(tmp_path / "ref.Circle.bin"), F32)
    s = _sim("stress", **params)
    s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
    s.step(1)
    assert s.count("Circle") == len(ref["x"])
    for v in ("neighbours", "parent"):
        assert np.array_equal(s.get("Circle", v, np.uint32), ref[v]), v
    for v in ("x", "y", "z"):
        assert np.array_equal(s.get("Circle", v, np.float32), ref[v]), v
    ours, theirs = s.get("Circle", "_id", np.uint32), ref["_id"]
    k = int((ref["parent"] == 0).sum())
    assert np.array_equal(ours[:k], theirs[:k]), "survivor order"
    assert np.array_equal(np.sort(ours[k:]), np.sort(theirs[k:])), "newborn ids as a set"
    assert np.array_equal(s.message_pbm("location"), fgbs.read_state(str(tmp_path / "ref.pbm.location.bin"))["_pbm"])
    s.close()


def test_boids2d_4m_vs_reference(tmp_path):
    """BASELINE configs[2] at 4 M boids: the 2D PBM and the 3-strip iterator."""
    n = 4_000_000
    r = 0.0007
    rng = np.random.default_rng(21)
    pop = {k: rng.uniform(-0.5, 0.5, n).astype(np.float32) for k in ("x", "y")}
    v = rng.uniform(-1, 1, (2, n)).astype(np.float32)
    v = (v / np.linalg.norm(v, axis=0) * rng.uniform(0.1, 1.0, n)).astype(np.float32)
    pop.update({"fx": v[0].copy(), "fy": v[1].copy()})
    params = {"interaction_radius": r * 2, "separation_radius": r / 2}
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, pop)
    fgbs.run_ref("boids2d", params, inp, str(tmp_path / "ref"), steps=1, dump_messages="location")
    ref = fgbs.read_state(str(tmp_path / "ref.Boid.bin"), F32)
    ref_pbm = fgbs.read_state(str(tmp_path / "ref.pbm.location.bin"))["_pbm"]
    ref_msg = fgbs.read_state(str(tmp_path / "ref.msg.location.bin"), F32)
    s = _sim("boids2d", **params)
    s.set_population("Boid", pop)
    s.step(1)
    assert np.array_equal(s.get("Boid", "_id", np.uint32), ref["_id"])
    pbm = s.message_pbm("location")
    assert np.array_equal(pbm, ref_pbm)
    assert _bins_multiset_equal(pbm, s.message_variable("location", "id", np.uint32, n), ref_msg["id"])
    for k in ("x", "y", "fx", "fy"):
        assert np.allclose(s.get("Boid", k, np.float32), ref[k], rtol=1e-4, atol=1e-5), k
    s.close()
