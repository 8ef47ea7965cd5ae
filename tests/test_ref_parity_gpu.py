"""Parity against the REFERENCE'S OWN CUDA IMPLEMENTATION (FLAME GPU 2 built unmodified from
/root/reference into oracle/_ref/ref_sim by oracle/ref_build/build_ref.sh), on the same seeded
inputs, one step from a common state:
  PBM bit-exact, per-bin message multisets equal, agent order + all integer state bit-exact,
  newborn ids equal as a set, float state within tolerance (summation order over ~33 neighbours).
The same runs also pin the CPU oracle to the reference (oracle == reference)."""
import os

import numpy as np
import pytest

import fgbs
import oracle_py as orc

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not fgbs.have_ref(), reason="oracle/_ref/ref_sim not built")]

RTOL, ATOL = 1e-5, 2e-6
F32 = {k: np.float32 for k in ("x", "y", "z", "drift", "fx", "fy", "fz", "result_x", "result_y", "result_z", "untouched")}


def _sim(model, **kw):
    from flamegpu2_b200 import sim as fsim

    return fsim.Simulation(model, **kw)


def _bins_multiset_equal(pbm, a, b):
    """a, b: per-message values in sorted-list order; equal as multisets inside every bin"""
    bins = np.repeat(np.arange(len(pbm) - 1), np.diff(pbm.astype(np.int64)))
    oa = np.lexsort((a, bins))
    ob = np.lexsort((b, bins))
    return np.array_equal(a[oa], b[ob])


@pytest.mark.parametrize("n,L", [(16384, 25.0), (200000, 58.0)])
def test_circles_step_vs_reference(tmp_path, n, L):
    rng = np.random.default_rng(n)
    pos = [rng.uniform(0, L, n).astype(np.float32) for _ in range(3)]
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, {"x": pos[0], "y": pos[1], "z": pos[2]})
    fgbs.run_ref("circles", {"env_max": L, "radius": 2.0}, inp, str(tmp_path / "ref"), steps=1, dump_messages="location")
    ref = fgbs.read_state(str(tmp_path / "ref.Circle.bin"), F32)
    ref_pbm = fgbs.read_state(str(tmp_path / "ref.pbm.location.bin"))["_pbm"]
    ref_msg = fgbs.read_state(str(tmp_path / "ref.msg.location.bin"), F32)

    s = _sim("circles", env_max=L, radius=2.0)
    s.set_population("Circle", {"x": pos[0], "y": pos[1], "z": pos[2]})
    s.step(1)
    ids = s.get("Circle", "_id", np.uint32)
    assert np.array_equal(ids, ref["_id"]), "agent order (stable auto-sort) must match the reference bit-exactly"
    assert np.array_equal(s.get("Circle", "_auto_sort_bin_index", np.uint32), ref["_auto_sort_bin_index"])
    pbm = s.message_pbm("location")
    assert np.array_equal(pbm, ref_pbm), "PBM must match the reference bit-exactly"
    mid = s.message_variable("location", "id", np.uint32, n)
    assert _bins_multiset_equal(pbm, mid, ref_msg["id"]), "bins must hold the same messages as the reference's"
    for v in ("x", "y", "z"):
        assert np.allclose(s.get("Circle", v, np.float32), ref[v], rtol=RTOL, atol=ATOL), v
    assert np.allclose(s.get("Circle", "drift", np.float32), ref["drift"], rtol=1e-3, atol=ATOL)
    s.close()

    # and the CPU oracle against the reference (this is what pins the oracle)
    g = orc.Grid(3, (0, 0, 0), (L, L, L), 2.0)
    i_o, x_o, y_o, z_o, d_o, pbm_o = g.circles_step(np.arange(1, n + 1, dtype=np.uint32), pos[0], pos[1], pos[2],
                                                    np.zeros(n, np.float32), want_pbm=True)
    assert np.array_equal(i_o, ref["_id"]) and np.array_equal(pbm_o, ref_pbm)
    assert np.allclose(x_o, ref["x"], rtol=RTOL, atol=ATOL) and np.allclose(z_o, ref["z"], rtol=RTOL, atol=ATOL)


def test_reference_mandatory3d_vs_reference(tmp_path, golden_dir):
    pos = np.fromfile(os.path.join(golden_dir, "mandatory3d_pos.f32"), dtype=np.float32).reshape(3, -1)
    expect = np.fromfile(os.path.join(golden_dir, "mandatory3d_expect.u32"), dtype=np.uint32)
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, {"x": pos[0], "y": pos[1], "z": pos[2]})
    params = {"which": 0, "max_x": 5, "max_y": 5, "max_z": 5, "radius": 1, "sort_period": 1}
    fgbs.run_ref("test", params, inp, str(tmp_path / "ref"), dump_messages="location")
    ref = fgbs.read_state(str(tmp_path / "ref.agent.bin"), F32)
    assert np.array_equal(ref["count"], expect[ref["_id"] - 1]), "the reference reproduces its own test's expectation"
    s = _sim("test", **params)
    s.set_population("agent", {"x": pos[0], "y": pos[1], "z": pos[2]})
    s.step(1)
    assert np.array_equal(s.get("agent", "_id", np.uint32), ref["_id"])
    for v in ("count", "badCount", "idsum"):
        assert np.array_equal(s.get("agent", v, np.uint32), ref[v]), v
    assert np.array_equal(s.message_pbm("location"), fgbs.read_state(str(tmp_path / "ref.pbm.location.bin"))["_pbm"])
    s.close()


def test_death_and_birth_vs_reference(tmp_path):
    n = 4096
    rng = np.random.default_rng(3)
    # death: survivor order and array payload bit-exact
    x = rng.integers(0, 13, n).astype(np.uint32)
    arr = np.stack([x + 1, x + 2, x + 3], axis=1).astype(np.uint32)
    inp = str(tmp_path / "d.bin")
    fgbs.write_state(inp, {"x": x, "arr": arr})
    fgbs.run_ref("test", {"which": 5}, inp, str(tmp_path / "refd"))
    ref = fgbs.read_state(str(tmp_path / "refd.agent.bin"))
    s = _sim("test", which=5)
    s.set_population("agent", {"x": x, "arr": arr})
    s.step(1)
    for v in ("x", "_id"):
        assert np.array_equal(s.get("agent", v, np.uint32), ref[v]), v
    assert np.array_equal(s.get("agent", "arr", np.uint32, 3), ref["arr"])
    s.close()
    # births (+ deaths): order of survivors and children, values, defaults; ids as a set
    ids0 = np.arange(n, dtype=np.uint32)
    inp = str(tmp_path / "b.bin")
    fgbs.write_state(inp, {"x": (ids0 + 1.0).astype(np.float32), "id": ids0})
    for which in (6, 7, 8):
        fgbs.run_ref("test", {"which": which}, inp, str(tmp_path / f"refb{which}"))
        ref = fgbs.read_state(str(tmp_path / f"refb{which}.agent.bin"), F32)
        s = _sim("test", which=which)
        s.set_population("agent", {"x": (ids0 + 1.0).astype(np.float32), "id": ids0})
        s.step(1)
        assert s.count("agent") == len(ref["id"]), which
        assert np.array_equal(s.get("agent", "id", np.uint32), ref["id"]), which
        assert np.array_equal(s.get("agent", "x", np.float32), ref["x"]), which
        assert np.array_equal(s.get("agent", "untouched", np.float32), ref["untouched"]), which
        assert np.array_equal(np.sort(s.get("agent", "_id", np.uint32)), np.sort(ref["_id"])), which
        s.close()


def test_function_conditions_vs_reference(tmp_path):
    # condition + death in one state: list order [disabled | executing survivors] must match the reference's
    n = 5000
    rng = np.random.default_rng(6)
    x = rng.integers(0, 1000, n).astype(np.int32)
    inp = str(tmp_path / "c.bin")
    fgbs.write_state(inp, {"x": x})
    fgbs.run_ref("test", {"which": 11}, inp, str(tmp_path / "refc"))
    ref = fgbs.read_state(str(tmp_path / "refc.agent.bin"))
    s = _sim("test", which=11)
    s.set_population("agent", {"x": x})
    s.step(1)
    assert np.array_equal(s.get("agent", "_id", np.uint32), ref["_id"])
    assert np.array_equal(s.get("agent", "x", np.int32).view(np.uint32), ref["x"])
    s.close()


def test_stress_step_vs_reference(tmp_path):
    n, L = 60000, 39.0
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
        assert np.array_equal(s.get("Circle", v, np.float32), ref[v]), v  # positions are only copied: bit-exact
    ours, theirs = s.get("Circle", "_id", np.uint32), ref["_id"]
    k = int((ref["parent"] == 0).sum())  # survivors have parent 0
    assert np.array_equal(ours[:k], theirs[:k]) and set(ours[k:]) == set(theirs[k:])
    assert np.array_equal(s.message_pbm("location"), fgbs.read_state(str(tmp_path / "ref.pbm.location.bin"))["_pbm"])
    s.close()


def test_boids3d_step_vs_reference(tmp_path):
    n = 4096
    rng = np.random.default_rng(12)
    pop = {k: rng.uniform(-0.5, 0.5, n).astype(np.float32) for k in ("x", "y", "z")}
    v = rng.uniform(-1, 1, (3, n)).astype(np.float32)
    v = (v / np.linalg.norm(v, axis=0) * rng.uniform(0.1, 1.0, n)).astype(np.float32)
    pop.update({"fx": v[0].copy(), "fy": v[1].copy(), "fz": v[2].copy()})
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, pop)
    fgbs.run_ref("boids3d", {}, inp, str(tmp_path / "ref"), dump_messages="location")
    ref = fgbs.read_state(str(tmp_path / "ref.Boid.bin"), F32)
    s = _sim("boids3d")
    s.set_population("Boid", pop)
    s.step(1)
    assert np.array_equal(s.get("Boid", "_id", np.uint32), ref["_id"])
    assert np.array_equal(s.message_pbm("location"), fgbs.read_state(str(tmp_path / "ref.pbm.location.bin"))["_pbm"])
    for k in ("x", "y", "z", "fx", "fy", "fz"):
        assert np.allclose(s.get("Boid", k, np.float32), ref[k], rtol=1e-4, atol=1e-5), k
    s.close()


def test_boids2d_steps_vs_reference(tmp_path):
    # the 2D PBM + 3-strip iterator path (BASELINE config 2 in miniature), one step (agent order -- the 2D sort
    # key has no collapsed axis -- and PBM bit-exact) and three free-running steps (per-agent state within tolerance)
    n = 20000
    rng = np.random.default_rng(21)
    pop = {k: rng.uniform(-0.5, 0.5, n).astype(np.float32) for k in ("x", "y")}
    v = rng.uniform(-1, 1, (2, n)).astype(np.float32)
    v = (v / np.linalg.norm(v, axis=0) * rng.uniform(0.1, 1.0, n)).astype(np.float32)
    pop.update({"fx": v[0].copy(), "fy": v[1].copy()})
    params = {"interaction_radius": 0.02, "separation_radius": 0.004}
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, pop)
    for steps in (1, 3):
        fgbs.run_ref("boids2d", params, inp, str(tmp_path / f"ref{steps}"), steps=steps, dump_messages="location")
        ref = fgbs.read_state(str(tmp_path / f"ref{steps}.Boid.bin"), F32)
        s = _sim("boids2d", **params)
        s.set_population("Boid", pop)
        s.step(steps)
        ids = s.get("Boid", "_id", np.uint32)
        if steps == 1:  # identical inputs: order and PBM are bit-exact; later a 1-ulp difference may move a boid across a bin edge
            assert np.array_equal(ids, ref["_id"])
            assert np.array_equal(s.message_pbm("location"), fgbs.read_state(str(tmp_path / "ref1.pbm.location.bin"))["_pbm"])
        oa, ob = np.argsort(ids), np.argsort(ref["_id"])
        assert np.array_equal(ids[oa], ref["_id"][ob])
        for k in ("x", "y", "fx", "fy"):
            assert np.allclose(s.get("Boid", k, np.float32)[oa], ref[k][ob], rtol=1e-4, atol=1e-5), (k, steps)
        s.close()


@pytest.mark.parametrize("which", [12, 13, 14])
def test_bucket_messaging_vs_reference(tmp_path, which):
    # bucket lists: PBM bit-exact, per-bucket message multisets equal, all (integer) agent results bit-exact
    n = 4096
    rng = np.random.default_rng(which)
    ids = rng.permutation(n).astype(np.int32)  # keys arrive unordered
    do_out = (rng.integers(0, 2, n) if which == 13 else np.ones(n)).astype(np.int32)
    upper = 12 + n // 2
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, {"id": ids, "do_output": do_out})
    fgbs.run_ref("test", {"which": which, "bucket_upper": upper}, inp, str(tmp_path / "ref"), steps=1, dump_messages="bucket")
    ref = fgbs.read_state(str(tmp_path / "ref.agent.bin"))
    ref_pbm = fgbs.read_state(str(tmp_path / "ref.pbm.bucket.bin"))["_pbm"]
    ref_msg = fgbs.read_state(str(tmp_path / "ref.msg.bucket.bin"))
    s = _sim("test", which=which, bucket_upper=upper)
    s.set_population("agent", {"id": ids, "do_output": do_out})
    s.step(1)
    for v in ("id", "count1", "count2", "sum"):
        assert np.array_equal(s.get("agent", v, np.uint32), ref[v]), v
    pbm = s.message_pbm("bucket")
    assert np.array_equal(pbm, ref_pbm), "PBM must match the reference bit-exactly"
    nm = int(pbm[-1])
    assert nm == len(ref_msg["id"]) == int(do_out.sum())
    mid = s.message_variable("bucket", "id", np.uint32, nm)
    assert _bins_multiset_equal(pbm, mid, ref_msg["id"])
    assert np.array_equal(s.message_variable("bucket", "_key", np.uint32, nm), ref_msg["_key"])
    # and the oracle agrees with both
    sent = do_out.astype(bool)
    pbm_o, _ = orc.bucket_build(12, upper, 12 + ids[sent] // 2)
    assert np.array_equal(pbm_o, ref_pbm)
    s.close()
