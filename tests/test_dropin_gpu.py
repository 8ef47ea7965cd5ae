"""The drop-in of INTEGRATION.md section A, compiled and run: oracle/_ref/ref_sim_fgb is the reference library with
MessageSpatial3D/2D/Bucket::CUDAModelHandler::buildIndex and CUDAScatter::scatter replaced by
integration/fgb_reference_shims.cu (calls into libflamegpu2_b200.so through the C ABI); everything else -- model
description, CUDASimulation::step(), cuRVE, the reference's own device iterators and agent functions -- is the
reference's.  It must reproduce the unmodified reference (oracle/_ref/ref_sim) on the same seeded inputs: PBM and
survivor order bit-exact, bins equal as multisets, float state within the summation-order tolerance."""
import numpy as np
import pytest

import fgbs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not fgbs.have_dropin(), reason="oracle/_ref/ref_sim_fgb not built")]

RTOL, ATOL = 1e-5, 2e-6
F32 = {k: np.float32 for k in ("x", "y", "z", "drift", "fx", "fy", "fz", "untouched")}


def _both(tmp_path, model, params, inp, **kw):
    a = fgbs.run_ref(model, params, inp, str(tmp_path / "ref"), **kw)
    b = fgbs.run_ref(model, params, inp, str(tmp_path / "fgb"), binary=fgbs.REF_SIM_FGB, **kw)
    return a, b


def _bins_multiset_equal(pbm, a, b):
    bins = np.repeat(np.arange(len(pbm) - 1), np.diff(pbm.astype(np.int64)))
    return np.array_equal(a[np.lexsort((a, bins))], b[np.lexsort((b, bins))])


@pytest.mark.parametrize("n,L,steps", [(16384, 25.0, 1), (200000, 58.0, 1), (200000, 58.0, 5)])
def test_dropin_circles(tmp_path, n, L, steps):
    rng = np.random.default_rng(n + steps)
    pos = [rng.uniform(0, L, n).astype(np.float32) for _ in range(3)]
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, {"x": pos[0], "y": pos[1], "z": pos[2]})
    _both(tmp_path, "circles", {"env_max": L, "radius": 2.0}, inp, steps=steps, dump_messages="location")
    ref = fgbs.read_state(str(tmp_path / "ref.Circle.bin"), F32)
    got = fgbs.read_state(str(tmp_path / "fgb.Circle.bin"), F32)
    if steps == 1:
        # identical inputs: the PBM and the (stable) agent order are bit-exact
        assert np.array_equal(got["_id"], ref["_id"])
        pbm = fgbs.read_state(str(tmp_path / "fgb.pbm.location.bin"))["_pbm"]
        assert np.array_equal(pbm, fgbs.read_state(str(tmp_path / "ref.pbm.location.bin"))["_pbm"])
        assert pbm[-1] == n
        assert _bins_multiset_equal(pbm, fgbs.read_state(str(tmp_path / "fgb.msg.location.bin"))["id"],
                                    fgbs.read_state(str(tmp_path / "ref.msg.location.bin"))["id"])
        for v in ("x", "y", "z"):
            assert np.allclose(got[v], ref[v], rtol=RTOL, atol=ATOL), v
    else:
        # free-running: a 1-ulp difference (in-bin order is atomic-arrival order in both builds) may move an agent
        # across a bin edge, so compare per agent id with the free-running tolerance
        oa, ob = np.argsort(got["_id"]), np.argsort(ref["_id"])
        assert np.array_equal(got["_id"][oa], ref["_id"][ob])
        for v in ("x", "y", "z"):
            assert np.allclose(got[v][oa], ref[v][ob], rtol=1e-4, atol=1e-4), v


def test_dropin_death_birth_condition(tmp_path):
    # CUDAScatter::scatter -> fgb_compact: death (survivor order, array payload), births appended behind the
    # survivors, function condition ([disabled | executing] order)
    n = 4096
    rng = np.random.default_rng(3)
    x = rng.integers(0, 13, n).astype(np.uint32)
    arr = np.stack([x + 1, x + 2, x + 3], axis=1).astype(np.uint32)
    inp = str(tmp_path / "d.bin")
    fgbs.write_state(inp, {"x": x, "arr": arr})
    _both(tmp_path, "test", {"which": 5}, inp)
    ref, got = fgbs.read_state(str(tmp_path / "ref.agent.bin")), fgbs.read_state(str(tmp_path / "fgb.agent.bin"))
    assert len(got["x"]) == len(ref["x"]) < n
    for v in ("x", "_id", "arr"):
        assert np.array_equal(got[v], ref[v]), v
    ids0 = np.arange(n, dtype=np.uint32)
    inp = str(tmp_path / "b.bin")
    fgbs.write_state(inp, {"x": (ids0 + 1.0).astype(np.float32), "id": ids0})
    for which in (6, 7, 8):
        _both(tmp_path, "test", {"which": which}, inp)
        ref, got = fgbs.read_state(str(tmp_path / "ref.agent.bin"), F32), fgbs.read_state(str(tmp_path / "fgb.agent.bin"), F32)
        for v in ("id", "x", "untouched"):
            assert np.array_equal(got[v], ref[v]), (which, v)
        assert np.array_equal(np.sort(got["_id"]), np.sort(ref["_id"])), which
    xi = rng.integers(0, 1000, 5000).astype(np.int32)
    inp = str(tmp_path / "c.bin")
    fgbs.write_state(inp, {"x": xi})
    _both(tmp_path, "test", {"which": 11}, inp)
    ref, got = fgbs.read_state(str(tmp_path / "ref.agent.bin")), fgbs.read_state(str(tmp_path / "fgb.agent.bin"))
    assert np.array_equal(got["_id"], ref["_id"]) and np.array_equal(got["x"], ref["x"])


def test_dropin_stress_and_boids2d_and_bucket(tmp_path):
    # spatial 3D messages + death + birth in one step
    n, L = 60000, 39.0
    rng = np.random.default_rng(77)
    pos = [rng.uniform(0, L, n).astype(np.float32) for _ in range(3)]
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, {"x": pos[0], "y": pos[1], "z": pos[2]})
    _both(tmp_path, "stress", {"env_max": L, "radius": 2.0, "death_mod": 10, "birth_mod": 20}, inp, dump_messages="location")
    ref, got = fgbs.read_state(str(tmp_path / "ref.Circle.bin"), F32), fgbs.read_state(str(tmp_path / "fgb.Circle.bin"), F32)
    for v in ("neighbours", "parent", "x", "y", "z"):
        assert np.array_equal(got[v], ref[v]), v
    k = int((ref["parent"] == 0).sum())
    assert np.array_equal(got["_id"][:k], ref["_id"][:k]) and set(got["_id"][k:]) == set(ref["_id"][k:])
    assert np.array_equal(fgbs.read_state(str(tmp_path / "fgb.pbm.location.bin"))["_pbm"],
                          fgbs.read_state(str(tmp_path / "ref.pbm.location.bin"))["_pbm"])
    # 2D lists
    n = 20000
    pop = {k: rng.uniform(-0.5, 0.5, n).astype(np.float32) for k in ("x", "y")}
    v = rng.uniform(-1, 1, (2, n)).astype(np.float32)
    v = (v / np.linalg.norm(v, axis=0) * rng.uniform(0.1, 1.0, n)).astype(np.float32)
    pop.update({"fx": v[0].copy(), "fy": v[1].copy()})
    inp = str(tmp_path / "in2.bin")
    fgbs.write_state(inp, pop)
    _both(tmp_path, "boids2d", {"interaction_radius": 0.02, "separation_radius": 0.004}, inp, dump_messages="location")
    ref, got = fgbs.read_state(str(tmp_path / "ref.Boid.bin"), F32), fgbs.read_state(str(tmp_path / "fgb.Boid.bin"), F32)
    assert np.array_equal(got["_id"], ref["_id"])
    assert np.array_equal(fgbs.read_state(str(tmp_path / "fgb.pbm.location.bin"))["_pbm"],
                          fgbs.read_state(str(tmp_path / "ref.pbm.location.bin"))["_pbm"])
    for k in ("x", "y", "fx", "fy"):
        assert np.allclose(got[k], ref[k], rtol=1e-4, atol=1e-5), k
    # bucket lists (keys arrive unordered)
    n = 4096
    ids = rng.permutation(n).astype(np.int32)
    inp = str(tmp_path / "in3.bin")
    fgbs.write_state(inp, {"id": ids, "do_output": np.ones(n, np.int32)})
    _both(tmp_path, "test", {"which": 12, "bucket_upper": 12 + n // 2}, inp, dump_messages="bucket")
    ref, got = fgbs.read_state(str(tmp_path / "ref.agent.bin")), fgbs.read_state(str(tmp_path / "fgb.agent.bin"))
    for v in ("id", "count1", "count2", "sum"):
        assert np.array_equal(got[v], ref[v]), v
    assert np.array_equal(fgbs.read_state(str(tmp_path / "fgb.pbm.bucket.bin"))["_pbm"],
                          fgbs.read_state(str(tmp_path / "ref.pbm.bucket.bin"))["_pbm"])


def test_dropin_timing_line(tmp_path, capsys):
    # not a bench value: the same 1 M-agent Circles run through both binaries, per-step times from the reference's
    # own getElapsedTimeSteps(); printed for profiles/ (the shimmed build must not be slower than the reference)
    n, L = 1000000, 100.0
    rng = np.random.default_rng(5)
    pos = [rng.uniform(0, L, n).astype(np.float32) for _ in range(3)]
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, {"x": pos[0], "y": pos[1], "z": pos[2]})
    a, b = _both(tmp_path, "circles", {"env_max": L, "radius": 2.0}, inp, steps=25, warmup=5)
    ta, tb = 1e3 * float(np.mean(a["step_seconds"])), 1e3 * float(np.mean(b["step_seconds"]))
    with capsys.disabled():
        print(f"\n[dropin] Circles 1M ms/step: reference {ta:.3f}, reference + libflamegpu2_b200 shims {tb:.3f}")
    assert tb <= 1.05 * ta
