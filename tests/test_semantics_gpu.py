"""Restatements of the reference's own tests for list lifecycle and scheduling semantics, run on this repo's CUDA
path and, where oracle/_ref/ref_sim is present, on the reference's own CUDA build from the same inputs:
  * message append vs truncate        tests/test_cases/runtime/messaging/test_append_truncate.cu:106-309
  * device agent creation, 20 combos  tests/test_cases/runtime/agent/test_device_agent_creation.cu:53-1107
  * unique ids across states          tests/test_cases/runtime/agent/test_device_agent_creation.cu:1196-1296
  * state transitions                 tests/test_cases/runtime/agent/detail/test_agent_state_transition.cu:67-257
  * functions sharing a layer         tests/test_cases/simulation/test_cuda_simulation_concurrency.cu:779-853
"""
import itertools

import numpy as np
import pytest

import fgbs

pytestmark = pytest.mark.gpu

F32 = {k: np.float32 for k in ("x", "y", "z", "untouched")}


def _sim(model, **kw):
    from flamegpu2_b200 import sim as fsim

    return fsim.Simulation(model, **kw)


# ---- append / truncate ---------------------------------------------------------------------------------
@pytest.mark.parametrize("graphs", [1, 0])
def test_append_keep_data(graphs):
    """Append_KeepData: two functions of one step write the same list; the second appends."""
    n = 1024
    s = _sim("test", which=16, graphs=graphs)
    s.set_population("agent", {"count0": np.zeros(n, np.uint32)})
    for _ in range(2):  # second step: the list was truncated at the step boundary, not accumulated
        s.step(1)
        assert s.message_count("msg") == 0  # non-persistent lists are emptied at the end of the step
        assert np.all(s.get("agent", "count0", np.uint32) == n)
        assert np.all(s.get("agent", "count1", np.uint32) == n)
        assert np.all(s.get("agent", "count2", np.uint32) == 0)
    s.close()


def test_optional_append_keep_data():
    """OptionalAppend_KeepData: ~70 % write x=0 in layer 1 (and flip do_out), the others write x=1 in layer 2."""
    n = 1024
    rng = np.random.default_rng(11)
    do_out = (rng.random(n) < 0.7).astype(np.uint32)
    k = int(do_out.sum())
    s = _sim("test", which=16, append_optional=1)
    s.set_population("agent", {"do_out": do_out})
    s.step(1)
    assert np.all(s.get("agent", "count0", np.uint32) == k)
    assert np.all(s.get("agent", "count1", np.uint32) == n - k)
    s.step(1)
    assert np.all(s.get("agent", "count0", np.uint32) == n - k)
    assert np.all(s.get("agent", "count1", np.uint32) == k)
    s.close()


def test_append_keep_data_resize(tmp_path):
    """Append_KeepData_Resize (bug #725): the appending population is twice the first, forcing the list to grow."""
    n = 1024
    s = _sim("test", which=17)
    s.set_population("agent", {"_n": np.zeros(n, np.uint32)})  # no user variables: "_n" only carries the size
    s.set_population("b", {"_n": np.zeros(2 * n, np.uint32)})
    s.set_population("c", {"count0": np.zeros(1, np.uint32)})
    s.step(1)
    assert s.get("c", "count0", np.uint32)[0] == 0
    assert s.get("c", "count1", np.uint32)[0] == n
    assert s.get("c", "count2", np.uint32)[0] == 2 * n
    s.close()
    if fgbs.have_ref():
        for a, m in (("agent", n), ("b", 2 * n), ("c", 1)):
            fgbs.write_state(str(tmp_path / f"{a}.bin"), {"_n": np.zeros(m, np.uint32)})
        fgbs.run_ref("test", {"which": 17}, None, str(tmp_path / "ref"),
                     pops=[(a, "", str(tmp_path / f"{a}.bin")) for a in ("agent", "b", "c")], dumps=[("c", "")])
        ref = fgbs.read_state(str(tmp_path / "ref.c.bin"))
        assert (ref["count0"][0], ref["count1"][0], ref["count2"][0]) == (0, n, 2 * n)


# ---- device agent creation: the reference's 20 combinations ---------------------------------------------
COMBOS = [(o, d, t, c) for o, d, t, c in itertools.product((0, 1), (0, 1), (0, 1, 2), (0, 1)) if not (c and t == 2)]


def _birth_states(opt, death, target, cond):
    """(state the population starts in, states to read back)"""
    if target == 2:
        return "", [("agent", ""), ("agent2", "")]
    if cond and not death:
        return "a", [("agent", "a"), ("agent", "b")] + ([("agent", "c")] if target == 1 else [])
    if target == 1:
        return "a", [("agent", "a"), ("agent", "b")]
    return "", [("agent", "")]


@pytest.mark.parametrize("opt,death,target,cond", COMBOS)
def test_device_agent_creation_combo(tmp_path, opt, death, target, cond):
    n = 1024
    ids0 = np.arange(n, dtype=np.uint32)
    x0 = (ids0 + 1.0).astype(np.float32)
    params = {"which": 15, "birth_optional": opt, "birth_death": death, "birth_target": target, "birth_condition": cond}
    start, reads = _birth_states(opt, death, target, cond)
    s = _sim("test", **params)
    s.set_population("agent", {"x": x0, "id": ids0}, state=start or None)
    s.step(1)
    ours = {}
    for a, st in reads:
        ours[(a, st)] = {v: s.get(a, v, np.float32 if v in F32 else np.uint32, state=st or None) for v in ("id", "x", "untouched", "_id")}
    s.close()

    # the reference test's own expectations (counts; parents keep x - id == 1, children have x - id == 12, defaults 15)
    executing = n // 2 if cond else n
    born = executing // 2 if opt else executing
    # which parents die: mandatory-with-death kills every executing parent, optional-with-death the non-birthing half
    dead = (executing if not opt else executing - born) if death else 0
    total = sum(len(v["id"]) for v in ours.values())
    assert total == n - dead + born, (total, n, dead, born)
    children = np.concatenate([v["x"] - v["id"] == 12.0 for v in ours.values()])
    parents = np.concatenate([v["x"] - v["id"] == 1.0 for v in ours.values()])
    assert children.sum() == born and parents.sum() == n - dead
    assert all(np.all(v["untouched"] == 15.0) for v in ours.values())
    for agent_type in {a for a, _ in reads}:  # ids are unique per agent type, over all of its states
        ids = np.concatenate([v["_id"] for (a, _), v in ours.items() if a == agent_type])
        assert len(np.unique(ids)) == len(ids) and np.all(ids != 0), "ids must be unique and set"

    if not fgbs.have_ref():
        return
    inp = str(tmp_path / "in.bin")
    fgbs.write_state(inp, {"x": x0, "id": ids0})
    fgbs.run_ref("test", params, None, str(tmp_path / "ref"), pops=[("agent", start, inp)], dumps=reads)
    for a, st in reads:
        ref = fgbs.read_state(str(tmp_path / f"ref.{a}{'.' + st if st else ''}.bin"), F32)
        mine = ours[(a, st)]
        assert len(mine["id"]) == len(ref["id"]), (a, st)
        # list order (survivors in order, then children in parent order) and every value bit-exact; ids as a set
        for v in ("id", "x", "untouched"):
            assert np.array_equal(mine[v], ref[v]), (a, st, v)
        assert np.array_equal(np.sort(mine["_id"]), np.sort(ref["_id"])), (a, st)


def test_unique_ids_across_states(tmp_path):
    """AgentID_MultipleStatesUniqueIDs; its last layer holds two functions (states a and b) that run concurrently."""
    n = 100
    s = _sim("test", which=20)
    # the reference test uploads the same AgentVector to both states: ids are assigned on first use
    s.set_population("agent", {"id_copy": np.zeros(n, np.uint32)}, state="a")
    s.set_population("agent", {"id_copy": np.zeros(n, np.uint32)}, state="b")
    s.step(1)
    pops = {st: {v: s.get("agent", v, np.uint32, state=st) for v in ("_id", "id_copy", "id_other")} for st in ("a", "b")}
    s.close()
    ids = np.concatenate([pops[st]["_id"] for st in ("a", "b")])
    assert len(ids) == 4 * n and len(np.unique(ids)) == 4 * n and np.all(ids != 0)
    for st in ("a", "b"):
        assert np.array_equal(pops[st]["_id"], pops[st]["id_copy"])  # copy_id / copy_id2 both ran
    pair = {}
    for st in ("a", "b"):
        for i, o in zip(pops[st]["_id"], pops[st]["id_other"]):
            assert o != 0
            pair[int(i)] = int(o)
    assert all(pair[o] == i for i, o in pair.items()), "parent/child id pairings must be mutual"


# ---- state transitions -----------------------------------------------------------------------------------
def test_state_transition_chain(tmp_path):
    """Src_0_Dest_10 and Src_10_Dest_0: Start -> End in step 1, End -> End2 in step 2 (layer order End->End2 first)."""
    n = 10
    y0 = np.tile(np.array([13, 14, 15, 16], np.int32), (n, 1))
    s = _sim("test", which=18)
    s.set_population("agent", {"x": np.full(n, 12, np.int32), "y": y0}, state="Start")
    s.step(1)
    assert (s.count("agent", "Start"), s.count("agent", "End"), s.count("agent", "End2")) == (0, n, 0)
    assert np.all(s.get("agent", "x", np.int32, state="End") == 11)
    assert np.array_equal(s.get("agent", "y", np.int32, 4, state="End"), np.tile(np.array([23, 24, 25, 26], np.int32), (n, 1)))
    s.step(1)
    assert (s.count("agent", "Start"), s.count("agent", "End"), s.count("agent", "End2")) == (0, 0, n)
    assert np.all(s.get("agent", "x", np.int32, state="End2") == 13)
    assert np.array_equal(s.get("agent", "y", np.int32, 4, state="End2"), np.tile(np.array([3, 4, 5, 6], np.int32), (n, 1)))
    s.step(1)  # nothing left to move
    assert (s.count("agent", "Start"), s.count("agent", "End"), s.count("agent", "End2")) == (0, 0, n)
    s.close()


def test_state_transition_conditional(tmp_path):
    """Src_10_Dest_10: every round the agents whose counter reached zero move Start -> End; order must match the reference."""
    rounds, per = 3, 10
    n = rounds * per
    val = (1 + np.arange(n) % rounds).astype(np.uint32)
    z0 = np.tile(np.array([13, 14, 15, 16], np.int32), (n, 1))
    s = _sim("test", which=19)
    s.set_population("agent", {"x": val, "y": val.copy(), "z": z0}, state="Start")
    ours = []
    for i in range(1, rounds + 1):
        s.step(1)
        ys, ye = s.get("agent", "y", np.uint32, state="Start"), s.get("agent", "y", np.uint32, state="End")
        assert len(ys) == (rounds - i) * per and len(ye) == i * per
        assert np.all(ys > i) and np.all(ye <= i)
        assert np.all(s.get("agent", "z", np.int32, 4, state="Start") == np.array([23, 24, 25, 26], np.int32))
        assert np.all(s.get("agent", "z", np.int32, 4, state="End") == np.array([3, 4, 5, 6], np.int32))
        ours.append((s.get("agent", "_id", np.uint32, state="Start"), s.get("agent", "_id", np.uint32, state="End")))
    s.close()
    if fgbs.have_ref():
        inp = str(tmp_path / "in.bin")
        fgbs.write_state(inp, {"x": val, "y": val.copy(), "z": z0})
        for i in range(1, rounds + 1):
            fgbs.run_ref("test", {"which": 19}, None, str(tmp_path / f"ref{i}"), steps=i, pops=[("agent", "Start", inp)],
                         dumps=[("agent", "Start"), ("agent", "End")])
            a = fgbs.read_state(str(tmp_path / f"ref{i}.agent.Start.bin"))
            b = fgbs.read_state(str(tmp_path / f"ref{i}.agent.End.bin"))
            assert np.array_equal(ours[i - 1][0], a["_id"]) and np.array_equal(ours[i - 1][1], b["_id"]), i


@pytest.mark.parametrize("graphs", [1, 0])
def test_conditional_transitions_keep_bounds_bounded(graphs):
    """Two conditional transitions A <-> B over many steps: the host-side list bounds (hence capacities, launch grids
    and graph keys) must stay within the population; they used to grow by the source bound every step."""
    n = 3000
    y = np.arange(n, dtype=np.uint32)
    s = _sim("test", which=22, graphs=graphs)
    s.set_population("agent", {"x": np.zeros(n, np.uint32), "y": y}, state="A")
    for _ in range(40):
        s.step(1)
    a, b = s.count("agent", "A"), s.count("agent", "B")
    assert a + b == n
    ids = np.concatenate([s.get("agent", "_id", np.uint32, state="A"), s.get("agent", "_id", np.uint32, state="B")])
    assert len(np.unique(ids)) == n
    for st in ("A", "B"):
        bound, cap = s.list_bound("agent", st)
        assert bound <= n and cap <= 2 * n + 512, (st, bound, cap)
    if graphs:
        assert s.graphs <= 8, "bounds settle, so the step graphs are reused instead of re-captured every step"
    s.close()


# ---- several functions in one layer ---------------------------------------------------------------------------
@pytest.mark.parametrize("graphs", [1, 0])
def test_concurrent_spatial_layers(tmp_path, graphs):
    """ConcurrentMessageOutputInputSpatial3D: four agent types with their own Spatial3D lists; all four outputs share
    layer 0 and all four readers layer 1 (one stream per function; with graphs: parallel branches of the step graph)."""
    n, L = 3000, 9.0
    rng = np.random.default_rng(5)
    pops = {}
    params = {"which": 21, "max_x": L, "max_y": L, "max_z": L, "radius": 1, "sort_period": 1}
    s = _sim("test", graphs=graphs, **params)
    for i in range(4):
        pops[i] = [rng.uniform(0, L, n + 17 * i).astype(np.float32) for _ in range(3)]
        s.set_population(f"agent_{i}", {"x": pops[i][0], "y": pops[i][1], "z": pops[i][2]})
    s.step(2)
    ours = {i: {v: s.get(f"agent_{i}", v, np.uint32) for v in ("_id", "count", "badCount", "idsum")} for i in range(4)}
    if graphs:
        assert s.graph_width >= 4, "the four functions of a layer must be parallel branches of the captured graph"
    s.close()
    # serial execution of the same model gives the same integers
    t = _sim("test", graphs=0, concurrency=0, **params)
    for i in range(4):
        t.set_population(f"agent_{i}", {"x": pops[i][0], "y": pops[i][1], "z": pops[i][2]})
    t.step(2)
    for i in range(4):
        for v in ("_id", "count", "badCount", "idsum"):
            assert np.array_equal(ours[i][v], t.get(f"agent_{i}", v, np.uint32)), (i, v)
        assert np.all(ours[i]["badCount"] == 0)
    t.close()
    if fgbs.have_ref():
        files = []
        for i in range(4):
            path = str(tmp_path / f"a{i}.bin")
            fgbs.write_state(path, {"x": pops[i][0], "y": pops[i][1], "z": pops[i][2]})
            files.append((f"agent_{i}", "", path))
        fgbs.run_ref("test", params, None, str(tmp_path / "ref"), steps=2, pops=files, dumps=[(f"agent_{i}", "") for i in range(4)])
        for i in range(4):
            ref = fgbs.read_state(str(tmp_path / f"ref.agent_{i}.bin"))
            for v in ("_id", "count", "badCount", "idsum"):
                assert np.array_equal(ours[i][v], ref[v]), (i, v)


# ---- host agent creation (scatter_new_agents) -----------------------------------------------------
@pytest.mark.parametrize("host_init", [1, 0])
def test_host_agent_creation_from_init_and_step(host_init):
    """HostAgentCreationTest.FromInit / FromStep (tests/test_cases/runtime/agent/test_host_agent_creation.cu:56-130): 512 agents made
    by a host function join 512 uploaded ones; values set on the host arrive, untouched variables hold their defaults, ids are unique"""
    n = 512
    s = _sim("test", which=23, host_init=host_init)
    s.set_population("agent", {"x": np.full(n, 12.0, np.float32)})
    s.simulate(1)
    x = s.get("agent", "x", np.float32)
    assert len(x) == 2 * n and (x == 12.0).sum() == n and (x == 1.0).sum() == n
    assert np.array_equal(x[:n], np.full(n, 12.0, np.float32)), "host-made agents are appended behind the existing ones"
    assert np.all(s.get("agent", "default", np.float32) == 15.0)
    ids = s.get("agent", "_id", np.uint32)
    assert len(np.unique(ids)) == 2 * n and np.all(ids != 0)
    s.close()


def test_host_agent_creation_multi_agent_and_state():
    """HostAgentCreationTest.FromStepMultiAgent (test_host_agent_creation.cu:36-43): a step function creates agents of two
    types, one of them into a non-initial state, over two steps"""
    n = 300
    s = _sim("test", which=23, host_init=2)
    s.set_population("agent", {"x": np.full(n, 12.0, np.float32)}, state="a")
    s.simulate(2)
    assert s.count("agent", "a") == n
    xb = s.get("agent", "x", np.float32, state="b")
    y2 = s.get("agent2", "y", np.float32)
    assert len(xb) == 1024 and np.all(xb == 1.0) and len(y2) == 1024 and np.all(y2 == 2.0)
    ids = np.concatenate([s.get("agent", "_id", np.uint32, state="a"), s.get("agent", "_id", np.uint32, state="b")])
    assert len(np.unique(ids)) == n + 1024
    s.close()
