"""GPU parity tests proper: the CUDA path (through the C ABI, via BatchedRailEnv) against the C oracle
and the golden vectors of the unmodified reference, on identical worlds, actions and malfunction
schedules.  Integer fields (positions, directions, states, counters, rewards, dones, adjacency,
orders, valid actions, deadlocks, distance maps) must match bit-exactly; float features to
rtol 1e-6 (BASELINE.json north_star), and in practice bit-exactly as well.
"""
import numpy as np
import pytest

from conftest import golden_names

pytestmark = pytest.mark.gpu

STATE_KEYS = ["pos", "dir", "state", "ctr", "mal", "nmal", "saved", "arrival", "old_pos", "old_dir", "sig_mal"]
INT_OBS = ["adjacency", "node_order", "edge_order", "valid_actions"]
FLOAT_OBS = ["attr", "forest", "dist_target"]
RTOL = 1e-6          # the contract (BASELINE.json north_star): float features to rtol 1e-6
BIT_EXACT_FLOATS = True   # what the kernels actually deliver, and what these tests demand: the same float32 bits


PLANS = [
    {"segcap": 3},                            # segment pool overflows into its global spill space
    {"entcap": 16},                           # the prediction entries spill to global memory
    {"segcap": 0, "entcap": 0},
    {"sortsmall": 1},                         # every bucket with two or more entries takes the warp sort (bitonic network)
    {"sortsmall": 1, "entcap": 0},            # ... with the entries and the sort scratch sharing the spill space
    {"ctas": 1},                              # all static tables staged in shared memory (TMA bulk copies)
    {"tables": 0},                            # all static tables read from global memory
    {"nt": 64}, {"nt": 128}, {"nt": 256}, {"nt": 512}, {"nt": 1024},
    # fused kernel / split launch (index kernel + `parts` tree CTAs per environment), forced either way
    {"parts": 0}, {"parts": 0, "entcap": 0, "sortsmall": 1}, {"parts": 1}, {"parts": 2}, {"parts": 3, "nt": 64}, {"parts": 7, "entcap": 0},
    {"parts": 2, "tables": 0}, {"parts": 4, "ctas": 1}, {"parts": 2, "sortsmall": 1, "segcap": 0},
    {"flatwalk": 3}, {"flatwalk": 1, "parts": 2}, {"flatwalk": 2, "entcap": 0, "segcap": 3}, {"bmglobal": 1}, {"bmglobal": 1, "parts": 3},
    # group mode of the fused kernel: G environments per CTA, the tree phase shared by all of its warps (the two test
    # environments leave spare slots in the group: the ragged case)
    {"parts": 0, "nt": 128, "group": 7}, {"parts": 0, "nt": 128, "group": 4}, {"parts": 0, "nt": 128, "group": 7, "entcap": 0},
    {"parts": 0, "nt": 128, "group": 5, "segcap": 3, "sortsmall": 1}, {"parts": 0, "group": 3, "nt": 256}, {"parts": 0, "group": 7, "nt": 64},
    {"parts": 0, "nt": 128, "group": 6, "tables": 0}, {"parts": 0, "group": 2, "nt": 256},
    {"sortsmall": 40}, {"sortsmall": 40, "parts": 2, "entcap": 0}, {"parts": 0, "nt": 160}, {"parts": 0, "nt": 192, "entcap": 0},
    # tree structures recomputed every step instead of taken from FlBatch.tree_cache (the default, which every other plan runs)
    {"treecache": 0}, {"treecache": 0, "parts": 2}, {"treecache": 0, "parts": 0, "nt": 128, "group": 7},
    # predicted paths walked every step instead of taken from FlBatch.path_cache; the cache under the other index plans
    {"pathcache": 0}, {"pathcache": 0, "treecache": 0, "parts": 2}, {"pathcache": 1, "segcap": 0, "entcap": 0}, {"pathcache": 1, "flatwalk": 0, "parts": 0, "nt": 128, "group": 5},
]


@pytest.fixture
def obs_plan():
    """Sets fl_observe_override knobs for one test and puts the defaults back afterwards."""
    import flatland_marl_b200 as fb
    lib = fb._lib.lib()
    used = []

    def set_plan(plan):
        for k, v in plan.items():
            assert lib.fl_observe_override(k.encode(), int(v)) == 0
            used.append(k)
    yield set_plan
    for k in used:
        lib.fl_observe_override(k.encode(), -1)


def _first_bad(a, b):
    bad = np.argwhere(np.asarray(a) != np.asarray(b))
    return None if len(bad) == 0 else tuple(int(x) for x in bad[0])


def assert_same(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, got.shape, want.shape)
    if got.dtype.kind == "f":
        ok = (got == want) | (np.isnan(got) & np.isnan(want))
        if not BIT_EXACT_FLOATS:
            ok |= np.isclose(got, want, rtol=RTOL, atol=0, equal_nan=True)
        if not ok.all():
            idx = tuple(int(x) for x in np.argwhere(~ok)[0])
            raise AssertionError("%s: first float mismatch at %s: got %r want %r (%d bad)" %
                                 (what, idx, got[idx], want[idx], int((~ok).sum())))
    else:
        idx = _first_bad(got, want)
        if idx is not None:
            raise AssertionError("%s: first mismatch at %s: got %r want %r (%d bad)" %
                                 (what, idx, got[idx], want[idx], int((got != want).sum())))


def cuda_obs_numpy(batch, e):
    o = {k: v[e].cpu().numpy() for k, v in batch.obs.items()}
    o["attr"] = o.pop("agent_attr")
    o["deadlocked"] = batch.t["deadlocked"][e].cpu().numpy()
    return o


def compare_obs(batch, e, oracle_obs, what):
    o = cuda_obs_numpy(batch, e)
    for k in INT_OBS + ["deadlocked"]:
        assert_same(o[k], oracle_obs[k], "%s %s" % (what, k))
    for k in FLOAT_OBS:
        assert_same(o[k], oracle_obs[k], "%s %s" % (what, k))


def compare_state(batch, e, oracle_state, what):
    s = batch.state_numpy(e)
    for k in STATE_KEYS:
        assert_same(s[k], oracle_state[k], "%s %s" % (what, k))


def run_against_oracle(worlds, actions, scheds, n_steps, check_every=1):
    """Steps E envs on the GPU and E oracle envs in lock-step; compares everything."""
    import torch
    import flatland_marl_b200 as fb
    from oracle import oracle as orc
    ws = []
    for w, s in zip(worlds, scheds):
        w = dict(w)
        w["sched"] = s
        ws.append(w)
    batch = fb.BatchedRailEnv(ws, sched_rows=max(len(s) for s in scheds))
    envs = [orc.OracleEnv(w) for w in ws]
    batch.reset()
    for e, env in enumerate(envs):
        env.reset()
        assert_same(batch.dist_numpy(e)[: env.dist_u16().shape[0]], env.dist_u16(), "env %d distance map" % e)
        compare_state(batch, e, env.state(), "env %d reset" % e)
        compare_obs(batch, e, env.obs(), "env %d reset obs" % e)
    alive = [True] * len(envs)
    for t in range(n_steps):
        act = np.stack([a[t] if t < len(a) else np.zeros_like(a[0]) for a in actions])
        _, rew, don = batch.step(torch.from_numpy(act).to(batch.device))
        rew, don = rew.cpu().numpy(), don.cpu().numpy()
        for e, env in enumerate(envs):
            if not alive[e]:
                continue
            orew, odon = env.step(act[e], scheds[e][t % len(scheds[e])])
            what = "env %d step %d" % (e, t + 1)
            assert_same(rew[e], orew, what + " rewards")
            assert_same(don[e], odon, what + " dones")
            if t % check_every == 0 or odon[-1]:
                compare_state(batch, e, env.state(), what)
                compare_obs(batch, e, env.obs(), what + " obs")
            if odon[-1]:
                alive[e] = False
        if not any(alive):
            break
    return batch


@pytest.mark.parametrize("name,parts", [(n, None) for n in golden_names()] +
                         [(n, 0) for n in ("t00_l1_greedy", "t02_l2_stacking", "t03_l0_random", "t08_l0_greedy", "simple_rail_n4")])
def test_cuda_matches_golden_and_oracle(golden, obs_plan, name, parts):
    """Same world, actions and schedule as the reference run that produced the fixture.  parts = None: the launch shape
    fl_observe picks for a two-environment batch (the split launch); parts = 0: the fused kernel of the large batches."""
    if parts is not None:
        obs_plan({"parts": parts})
    g = golden(name)
    n_steps = int(g["n_steps"])
    # env 0 replays the golden action stream; env 1 a different stream on the same world
    rng = np.random.RandomState(123)
    other = rng.randint(0, 5, size=g["actions"].shape).astype(np.uint8)
    batch = run_against_oracle([g, g], [g["actions"], other], [g["sched"], g["sched"]], n_steps,
                               check_every=1 if int(g["N"]) <= 80 else 4)
    # and directly against what the reference recorded at the final step
    final = {k: g["tr_" + k][n_steps] for k in STATE_KEYS}
    s = batch.state_numpy(0)
    for k in STATE_KEYS:
        assert_same(s[k], final[k], "%s final reference %s" % (name, k))
    assert_same(batch.dist_numpy(0)[: g["dist"].shape[0]], g["dist"], name + " reference distance map")


@pytest.mark.parametrize("name", ["t02_l2_stacking", "t03_l2_stacking", "t03_l1_greedy", "t08_l0_greedy", "t00_l0_random"])
def test_motion_check_through_rail_cell_tables(golden, obs_plan, name):
    """k_step's MotionCheck looks neighbours up in per-rail-cell tables from 64 agents on; forced here ("stepmap" 1) on the
    fixtures with stacked trains, contended cells, swaps and long chains, against the oracle and the reference's final state."""
    obs_plan({"stepmap": 1})
    g = golden(name)
    n_steps = int(g["n_steps"])
    rng = np.random.RandomState(321)
    other = np.where(rng.rand(*g["actions"].shape) < 0.6, 2, rng.randint(0, 5, size=g["actions"].shape)).astype(np.uint8)
    batch = run_against_oracle([g, g], [g["actions"], other], [g["sched"], g["sched"]], n_steps, check_every=4)
    final = {k: g["tr_" + k][n_steps] for k in STATE_KEYS}
    s = batch.state_numpy(0)
    for k in STATE_KEYS:
        assert_same(s[k], final[k], "%s final reference %s" % (name, k))


def test_step_after_done_and_auto_reset(golden):
    import torch
    import flatland_marl_b200 as fb
    g = golden("t00_l1_greedy")
    w = dict(g)
    n_steps = int(g["n_steps"])
    batch = fb.BatchedRailEnv([w], sched_rows=n_steps)
    batch.reset()
    for t in range(n_steps):
        batch.step(torch.from_numpy(g["actions"][t][None]).to(batch.device))
    assert int(batch.t["done_all"][0]) == 1
    batch.step_only(torch.from_numpy(g["actions"][0][None]).to(batch.device))
    assert int(batch.t["status"][0]) & fb._lib.ST_STEP_AFTER_DONE
    # reference-API facade raises like rail_env.py:508-509
    env = fb.RailEnv.from_world(w)
    env.reset()
    for t in range(n_steps):
        acts = {i: int(a) for i, a in enumerate(g["actions"][t]) if a != 255}
        obs, rew, done, info = env.step(acts)
    assert done["__all__"]
    assert [rew[i] for i in range(int(g["N"]))] == list(g["rewards"][n_steps - 1])
    with pytest.raises(Exception):
        env.step({})
    # auto reset: the finished env starts a new episode in place
    batch2 = fb.BatchedRailEnv([w], sched_rows=n_steps, auto_reset=True)
    batch2.reset()
    for t in range(n_steps + 1):
        batch2.step(torch.from_numpy(g["actions"][t % n_steps][None]).to(batch2.device))
    assert int(batch2.t["elapsed"][0]) == 0 and int(batch2.t["done_all"][0]) == 0
    assert int(batch2.t["status"][0]) & fb._lib.ST_AUTO_RESET
    assert (batch2.t["state"][0].cpu().numpy() == 0).all()


def test_facade_matches_reference_api_shapes(golden):
    """The reference-API facade returns what flatland_cutils returns: nested lists with the shapes
    (N,83) (N,31,12) (N,30,3) (N,31) (N,30) and the get_properties triple (treeobs.cpp:612-640)."""
    import flatland_marl_b200 as fb
    g = golden("t00_l0_random")
    env = fb.RailEnv.from_world(dict(g))
    (attr, (forest, adj, norder, eorder)), info = env.reset()
    n = int(g["N"])
    assert np.array(attr).shape == (n, 83) and np.array(forest).shape == (n, 31, 12)
    assert np.array(adj).shape == (n, 30, 3) and np.array(norder).shape == (n, 31) and np.array(eorder).shape == (n, 30)
    np.testing.assert_array_equal(np.array(attr, dtype=np.float32), g["obs0_attr"])
    np.testing.assert_array_equal(np.array(forest, dtype=np.float32), g["obs0_forest"])
    np.testing.assert_array_equal(np.array(adj), g["obs0_adjacency"])
    cfg, props, valid = env.obs_builder.get_properties()
    assert cfg == {"curr_step": 0, "n_agents": n, "max_timesteps": int(g["T"]), "height": int(g["H"]), "width": int(g["W"])}
    assert set(props) == {"dist_target", "deadlocked", "ready_not_depart", "earliest_departure", "latest_arrival", "speed"}
    np.testing.assert_array_equal(np.array(valid, dtype=np.uint8), g["obs0_valid_actions"])
    assert set(info) == {"action_required", "malfunction", "speed", "state"}
    assert not any(env.action_required(a) for a in env.agents)  # everyone is WAITING at reset


def test_host_buffer_step_matches_device_step(golden):
    """fl_step_observe_host (pinned host in/out) gives the same bytes as the device-resident path."""
    import torch
    import flatland_marl_b200 as fb
    g = golden("t02_l1_greedy")
    a = fb.BatchedRailEnv([dict(g)] * 3, sched_rows=int(g["n_steps"]))
    b = fb.BatchedRailEnv([dict(g)] * 3, sched_rows=int(g["n_steps"]))
    a.reset(); b.reset()
    for t in range(40):
        act = np.repeat(g["actions"][t][None], 3, 0)
        a.step(torch.from_numpy(act).to(a.device))
        h = b.step_host(act, wire="compact" if t % 2 else "full", n_chunks=1 + t % 3)
        for k in a.obs:
            np.testing.assert_array_equal(h[k].numpy(), a.obs[k].cpu().numpy(), err_msg="%s step %d" % (k, t))
        np.testing.assert_array_equal(h["rewards"].numpy(), a.rewards.cpu().numpy())
        np.testing.assert_array_equal(h["dones"].numpy(), a.dones.cpu().numpy())


@pytest.mark.parametrize("n_agents", [1, 4, 6])
def test_rail_cycle_world_matches_oracle(n_agents):
    """Switch-free rail cycle + dead-end spur (tests/handmade_worlds.py): the walk that comes back to its own
    first state must end as a terminal node exactly where the reference's visited set stops it."""
    from handmade_worlds import loop_world
    w = loop_world(n_agents)
    rng = np.random.RandomState(5 + n_agents)
    acts = [np.where(rng.rand(w["T"], n_agents) < 0.8, 2, rng.randint(0, 5, (w["T"], n_agents))).astype(np.uint8) for _ in range(3)]
    sched = np.zeros((w["T"], n_agents), np.uint8)
    sched[7, 0] = 5                                   # one malfunction for good measure
    run_against_oracle([w, w, w], acts, [sched, sched, sched], w["T"])


@pytest.mark.parametrize("plan", PLANS, ids=lambda d: ",".join("%s=%s" % kv for kv in d.items()))
def test_every_kernel_plan_matches_oracle(golden, obs_plan, plan):
    """k_observe picks its shared-memory plan, CTA size and fallbacks from the batch shape; each of them must give
    the same bytes.  The overrides go through fl_observe_override (the launch path never reads the environment)."""
    obs_plan(plan)
    g = golden("t03_l1_greedy")
    rng = np.random.RandomState(77)
    other = np.where(rng.rand(*g["actions"].shape) < 0.7, 2, rng.randint(0, 5, size=g["actions"].shape)).astype(np.uint8)
    run_against_oracle([g, g], [g["actions"], other], [g["sched"], g["sched"]], 45, check_every=3)


@pytest.mark.parametrize("group", [4, 7])
def test_group_mode_full_and_ragged_groups(golden, obs_plan, group):
    """Group mode with more environments than one CTA holds: 9 environments are two full groups and a ragged one at G = 4,
    one full and one ragged group at G = 7; warps walk the trees of their neighbours' agents, the bytes stay the same."""
    obs_plan({"parts": 0, "nt": 128, "group": group})
    g = golden("t03_l1_greedy")
    rng = np.random.RandomState(11)
    acts = [g["actions"]] + [np.where(rng.rand(*g["actions"].shape) < p, 2, rng.randint(0, 5, size=g["actions"].shape)).astype(np.uint8)
                             for p in (0.9, 0.8, 0.7, 0.6, 0.5, 0.4, 0.3, 0.2)]
    b = run_against_oracle([g] * 9, acts, [g["sched"]] * 9, 36, check_every=4)
    assert b.observe_plan()["group"] == group


def test_tall_grid_prediction_key_collisions():
    """The reference keys predicted positions by c * W + r (treeobs.cpp:50-65, 379-465); on a grid with H > W two rail
    cells can share a key and then see each other's predictions.  Environment 517 of the Test_03 pack (35 x 30) has
    such a pair on agents' paths from the very first observation (walks.cuh: kcls)."""
    import bench
    w = bench.load_worlds("Test_03", 1, offset=517)[0]
    assert int(w["H"]) > int(w["W"])
    n = int(w["N"])
    rng = np.random.RandomState(3)
    acts = np.where(rng.rand(40, n) < 0.7, 2, rng.randint(0, 5, (40, n))).astype(np.uint8)
    run_against_oracle([w], [acts], [w["sched"]], 40, check_every=4)
