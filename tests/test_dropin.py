"""Drop-in proof for the reference-facing façade (`flatland_marl_b200.RailEnv` + `TreeObsForRailEnv`).

tests/golden/dropin_t00.npz records the reference's own demo loop (solution/demo.py:101-127): unmodified
`LocalTestEnvWrapper`, unmodified `Actor`, unmodified reference `RailEnv` + `flatland_cutils` (make_dropin_golden.py).

* `test_reference_consumer_runs_unchanged_on_facade` (build container, needs /root/reference): the SAME unmodified
  `LocalTestEnvWrapper` and `Actor` classes are imported from the reference tree and run the demo loop over the façade;
  every array the consumer sees, the rewards, dones and `final_metric()` must equal the reference run — closed loop (the
  actor's actions depend on the observations).  There is no GPU in the build container and the product has no CPU
  fallback, so the façade's batch is a test double that duck-types `BatchedRailEnv` with the C oracle as its engine: this
  test pins the façade and the consumer contract, the GPU test below pins the engine.
* `test_facade_replays_reference_demo_loop` (GPU box, no reference there): the façade on the CUDA path replays the recorded
  action dicts with the calls `LocalTestEnvWrapper` makes and must produce the recorded CRCs."""
import os
import sys
import zlib

import numpy as np
import pytest

from conftest import GOLDEN_DIR

OBS_KEYS = ["agent_attr", "forest", "adjacency", "node_order", "edge_order", "dist_target", "deadlocked", "ready_not_depart",
            "earliest_departure", "latest_arrival", "speed", "valid_actions"]
CFG_KEYS = ["curr_step", "n_agents", "max_timesteps", "height", "width"]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def load_fixture():
    with np.load(os.path.join(GOLDEN_DIR, "dropin_t00.npz")) as z:
        return {k: z[k] for k in z.files}


def check_obs(g, row, o):
    for k in OBS_KEYS:
        assert crc(np.asarray(o[k])) == int(g["crc_" + k][row]), "row %d: the consumer's %s differs from the reference run" % (row, k)
    assert crc(np.array([int(o[k]) for k in CFG_KEYS], np.int64)) == int(g["crc_env_config"][row]), "row %d env_config" % row


class OracleBatch:
    """Test double with the members of BatchedRailEnv that the façade touches, one environment, engine = the C oracle."""
    E = 1

    def __init__(self, world):
        import torch
        from oracle import oracle as orc
        self.torch, self.world, self.env = torch, world, orc.OracleEnv(world)
        self.N, self.device, self.row = int(world["N"]), torch.device("cpu"), 0
        self.t, self.obs = {}, {}

    def _refresh(self, rew, don):
        t = self.torch
        o = self.env.obs()
        self.deadlocked = o["deadlocked"]
        self.obs = {"agent_attr": t.from_numpy(o["attr"][None]), "forest": t.from_numpy(o["forest"][None]),
                    "adjacency": t.from_numpy(o["adjacency"][None]), "node_order": t.from_numpy(o["node_order"][None]),
                    "edge_order": t.from_numpy(o["edge_order"][None]), "valid_actions": t.from_numpy(o["valid_actions"][None]),
                    "dist_target": t.from_numpy(o["dist_target"][None])}
        self.rewards, self.dones = t.from_numpy(rew[None].copy()), t.from_numpy(don[None].copy())
        self.t["elapsed"] = t.tensor([self.env.elapsed])

    def reset(self, env_mask=None, same_agents=False):
        self.env.reset()
        if not same_agents:
            self.row = 0
        self._refresh(np.zeros(self.N, np.int32), np.zeros(self.N + 1, np.uint8))

    def step(self, actions):
        rew, don = self.env.step(actions.numpy()[0], self.world["sched"][self.row])
        self.row += 1
        self._refresh(rew, don)

    def state_numpy(self, e):
        s = self.env.state()
        s["deadlocked"] = self.deadlocked
        return s


@pytest.mark.reference
def test_reference_consumer_runs_unchanged_on_facade():
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("needs the reference tree (build container)")
    rh.load()
    sol = os.path.join(rh.REFERENCE_ROOT, "solution")
    if sol not in sys.path:
        sys.path.insert(0, sol)
    sys.path.insert(0, GOLDEN_DIR)
    from eval_env import LocalTestEnvWrapper                 # UNMODIFIED reference consumer (solution/eval_env.py:97-114)
    import make_dropin_golden as md
    import flatland_marl_b200 as fb
    g = load_fixture()
    world = {k: g[k] for k in ("H", "W", "N", "T", "grid", "init_pos", "init_dir", "target", "speed", "earliest", "latest", "sched")}
    env = fb.RailEnv(width=int(g["W"]), height=int(g["H"]), number_of_agents=int(g["N"]), world=world, batch=OracleBatch(world), index=0)
    wrapper = LocalTestEnvWrapper(env)
    actor = md.make_actor()                                   # UNMODIFIED solution/plfActor.py Actor + nn/net_tree.py Network
    rows = []

    def on_step(obs, passed_on, rew, done):
        row = len(rows)
        check_obs(g, row, obs[0])
        if passed_on is not None:
            n = int(g["N"])
            a = np.full(n, 255, np.uint8)
            for i, v in passed_on.items():
                a[i] = v
            np.testing.assert_array_equal(a, g["actions"][row - 1], err_msg="step %d: actions passed on" % row)
            assert [rew[i] for i in range(n)] == list(g["rewards"][row - 1])
            assert [int(done[i]) for i in range(n)] + [int(done["__all__"])] == list(g["dones"][row - 1])
        rows.append(row)

    metric = md.demo_loop(wrapper, actor, on_step)
    assert len(rows) == len(g["actions"]) + 1
    np.testing.assert_array_equal(np.array(metric, np.float64), g["final_metric"])


def consumer_view(env, feature):
    """What LocalTestEnvWrapper hands its caller: update_obs_properties + parse_features (solution/eval_env.py:56-79)."""
    env_config, agents_properties, valid_actions = env.obs_builder.get_properties()
    props = {}
    props.update(env_config)
    props.update(agents_properties)
    props["valid_actions"] = valid_actions
    o = {"agent_attr": np.array(feature[0]), "forest": np.array(feature[1][0]), "adjacency": np.array(feature[1][1]),
         "node_order": np.array(feature[1][2]), "edge_order": np.array(feature[1][3])}
    o["forest"][o["forest"] == np.inf] = -1
    o.update(props)
    return o


@pytest.mark.gpu
def test_facade_replays_reference_demo_loop():
    import flatland_marl_b200 as fb
    g = load_fixture()
    n = int(g["N"])
    world = {k: g[k] for k in ("H", "W", "N", "T", "grid", "init_pos", "init_dir", "target", "speed", "earliest", "latest", "sched")}
    env = fb.RailEnv.from_world(world)
    feature, _ = env.reset()                                                       # eval_env.py:101-106
    check_obs(g, 0, consumer_view(env, feature))
    for t in range(len(g["actions"])):
        required = {i: env.action_required(a) for i, a in enumerate(env.agents)}    # eval_env.py:27-39
        acts = {i: int(a) for i, a in enumerate(g["actions"][t]) if a != 255}
        assert all(required[i] for i in acts), "step %d: the reference passed on an action the façade says is not required" % t
        feature, reward, done, info = env.step(acts)                                # eval_env.py:108-114
        check_obs(g, t + 1, consumer_view(env, feature))
        assert [reward[i] for i in range(n)] == list(g["rewards"][t])
        assert [int(done[i]) for i in range(n)] + [int(done["__all__"])] == list(g["dones"][t])
    assert done["__all__"]
    # eval_env.py:81-94 final_metric on the façade's attributes
    n_arrival = sum(1 for a in env.agents if a.position is None and a.state != fb.TrainState.READY_TO_DEPART)
    total = sum(env.rewards_dict.values())
    metric = (n_arrival / env.get_num_agents(), total, 1 + total / env._max_episode_steps / env.get_num_agents())
    np.testing.assert_array_equal(np.array(metric, np.float64), g["final_metric"])
