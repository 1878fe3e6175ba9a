"""Records the reference's own demo loop (solution/demo.py:101-127) — UNMODIFIED `LocalTestEnvWrapper`
(solution/eval_env.py:97-114), UNMODIFIED `Actor` (solution/plfActor.py; random-init weights `init_weights(0)`: the
reference ships no checkpoint) and the unmodified reference `RailEnv` + `flatland_cutils.TreeObsForRailEnv` — so that the
GPU box, where the reference does not exist, can check that `flatland_marl_b200.RailEnv` hands the same consumer the same
observations, rewards, dones and final metric for the same actions.

  python tests/golden/make_dropin_golden.py          # needs /root/reference; writes tests/golden/dropin_t00.npz

Per step the file holds the action dict the wrapper passed on (after parse_actions dropped the agents for which no action is
required), CRC32s of every array `parse_features` produced (float64 / int64, as the reference's consumer sees them) and of the
property lists, the rewards and dones, and at the end `final_metric()`."""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import ref_harness as rh  # noqa: E402
import make_golden as mg  # noqa: E402

OBS_KEYS = ["agent_attr", "forest", "adjacency", "node_order", "edge_order", "dist_target", "deadlocked", "ready_not_depart",
            "earliest_departure", "latest_arrival", "speed", "valid_actions"]
CFG_KEYS = ["curr_step", "n_agents", "max_timesteps", "height", "width"]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def obs_crcs(obs):
    """obs: what LocalTestEnvWrapper.reset/step return ([dict]); every value as the numpy array the consumer gets."""
    o = obs[0]
    out = {k: crc(np.asarray(o[k])) for k in OBS_KEYS}
    out["env_config"] = crc(np.array([int(o[k]) for k in CFG_KEYS], np.int64))
    return out


def make_actor():
    import torch
    from plfActor import Actor
    from nn.net_tree import Network
    import flatland_marl_b200.policy_weights as pw
    net = Network()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in pw.init_weights(0).items()}, strict=True)
    net.eval()
    actor = Actor.__new__(Actor)
    actor.net = net
    return actor


def demo_loop(env_wrapper, actor, on_step):
    """solution/demo.py:101-127 without rendering."""
    n_agents = env_wrapper.env.number_of_agents
    obs = env_wrapper.reset()
    on_step(obs, None, None, None)
    while True:
        va = env_wrapper.get_valid_actions()
        action = actor.get_actions(obs, va, n_agents)
        passed_on = env_wrapper.parse_actions(action)          # what step() will hand to env.step (eval_env.py:33-39,109)
        obs, all_rewards, done = env_wrapper.step(action)
        on_step(obs, passed_on, all_rewards, done)
        if done["__all__"]:
            return env_wrapper.final_metric()


def main():
    rh.load()
    sys.path.insert(0, os.path.join(rh.REFERENCE_ROOT, "solution"))
    from eval_env import LocalTestEnvWrapper
    env = rh.make_env("Test_00", rh.csv_seed(0, 3), mal_interval=80)
    wrapper = LocalTestEnvWrapper(env)
    actor = make_actor()
    rows = {k: [] for k in OBS_KEYS + ["env_config"]}
    actions, rewards, dones = [], [], []
    world = {}

    def on_step(obs, passed_on, rew, done):
        if not world:                                            # right after reset: the generated world and its schedule
            world.update(mg.static_of(env))
            world["sched"] = mg.draw_schedule(env, int(world["T"]) + 2)
        for k, v in obs_crcs(obs).items():
            rows[k].append(v)
        if passed_on is not None:
            n = int(world["N"])
            a = np.full(n, 255, np.uint8)
            for i, v in passed_on.items():
                a[i] = v
            actions.append(a)
            rewards.append([rew[i] for i in range(n)])
            dones.append([done[i] for i in range(n)] + [done["__all__"]])

    metric = demo_loop(wrapper, actor, on_step)
    out = dict(world)
    out["actions"] = np.stack(actions)
    out["rewards"] = np.array(rewards, np.int32)
    out["dones"] = np.array(dones, np.uint8)
    out["final_metric"] = np.array(metric, np.float64)
    for k, v in rows.items():
        out["crc_" + k] = np.array(v, np.uint32)
    path = os.path.join(HERE, "dropin_t00.npz")
    np.savez_compressed(path, **out)
    print("dropin_t00: %d steps, final metric %s, %.1f KB" % (len(actions), metric, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
