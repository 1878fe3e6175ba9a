"""Pins the C oracle (oracle/flatland_oracle.c) against golden vectors recorded from the unmodified
reference (tests/golden/make_golden.py): every agent field after every step, rewards, dones, the
distance maps, a CRC of every observation tensor at every step (bit-exact, floats included) and the
full observation tensors at the sampled steps."""
import zlib

import numpy as np
import pytest

from conftest import golden_names
from oracle import oracle as orc

STATE_KEYS = ["pos", "dir", "state", "ctr", "mal", "nmal", "saved", "arrival", "old_pos", "old_dir", "sig_mal"]
OBS_KEYS = ["attr", "forest", "adjacency", "node_order", "edge_order", "valid_actions", "dist_target", "deadlocked"]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def check_obs(g, k, o, name):
    for key in OBS_KEYS:
        if "obs%d_%s" % (k, key) in g:
            np.testing.assert_array_equal(o[key], g["obs%d_%s" % (k, key)], err_msg="%s step %d %s" % (name, k, key))
        assert crc(o[key]) == int(g["crc_" + key][k]), "%s: crc mismatch at step %d for %s" % (name, k, key)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_episode(golden, name):
    g = golden(name)
    env = orc.OracleEnv(g)
    env.reset()
    np.testing.assert_array_equal(env.dist_u16(), g["dist"], err_msg=name + " distance map")
    np.testing.assert_array_equal(env.target_slots(), g["tgt_slot"])
    check_obs(g, 0, env.obs(), name)
    for t in range(int(g["n_steps"])):
        rewards, dones = env.step(g["actions"][t], g["sched"][t])
        s = env.state()
        for key in STATE_KEYS:
            np.testing.assert_array_equal(s[key], g["tr_" + key][t + 1], err_msg="%s step %d %s" % (name, t + 1, key))
        np.testing.assert_array_equal(rewards, g["rewards"][t], err_msg="%s step %d rewards" % (name, t + 1))
        np.testing.assert_array_equal(dones, g["dones"][t], err_msg="%s step %d dones" % (name, t + 1))
        check_obs(g, t + 1, env.obs(), name)
    if g["dones"][-1][-1]:
        with pytest.raises(Exception):
            env.step(g["actions"][0], g["sched"][0])


def test_oracle_motion_check_matches_reference(golden):
    import os
    with np.load(os.path.join(os.path.dirname(__file__), "golden", "motion_cases.npz")) as z:
        cur, nxt, n, ok = z["cur"], z["nxt"], z["n"], z["can_move"]
    for k in range(len(n)):
        got = orc.motion_check(cur[k, : n[k]], nxt[k, : n[k]])
        np.testing.assert_array_equal(got, ok[k, : n[k]], err_msg="case %d" % k)


def test_oracle_runs_rail_cycle_world():
    """The hand-made loop world (tests/handmade_worlds.py) exercises the `visited` branch of
    _explore_branch; the oracle must produce terminal nodes (dist_to_next_branch = inf -> -1) there."""
    from handmade_worlds import loop_world
    w = loop_world(4)
    env = orc.OracleEnv(w)
    env.reset()
    rng = np.random.RandomState(3)
    saw_cycle_node = False
    for t in range(50):
        act = np.where(rng.rand(4) < 0.8, 2, rng.randint(0, 5, 4)).astype(np.uint8)
        _, dones = env.step(act, np.zeros(4, np.uint8))
        f = env.obs()["forest"]
        # a real node (some feature != -1) whose dist_to_next_branch (feature 5) is -1 ended on a revisited state
        real = (f != -1).any(axis=2)
        saw_cycle_node |= bool((real[:, 1:] & (f[:, 1:, 5] == -1)).any())
        if dones[-1]:
            break
    assert saw_cycle_node
