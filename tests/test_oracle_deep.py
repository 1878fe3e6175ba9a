"""Pins the C oracle to the deep-episode fixtures (tests/golden/make_deep_golden.py): whole episodes of Test_03,
Test_08 and Test_14 recorded from the unmodified reference — hundreds of trains on the map, gridlock, the episode
end at N = 425 — followed by a second episode after `reset(False, False)`.  Every row holds a CRC32 per state
field and per observation tensor; sampled rows hold the full tensors."""
import os
import zlib

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import oracle as orc

STATE_KEYS = ["pos", "dir", "state", "ctr", "mal", "nmal", "saved", "arrival", "old_pos", "old_dir", "sig_mal"]
OBS_KEYS = ["attr", "forest", "adjacency", "node_order", "edge_order", "valid_actions", "dist_target", "deadlocked"]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def deep_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("deep_") and f.endswith(".npz"))


def load_deep(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def action_required_of(state, ctr):
    """rail_env.py:243-258"""
    return (state == 1) | ((state >= 3) & (state <= 5) & (ctr == 0))


def obs_rows_to_check(g):
    """Every row for the small configurations; for Test_14 (425 agents) the oracle's observation is checked on every 8th
    row, on all sampled rows, around the busiest stretch and around both episode boundaries (the state is checked on
    every row regardless)."""
    n_rows = int(g["n_rows"])
    if int(g["N"]) <= 100:
        return set(range(n_rows))
    rows = set(range(0, n_rows, 8)) | set(int(r) for r in g["sample_rows"])
    end1 = int(g["ep_len"][0])
    rows |= set(range(max(0, end1 - 30), min(n_rows, end1 + 12)))
    rows |= set(range(n_rows - 10, n_rows))
    busiest = int(np.argmax(g["n_onmap"][: end1 + 1]))
    rows |= set(range(max(0, busiest - 10), busiest + 1))
    return rows


def check_row(g, row, env, rewards, dones, name, with_obs):
    s = env.state()
    for k in STATE_KEYS:
        if "row%d_%s" % (row, k) in g:
            np.testing.assert_array_equal(s[k], g["row%d_%s" % (row, k)], err_msg="%s row %d %s" % (name, row, k))
        assert crc(s[k]) == int(g["crc_" + k][row]), "%s: state crc mismatch at row %d for %s" % (name, row, k)
    assert crc(rewards) == int(g["crc_rewards"][row]) and crc(dones) == int(g["crc_dones"][row]), "%s row %d rewards/dones" % (name, row)
    req = np.unpackbits(g["info_action_required"][row])[: env.N].astype(bool)
    np.testing.assert_array_equal(action_required_of(s["state"], s["ctr"]), req, err_msg="%s row %d action_required" % (name, row))
    assert int(((s["state"] >= 3) & (s["state"] <= 5)).sum()) == int(g["n_onmap"][row])
    if with_obs:
        o = env.obs()
        for k in OBS_KEYS:
            if "row%d_%s" % (row, k) in g:
                np.testing.assert_array_equal(o[k], g["row%d_%s" % (row, k)], err_msg="%s row %d %s" % (name, row, k))
            assert crc(o[k]) == int(g["crc_" + k][row]), "%s: obs crc mismatch at row %d for %s" % (name, row, k)


@pytest.mark.parametrize("name", deep_names())
def test_oracle_matches_reference_deep_episode(name):
    g = load_deep(name)
    N = int(g["N"])
    env = orc.OracleEnv(g)
    rows = obs_rows_to_check(g)
    zero_r, zero_d = np.zeros(N, np.int32), np.zeros(N + 1, np.uint8)
    row = arow = 0
    rewards_seen = {}
    for ep, steps in enumerate(g["ep_len"]):
        env.reset()                                         # reset(False, False): nothing regenerated, the schedule carries on
        if ep == 0:
            np.testing.assert_array_equal(env.dist_u16(), g["dist"], err_msg=name + " distance map")
        check_row(g, row, env, zero_r, zero_d, name, row in rows)
        row += 1
        for _ in range(int(steps)):
            rew, don = env.step(g["actions"][arow], g["sched"][arow])
            arow += 1
            check_row(g, row, env, rew, don, name, row in rows)
            if rew.any():
                rewards_seen[row] = rew.copy()
            row += 1
    assert row == int(g["n_rows"])
    assert sorted(rewards_seen) == [int(r) for r in g["reward_rows"]]
    for k, r in enumerate(g["reward_rows"]):
        np.testing.assert_array_equal(rewards_seen[int(r)], g["reward_vals"][k])
    if N >= 400:
        assert int(g["n_onmap"].max()) >= 300             # the regime the fixture exists for
