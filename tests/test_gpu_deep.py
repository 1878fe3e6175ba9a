"""The CUDA path against the deep-episode fixtures recorded from the unmodified reference (tests/golden/make_deep_golden.py):
whole episodes of Test_00 / Test_03 / Test_08 / Test_14 — up to 356 of 425 trains on the map, gridlock, sticky deadlocks, paths
beyond the 500-row prediction horizon mid-episode, the episode end at N = 425 — and a SECOND episode after the in-place
auto-reset (= env.reset(False, False) of the reference, including its quirk that arrival_time survives).  Every row is
compared through the CRC32 the reference run recorded per state field and per observation tensor, i.e. bit-exactly, floats
included; on a mismatch the oracle is replayed to that row to name the first differing element."""
import zlib

import numpy as np
import pytest

from test_oracle_deep import OBS_KEYS, STATE_KEYS, action_required_of, deep_names, load_deep
from test_gpu_parity import PLANS, assert_same, cuda_obs_numpy, obs_plan  # noqa: F401

pytestmark = pytest.mark.gpu


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def explain(g, row, key, got):
    """Replays the oracle up to `row` and reports the first element of `key` that differs from `got`."""
    from oracle import oracle as orc
    env = orc.OracleEnv(g)
    r = a = 0
    for steps in g["ep_len"]:
        env.reset()
        if r == row:
            break
        r += 1
        for _ in range(int(steps)):
            env.step(g["actions"][a], g["sched"][a])
            a += 1
            if r == row:
                break
            r += 1
        if r == row:
            break
    want = env.state()[key] if key in STATE_KEYS else env.obs()[key]
    assert_same(got, want, "row %d %s (vs oracle replay)" % (row, key))
    raise AssertionError("row %d %s: CRC differs from the reference's although the oracle replay agrees" % (row, key))


def compare_row(g, batch, e, row, rew, don, name):
    s = batch.state_numpy(e)
    for k in STATE_KEYS:
        if crc(s[k]) != int(g["crc_" + k][row]):
            explain(g, row, k, s[k])
    o = cuda_obs_numpy(batch, e)
    for k in OBS_KEYS:
        if crc(o[k]) != int(g["crc_" + k][row]):
            explain(g, row, k, o[k])
    assert crc(rew) == int(g["crc_rewards"][row]), "%s row %d rewards" % (name, row)
    assert crc(don) == int(g["crc_dones"][row]), "%s row %d dones" % (name, row)


@pytest.mark.parametrize("name,parts", [(n, None) for n in deep_names()] + [("deep_t00_l1", 0), ("deep_t08_l0", 0), ("deep_t14_l0", 0)])
def test_cuda_matches_reference_deep_episode_and_second_episode(obs_plan, name, parts):
    """parts = None: the launch shape fl_observe picks for this small batch (the split launch); 0: the fused kernel."""
    import torch
    import flatland_marl_b200 as fb
    if parts is not None:
        obs_plan({"parts": parts})
    g = load_deep(name)
    N = int(g["N"])
    batch = fb.BatchedRailEnv([dict(g), dict(g)], sched_rows=len(g["sched"]), auto_reset=True)
    batch.reset()
    assert_same(batch.dist_numpy(0)[: g["dist"].shape[0]], g["dist"], name + " reference distance map")
    zero_r, zero_d = np.zeros(N, np.int32), np.zeros(N + 1, np.uint8)
    compare_row(g, batch, 0, 0, zero_r, zero_d, name)
    row, arow = 1, 0
    for ep, steps in enumerate(g["ep_len"]):
        if ep > 0:
            # eval_env.py:81-94 final_metric of the finished episode: arrival ratio, total and normalised reward
            m = batch.final_metric()
            assert m["episodes"] == 2 * ep
            if ep == 1:
                assert m["arrival_ratio"] == g["ep_metric"][0][0] and m["mean_total_reward"] == g["ep_metric"][0][1]
                assert abs(m["mean_norm_reward"] - g["ep_metric"][0][2]) < 1e-12
            # the finished environments restart in place on the next call; the call consumes no action row and no schedule row
            _, rew, don = batch.step(torch.zeros((2, N), dtype=torch.uint8, device=batch.device))
            assert int(batch.t["status"][0]) & fb._lib.ST_AUTO_RESET
            compare_row(g, batch, 0, row, rew[0].cpu().numpy(), don[0].cpu().numpy(), name)
            row += 1
        for _ in range(int(steps)):
            act = np.stack([g["actions"][arow]] * 2)
            arow += 1
            _, rew, don = batch.step(torch.from_numpy(act).to(batch.device))
            compare_row(g, batch, 0, row, rew[0].cpu().numpy(), don[0].cpu().numpy(), name)
            if row % 97 == 0:                                      # the twin environment computes the same bytes
                for k in batch.obs:
                    assert torch.equal(batch.obs[k][0], batch.obs[k][1])
            row += 1
    assert row == int(g["n_rows"])
    stats = batch.t["stats"][0].cpu().numpy()
    assert int(stats[0]) == int((~np.isnan(g["ep_metric"][:, 0])).sum())          # episodes that reached their end


def test_facade_info_dict_and_same_agents_reset_match_reference():
    """RailEnv façade over a whole Test_03 episode and a second one after reset(False, False): get_info_dict()
    (rail_env.py:452-468) against the action_required / malfunction / state the reference returned at every step."""
    import flatland_marl_b200 as fb
    g = load_deep("deep_t03_l3")
    N = int(g["N"])
    env = fb.RailEnv.from_world(dict(g))
    row = arow = 0
    for ep, steps in enumerate(g["ep_len"]):
        obs, info = env.reset() if ep == 0 else env.reset(False, False)

        def check_info(info, row):
            s = env._state_host()
            req = np.unpackbits(g["info_action_required"][row])[:N].astype(bool)
            assert [bool(info["action_required"][i]) for i in range(N)] == list(req), "row %d action_required" % row
            assert [int(info["malfunction"][i]) for i in range(N)] == [int(x) for x in s["mal"]]
            assert [int(info["state"][i]) for i in range(N)] == [int(x) for x in s["state"]]
            assert all(float(info["speed"][i]) == float(g["speed"][i]) for i in range(N))
            for k in STATE_KEYS:
                assert crc(s[k]) == int(g["crc_" + k][row]), "row %d %s" % (row, k)
            assert list(action_required_of(s["state"], s["ctr"])) == list(req)
        check_info(info, row)
        row += 1
        for _ in range(int(steps)):
            a = g["actions"][arow]
            arow += 1
            obs, rew, done, info = env.step({i: int(a[i]) for i in range(N) if a[i] != 255})
            check_info(info, row)
            row += 1
        if ep == 0:
            assert done["__all__"]
            k = list(g["reward_rows"]).index(row - 1)
            assert [rew[i] for i in range(N)] == list(g["reward_vals"][k])


def _resume_rows(g, lo=900):
    return [int(r) for r in g["sample_rows"] if lo <= r < int(g["ep_len"][0]) - 20]


@pytest.mark.parametrize("plan", [{}] + PLANS + [{"segcap": 0, "entcap": 0, "sortsmall": 1}],
                         ids=lambda d: ",".join("%s=%s" % kv for kv in d.items()) or "default")
def test_test14_dense_traffic_every_kernel_plan(obs_plan, plan):
    """Test_14 with hundreds of trains on the map under every k_observe plan (forced segment / entry spill, warp sort of
    large buckets, every CTA size): the recorded reference state of a mid-episode row is uploaded, then 12 steps are
    compared with the reference's CRCs."""
    import torch
    import flatland_marl_b200 as fb
    g = load_deep("deep_t14_l0")
    N = int(g["N"])
    rows = _resume_rows(g)
    assert rows and max(int(g["n_onmap"][r]) for r in rows) >= 300
    obs_plan(plan)
    batch = fb.BatchedRailEnv([dict(g)] * len(rows), sched_rows=len(g["sched"]))
    batch.reset()
    for e, r in enumerate(rows):
        s = {k: g["row%d_%s" % (r, k)] for k in STATE_KEYS}
        batch.set_state(e, s, deadlocked=g["row%d_deadlocked" % r], elapsed=int(g["row_elapsed"][r]))
    batch.observe()
    for k in range(12):
        act = np.stack([g["actions"][r + k] for r in rows])       # episode 1: action row = row index
        _, rew, don = batch.step(torch.from_numpy(act).to(batch.device))
        for e, r in enumerate(rows):
            compare_row(g, batch, e, r + k + 1, rew[e].cpu().numpy(), don[e].cpu().numpy(), "deep_t14_l0 from row %d" % r)
