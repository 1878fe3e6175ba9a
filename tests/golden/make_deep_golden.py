"""Deep-episode golden vectors: whole episodes of the large configurations recorded from the UNMODIFIED
reference (flatland-rl RailEnv.step + flatland_cutils TreeObsForRailEnv through oracle/ref_harness.py),
followed by a second episode after `env.reset(False, False)` — the reference call that the CUDA path's
in-place auto-reset stands for (rail_env.py:260-357 with nothing regenerated: no RNG is consumed, the
malfunction generator simply carries on).

  python tests/golden/make_deep_golden.py [name ...]        # needs /root/reference; Test_14 takes ~1 h

Round-1 fixtures (make_golden.py) stop at 40 steps of Test_14 with 12 trains on the map; these cover the
regime the large configurations exist for: hundreds of trains on the map, a dense prediction index, paths
beyond the 500-row horizon mid-episode, gridlock / sticky deadlocks, the episode end at N = 425.

The action stream is fixed BEFORE the reference runs: a distance-map-greedy policy with random stops, invalid
values and absent keys, computed by stepping the C oracle on the same world and malfunction schedule.  (If the
oracle disagreed with the reference the stream would merely be a different arbitrary stream; what is recorded is
only what the reference did with it.)  To keep the files small every row (reset observation + one row per step,
for both episodes) holds a CRC32 per state field and per observation tensor; full state and observation
tensors are kept at a few sampled rows; rewards are stored where non-zero (episode ends) and as CRCs.
"""
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import ref_harness as rh  # noqa: E402
from oracle import oracle as orc  # noqa: E402
import make_golden as mg  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ABSENT = 255
STATE_KEYS = ["pos", "dir", "state", "ctr", "mal", "nmal", "saved", "arrival", "old_pos", "old_dir", "sig_mal"]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def plan_actions(world, sched, ep_steps, seed, p_greedy=0.9, p_stop=0.5, absent_p=0.05, invalid_p=0.01):
    """Action stream for both episodes from a greedy walk of the C oracle (see the module docstring)."""
    env = orc.OracleEnv(world)
    N = int(world["N"])
    grid = np.asarray(world["grid"])
    rng = np.random.RandomState(seed)
    out = []
    row = 0
    for steps in ep_steps:
        env.reset()
        dist = env.dist_u16().astype(np.int64)
        slots = env.target_slots()
        for _ in range(steps):
            s = env.state()
            acts = np.zeros(N, np.uint8)
            u = rng.rand(N)
            v = rng.rand(N)
            rnd = rng.randint(0, 5, N)
            for i in range(N):
                if u[i] > p_greedy:
                    acts[i] = 4 if v[i] < p_stop else rnd[i]
                    continue
                st = s["state"][i]
                if st == 1:
                    acts[i] = 2
                elif st in (3, 4):
                    r, c = s["pos"][i]
                    d = int(s["dir"][i])
                    nib = (int(grid[r, c]) >> ((3 - d) * 4)) & 0xF
                    best, ba = 1 << 30, 2
                    for a in (1, 2, 3):
                        nd = (d + a - 2) % 4
                        if (nib >> (3 - nd)) & 1:
                            rr, cc = r + (-1, 0, 1, 0)[nd], c + (0, 1, 0, -1)[nd]
                            if dist[slots[i], rr, cc, nd] < best:
                                best, ba = dist[slots[i], rr, cc, nd], a
                    acts[i] = ba
            w = rng.rand(N)
            acts[w < invalid_p] = 5 + rng.randint(0, 3, int((w < invalid_p).sum()))
            acts[w > 1.0 - absent_p] = ABSENT
            out.append(acts)
            _, don = env.step(acts, sched[row])
            row += 1
            if don[N]:
                break
    return np.stack(out)


def record_deep(cfg, test, level, mal_interval, second_steps, seed, n_samples=5, extra_samples=(), policy={}):
    t_start = time.time()
    env = rh.make_env(cfg, rh.csv_seed(test, level), mal_interval=mal_interval)
    obs, info = env.reset()
    st = mg.static_of(env)
    N, T = int(st["N"]), int(st["T"])
    sched = mg.draw_schedule(env, T + min(second_steps, T))
    print("%s: reset done after %.0f s (N=%d, T=%d)" % (cfg, time.time() - t_start, N, T), flush=True)
    second_steps = min(second_steps, T)
    actions = plan_actions(st, sched, [T, second_steps], seed, **policy)
    print("%s: %d planned action rows after %.0f s" % (cfg, len(actions), time.time() - t_start), flush=True)

    crcs = {k: [] for k in STATE_KEYS + mg.OBS_KEYS + ["rewards", "dones"]}
    info_req, n_onmap, n_done, n_deadlocked, row_elapsed = [], [], [], [], []
    samples = {}
    reward_rows, reward_vals = [], []
    want = set()

    def record(row, obs, info, rew, don):
        s = mg.state_of(env)
        o = mg.obs_of(env, obs)
        for k in STATE_KEYS:
            crcs[k].append(crc(s[k]))
        for k in mg.OBS_KEYS:
            crcs[k].append(crc(o[k]))
        crcs["rewards"].append(crc(rew))
        crcs["dones"].append(crc(don))
        info_req.append(np.packbits(np.array([bool(info["action_required"][i]) for i in range(N)], np.uint8)))
        assert all(int(info["malfunction"][i]) == int(s["mal"][i]) and int(info["state"][i]) == int(s["state"][i]) for i in range(N))
        n_onmap.append(int(((s["state"] >= 3) & (s["state"] <= 5)).sum()))
        n_done.append(int((s["state"] == 6).sum()))
        n_deadlocked.append(int(o["deadlocked"].sum()))
        row_elapsed.append(int(env._elapsed_steps))
        if rew.any():
            reward_rows.append(row)
            reward_vals.append(rew.copy())
        if row in want:
            for k in STATE_KEYS:
                samples["row%d_%s" % (row, k)] = s[k]
            for k in mg.OBS_KEYS:
                samples["row%d_%s" % (row, k)] = o[k]

    zero_r, zero_d = np.zeros(N, np.int32), np.zeros(N + 1, np.uint8)
    ep_len, ep_metric = [], []
    row, arow = 0, 0
    for ep, steps in enumerate([T, second_steps]):
        if ep == 1:
            obs, info = env.reset(False, False)            # what FL_FLAG_AUTO_RESET stands for
        base = row
        want = set(base + int(x) for x in np.linspace(0, steps, n_samples)) | set(base + int(x) for x in (extra_samples if ep == 0 else ()))
        record(row, obs, info, zero_r, zero_d)
        row += 1
        k = 0
        for k in range(1, steps + 1):
            a = actions[arow]
            arow += 1
            obs, rew, done, info = env.step({i: int(a[i]) for i in range(N) if a[i] != ABSENT})
            r = np.array([rew[i] for i in range(N)], np.int32)
            d = np.array([done[i] for i in range(N)] + [done["__all__"]], np.uint8)
            record(row, obs, info, r, d)
            row += 1
            if k % 100 == 0:
                print("%s: episode %d step %d, %d on map, %d done, %d deadlocked, %.0f s" %
                      (cfg, ep, k, n_onmap[-1], n_done[-1], n_deadlocked[-1], time.time() - t_start), flush=True)
            if done["__all__"]:
                break
        ep_len.append(k)
        if done["__all__"]:
            # solution/eval_env.py:81-94 final_metric, same expressions
            arrived = sum(1 for ag in env.agents if ag.position is None and int(ag.state) != 1)
            total = sum(env.rewards_dict.values())
            ep_metric.append([arrived / N, total, 1 + total / env._max_episode_steps / N])
        else:
            ep_metric.append([np.nan, np.nan, np.nan])
    assert arow == len(actions), (arow, len(actions))
    out = dict(st)
    out["sched"] = sched[:arow]
    out["actions"] = actions
    out["ep_len"] = np.array(ep_len, np.int32)
    out["ep_metric"] = np.array(ep_metric, np.float64)
    out["n_rows"] = np.int32(row)
    for k, v in crcs.items():
        out["crc_" + k] = np.array(v, np.uint32)
    out["info_action_required"] = np.stack(info_req)
    out["n_onmap"] = np.array(n_onmap, np.int32)
    out["n_done"] = np.array(n_done, np.int32)
    out["n_deadlocked"] = np.array(n_deadlocked, np.int32)
    out["row_elapsed"] = np.array(row_elapsed, np.int32)
    out["reward_rows"] = np.array(reward_rows, np.int32)
    out["reward_vals"] = np.stack(reward_vals) if reward_vals else np.zeros((0, N), np.int32)
    out["sample_rows"] = np.array(sorted(set(int(k[3:].split("_")[0]) for k in samples)), np.int32)
    out.update(samples)
    return out


JOBS = {
    # name: (config, test, level, malfunction interval, steps of the second episode, action seed, extra sample rows)
    # Test_00: nearly all trains arrive, and arrive AGAIN in the second episode, where the reference leaves them standing on
    # their target cell in state DONE (EnvAgent.reset keeps arrival_time, so handle_done_state does nothing)
    "deep_t00_l1": ("Test_00", 0, 1, 60, 10 ** 6, 30, ()),
    "deep_t03_l3": ("Test_03", 3, 3, 450, 150, 31, ()),
    "deep_t08_l0": ("Test_08", 8, 0, 720, 100, 32, (600, 900)),
    "deep_t14_l0": ("Test_14", 14, 0, 3600, 40, 33, (950, 1700, 2300)),
}


def main():
    only = sys.argv[1:]
    for name, (cfg, test, level, mal, second, seed, extra) in JOBS.items():
        if only and name not in only:
            continue
        d = record_deep(cfg, test, level, mal, second, seed, extra_samples=extra,
                        policy=dict(p_greedy=0.97) if name == "deep_t00_l1" else {})
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print("%-16s %8.1f KB rows=%d ep_len=%s max on map=%d max deadlocked=%d metric=%s" %
              (name, os.path.getsize(path) / 1024, int(d["n_rows"]), d["ep_len"].tolist(), int(d["n_onmap"].max()),
               int(d["n_deadlocked"].max()), d["ep_metric"].tolist()), flush=True)


if __name__ == "__main__":
    main()
