"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference
(flatland-rl `RailEnv.step` + the flatland_cutils C++ `TreeObsForRailEnv`, see oracle/ref_harness.py)
in the build container.  The reference ships no tests or known-answer vectors of its own
(SURVEY.md §4), so these files are the pin for the oracle:

  python tests/golden/make_golden.py            # regenerate everything (needs /root/reference)

Every fixture holds the generated world itself (grid, agents, timetable, malfunction schedule), the
action stream, the reference's per-step agent state, rewards and dones, a CRC32 of every observation
tensor at every step and the full observation tensors at a few sampled steps.  Floats are stored as
the float32 values the reference produced.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ABSENT = 255  # action value meaning "key not present in action_dict"


def draw_schedule(env, n_steps):
    """Pre-draws the malfunction onsets exactly as RailEnv.step would consume them
    (malfunction_generators.py:46-53 called once per agent per step, rail_env.py:524)."""
    import copy
    rng = copy.deepcopy(env.np_random)
    gen = env.malfunction_generator
    n = len(env.agents)
    sched = np.zeros((n_steps, n), dtype=np.uint8)
    for t in range(n_steps):
        for i in range(n):
            sched[t, i] = gen.generate(rng).num_broken_steps
    return sched


def static_of(env):
    ags = env.agents
    dm = env.distance_map.get()
    targets = []
    slot = []
    for a in ags:
        if a.target not in targets:
            targets.append(a.target)
        slot.append(targets.index(a.target))
    first = [slot.index(s) for s in range(len(targets))]
    dist = dm[first].copy()
    dist_u16 = np.where(np.isinf(dist), 65535, dist).astype(np.uint16)
    return dict(
        H=np.int32(env.height), W=np.int32(env.width), N=np.int32(len(ags)),
        T=np.int32(env._max_episode_steps),
        grid=np.asarray(env.rail.grid, dtype=np.uint16),
        init_pos=np.array([a.initial_position for a in ags], dtype=np.int16),
        init_dir=np.array([int(a.initial_direction) for a in ags], dtype=np.uint8),
        target=np.array([a.target for a in ags], dtype=np.int16),
        speed=np.array([a.speed_counter.speed for a in ags], dtype=np.float64),
        earliest=np.array([a.earliest_departure for a in ags], dtype=np.int32),
        latest=np.array([a.latest_arrival for a in ags], dtype=np.int32),
        tgt_slot=np.array(slot, dtype=np.int16),
        dist=dist_u16,  # [n_unique_targets, H, W, 4], 65535 = unreachable
    )


def state_of(env):
    ags = env.agents
    n = len(ags)
    out = dict(
        pos=np.full((n, 2), -1, np.int16), dir=np.zeros(n, np.uint8), state=np.zeros(n, np.uint8),
        ctr=np.zeros(n, np.uint8), mal=np.zeros(n, np.uint8), nmal=np.zeros(n, np.uint16),
        saved=np.zeros(n, np.uint8), arrival=np.full(n, -1, np.int32),
        old_pos=np.full((n, 2), -1, np.int16), old_dir=np.full(n, -1, np.int8),
        sig_mal=np.zeros(n, np.uint8))
    for i, a in enumerate(ags):
        if a.position is not None:
            out["pos"][i] = a.position
        out["dir"][i] = int(a.direction)
        out["state"][i] = int(a.state)
        out["ctr"][i] = a.speed_counter.counter
        out["mal"][i] = a.malfunction_handler.malfunction_down_counter
        out["nmal"][i] = a.malfunction_handler.num_malfunctions
        out["saved"][i] = 0 if a.action_saver.saved_action is None else int(a.action_saver.saved_action)
        out["arrival"][i] = -1 if a.arrival_time is None else a.arrival_time
        if a.old_position is not None:
            out["old_pos"][i] = a.old_position
        out["old_dir"][i] = -1 if a.old_direction is None else int(a.old_direction)
        out["sig_mal"][i] = int(a.state_machine.st_signals.in_malfunction)
    return out


def obs_of(env, obs):
    attr, forest = obs
    env_cfg, props, va = env.obs_builder.get_properties()
    dist_target = np.array(props["dist_target"], dtype=np.float32)
    return dict(
        attr=np.array(attr, dtype=np.float32),
        forest=np.array(forest[0], dtype=np.float32),
        adjacency=np.array(forest[1], dtype=np.int32),
        node_order=np.array(forest[2], dtype=np.int32),
        edge_order=np.array(forest[3], dtype=np.int32),
        valid_actions=np.array(va, dtype=np.uint8),
        dist_target=dist_target,
        deadlocked=np.array(props["deadlocked"], dtype=np.uint8),
    )


OBS_KEYS = ["attr", "forest", "adjacency", "node_order", "edge_order", "valid_actions",
            "dist_target", "deadlocked"]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def greedy_action(env, i):
    """Distance-map descent used by the 'greedy' action policy (drives agents to their targets so
    DONE states and all-arrived episode ends are covered)."""
    a = env.agents[i]
    st = int(a.state)
    if st == 1:
        return 2
    if st not in (3, 4) or a.position is None:
        return 0
    dm = env.distance_map.get()
    grid = env.rail.grid
    r, c = a.position
    d = int(a.direction)
    nib = (int(grid[r, c]) >> ((3 - d) * 4)) & 0xF
    best, best_a = np.inf, 2
    for act in (1, 2, 3):
        nd = (d + act - 2) % 4
        if not (nib >> (3 - nd)) & 1:
            continue
        rr, cc = r + (-1, 0, 1, 0)[nd], c + (0, 1, 0, -1)[nd]
        v = dm[i, rr, cc, nd]
        if v < best:
            best, best_a = v, act
    return best_a


def run_episode(env, policy, seed, max_steps=None, n_samples=6, absent_p=0.1, invalid_p=0.02):
    rng = np.random.RandomState(seed)
    obs, _ = env.reset()
    st = static_of(env)
    n = int(st["N"])
    T = int(st["T"])
    steps = T if max_steps is None else min(T, max_steps)
    sched = draw_schedule(env, steps)
    sample_at = sorted(set([0, 1, 2] + [int(x) for x in np.linspace(3, steps, n_samples)]))
    rec = {k: [] for k in ["pos", "dir", "state", "ctr", "mal", "nmal", "saved", "arrival",
                           "old_pos", "old_dir", "sig_mal"]}
    crcs = {k: [] for k in OBS_KEYS}
    samples = {}
    actions = np.zeros((steps, n), np.uint8)
    rewards = np.zeros((steps, n), np.int32)
    dones = np.zeros((steps, n + 1), np.uint8)

    def record(k, obs):
        s = state_of(env)
        for key in rec:
            rec[key].append(s[key])
        o = obs_of(env, obs)
        for key in OBS_KEYS:
            crcs[key].append(crc(o[key]))
        if k in sample_at:
            for key in OBS_KEYS:
                samples["obs%d_%s" % (k, key)] = o[key]

    record(0, obs)
    n_done = 0
    n_stacked = 0
    for t in range(steps):
        act = {}
        for i in range(n):
            u = rng.rand()
            if policy == "random":
                a = int(rng.randint(0, 5))
            elif policy == "forward":
                a = 2 if rng.rand() < 0.85 else int(rng.randint(0, 5))
            else:
                a = greedy_action(env, i) if rng.rand() < 0.9 else int(rng.randint(0, 5))
            if u < invalid_p:
                a = 5 + int(rng.randint(0, 3))  # invalid action values (-> DO_NOTHING)
            if u > 1.0 - absent_p:
                actions[t, i] = ABSENT
            else:
                actions[t, i] = a
                act[i] = a
        obs, rew, done, _ = env.step(act)
        rewards[t] = [rew[i] for i in range(n)]
        dones[t, :n] = [done[i] for i in range(n)]
        dones[t, n] = done["__all__"]
        record(t + 1, obs)
        onmap = [a.position for a in env.agents if a.position is not None]
        n_stacked += len(onmap) != len(set(onmap))
        n_done = t + 1
        if done["__all__"]:
            break
    out = dict(st)
    out["sched"] = sched[:n_done]
    out["actions"] = actions[:n_done]
    out["rewards"] = rewards[:n_done]
    out["dones"] = dones[:n_done]
    out["n_steps"] = np.int32(n_done)
    out["n_stacked_steps"] = np.int32(n_stacked)
    out["sample_steps"] = np.array([k for k in sample_at if k <= n_done], np.int32)
    for key in rec:
        out["tr_" + key] = np.stack(rec[key])
    for key in OBS_KEYS:
        out["crc_" + key] = np.array(crcs[key], np.uint32)
    out.update({k: v for k, v in samples.items() if int(k[3:].split("_")[0]) <= n_done})
    return out


def simple_rail_env(n_agents, seed, mal_interval):
    """Hand-made 7x10 map with dead-end stations (flatland/utils/simple_rail.py:9-56): the only
    reference map type that exercises the dead-end branches (sparse_rail_generator makes none)."""
    ref = rh.load()
    from flatland.utils.simple_rail import make_simple_rail
    rail, rail_map, optionals = make_simple_rail()
    env = ref["RailEnv"](
        width=rail_map.shape[1], height=rail_map.shape[0],
        rail_generator=ref["rail_from_grid_transition_map"](rail, optionals),
        line_generator=ref["SparseLineGen"](speed_ratio_map=dict(rh.SPEED_RATIOS)),
        number_of_agents=n_agents,
        malfunction_generator=ref["ParamMalfunctionGen"](ref["MalfunctionParameters"](
            malfunction_rate=1.0 / mal_interval, min_duration=3, max_duration=8)),
        obs_builder_object=ref["TreeCutils"](31, 500), random_seed=seed)
    return env


def motion_cases(n_cases, seed, stack_p=0.0):
    """Random MotionCheck scenarios (agent_chains.py:19-37,151-236): dense little grids with
    off-map entrants, stoppers, swaps, chains and cycles; records the reference's verdict."""
    ref = rh.load()
    rng = np.random.RandomState(seed)
    MAXN = 12
    cur = np.full((n_cases, MAXN, 2), -9, np.int16)
    nxt = np.full((n_cases, MAXN, 2), -9, np.int16)
    n_ag = np.zeros(n_cases, np.int32)
    ok = np.zeros((n_cases, MAXN), np.uint8)
    for k in range(n_cases):
        h, w = rng.randint(2, 5), rng.randint(2, 5)
        n = rng.randint(1, min(MAXN, h * w) + 1)
        cells = [(r, c) for r in range(h) for c in range(w)]
        rng.shuffle(cells)
        mc = ref["MotionCheck"]()
        n_ag[k] = n
        pos = []
        for i in range(n):
            p = None if rng.rand() < 0.2 else cells[i]
            if p is not None and i > 0 and rng.rand() < stack_p:
                p = cells[rng.randint(i)]  # several trains on one cell (MALFUNCTION_OFF_MAP + STOP placement)
            if p is None:
                q = None if rng.rand() < 0.3 else cells[rng.randint(len(cells))]
            elif rng.rand() < 0.3:
                q = p
            else:
                dr, dc = [(-1, 0), (0, 1), (1, 0), (0, -1)][rng.randint(4)]
                q = (p[0] + dr, p[1] + dc)
                if not (0 <= q[0] < h and 0 <= q[1] < w):
                    q = p
            pos.append(p)
            cur[k, i] = (-1, i) if p is None else p
            nxt[k, i] = (-1, i) if q is None else q
            mc.addAgent(i, p, q)
        mc.find_conflicts()
        for i in range(n):
            ok[k, i] = bool(mc.check_motion(i, pos[i]))
    return dict(cur=cur, nxt=nxt, n=n_ag, can_move=ok)


def save(name, d):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("%-28s %8.1f KB  steps=%s stacked_steps=%s" % (name, os.path.getsize(path) / 1024, d.get("n_steps"),
                                                          d.get("n_stacked_steps")))


def main():
    jobs = [
        # name, config, level, policy, action seed, mal_interval override, max_steps
        ("t00_l0_random", "Test_00", (0, 0), "random", 1, None, None),
        ("t00_l1_greedy", "Test_00", (0, 1), "greedy", 2, 60, None),
        ("t00_l2_forward", "Test_00", (0, 2), "forward", 3, None, None),
        ("t02_l0_forward", "Test_02", (2, 0), "forward", 4, 100, None),
        ("t02_l1_greedy", "Test_02", (2, 1), "greedy", 5, None, None),
        ("t03_l0_random", "Test_03", (3, 0), "random", 6, None, 220),
        ("t03_l1_greedy", "Test_03", (3, 1), "greedy", 7, 450, None),
        ("t08_l0_greedy", "Test_08", (8, 0), "greedy", 8, 720, 400),
        ("t14_l0_forward", "Test_14", (14, 0), "forward", 9, 3600, 40),
        # heavy malfunctions + uniform random actions: MALFUNCTION_OFF_MAP + STOP puts trains on occupied cells
        ("t02_l2_stacking", "Test_02", (2, 2), "random", 10, 25, None),
        ("t03_l2_stacking", "Test_03", (3, 2), "random", 11, 40, 300),
    ]
    only = sys.argv[1:]
    for name, cfg, (test, level), policy, aseed, mal, max_steps in jobs:
        if only and name not in only:
            continue
        env = rh.make_env(cfg, rh.csv_seed(test, level), mal_interval=mal)
        save(name, run_episode(env, policy, aseed, max_steps=max_steps))
    for n_agents, seed in ((1, 11), (2, 12), (4, 13)):
        name = "simple_rail_n%d" % n_agents
        if only and name not in only:
            continue
        env = simple_rail_env(n_agents, seed, 30)
        save(name, run_episode(env, "greedy" if n_agents < 4 else "random", 20 + n_agents))
    if not only or "motion_cases" in only:
        d = motion_cases(3000, 99)
        d2 = motion_cases(3000, 100, stack_p=0.35)
        save("motion_cases", {k: np.concatenate([d[k], d2[k]]) for k in d})


if __name__ == "__main__":
    main()
