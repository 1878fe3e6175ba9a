"""Generated worlds: what `RailEnv.reset` leaves behind once the reference's Python generators
(sparse_rail_generator, sparse_line_generator, timetable_generator — out of scope)
have run.  A world is a dict of numpy arrays:

  H W N T            ints (T = _max_episode_steps)
  grid               uint16 [H, W]   transition bitmasks (core/transition_map.py:144)
  init_pos target    int16  [N, 2]   (row, col)
  init_dir           uint8  [N]
  speed              float64 [N]
  earliest latest    int32  [N]
  sched              uint8  [S, N]   pre-drawn malfunction durations, 0 = no onset (optional)
"""
import copy

import numpy as np

WORLD_KEYS = ("H", "W", "N", "T", "grid", "init_pos", "init_dir", "target", "speed", "earliest", "latest")


def world_from_reference_env(env, n_sched=None):
    """Extracts a world from a reference `RailEnv` AFTER its reset() (rail_env.py:260-357), and
    pre-draws the malfunction schedule from a clone of env.np_random: the generator is called once
    per agent per step regardless of state (rail_env.py:524, malfunction_generators.py:46-53), so a
    [S, N] table reproduces the reference's draws exactly."""
    ags = env.agents
    w = dict(
        H=int(env.height), W=int(env.width), N=len(ags), T=int(env._max_episode_steps),
        grid=np.asarray(env.rail.grid, dtype=np.uint16).copy(),
        init_pos=np.array([a.initial_position for a in ags], dtype=np.int16),
        init_dir=np.array([int(a.initial_direction) for a in ags], dtype=np.uint8),
        target=np.array([a.target for a in ags], dtype=np.int16),
        speed=np.array([a.speed_counter.speed for a in ags], dtype=np.float64),
        earliest=np.array([a.earliest_departure for a in ags], dtype=np.int32),
        latest=np.array([a.latest_arrival for a in ags], dtype=np.int32),
    )
    n_sched = w["T"] if n_sched is None else n_sched
    rng = copy.deepcopy(env.np_random)
    gen = env.malfunction_generator
    sched = np.zeros((n_sched, w["N"]), dtype=np.uint8)
    for t in range(n_sched):
        for i in range(w["N"]):
            sched[t, i] = gen.generate(rng).num_broken_steps
    w["sched"] = sched
    return w


def draw_schedule(rng, n_steps, n_agents, rate, min_duration=20, max_duration=50):
    """Same draws as ParamMalfunctionGen.generate (malfunction_generators.py:25-53) from a numpy
    RandomState: rand() < 1-exp(-rate) -> randint(min, max+1)+1, one call per agent per step."""
    sched = np.zeros((n_steps, n_agents), dtype=np.uint8)
    if rate <= 0:
        return sched
    prob = 1.0 - np.exp(-rate)
    for t in range(n_steps):
        for i in range(n_agents):
            if rng.rand() < prob:
                sched[t, i] = rng.randint(min_duration, max_duration + 1) + 1
    return sched


def draw_schedule_fast(rng, n_steps, n_agents, rate, min_duration=20, max_duration=50):
    """Vectorised synthetic schedule with the same distribution (NOT the same stream) — for
    benchmark batches only, where no reference run is being mirrored."""
    sched = np.zeros((n_steps, n_agents), dtype=np.uint8)
    if rate > 0:
        hit = rng.rand(n_steps, n_agents) < (1.0 - np.exp(-rate))
        sched[hit] = rng.randint(min_duration, max_duration + 1, size=int(hit.sum())) + 1
    return sched


def unique_target_slots(world):
    """Unique targets in first-appearance order, as DistanceMap._compute dedupes them
    (distance_map.py:71-79).  Returns (slot_rc [n,2] int16, slot [N] uint16)."""
    seen, slot = [], []
    for t in world["target"]:
        key = (int(t[0]), int(t[1]))
        if key not in seen:
            seen.append(key)
        slot.append(seen.index(key))
    return np.array(seen, dtype=np.int16).reshape(-1, 2), np.array(slot, dtype=np.uint16)


def max_count_of(speed):
    """SpeedCounter.max_count = int(1/speed) - 1 in float64 (speed_counter.py:40-42)."""
    return np.array([int(1 / float(s)) - 1 for s in speed], dtype=np.uint8)


def load_worlds_npz(path):
    """Loads a pack of worlds saved by save_worlds_npz (arrays stacked on a leading axis)."""
    with np.load(path) as z:
        k = int(z["n_worlds"])
        out = []
        for j in range(k):
            w = {key: (int(z[key][j]) if key in ("H", "W", "N", "T") else z[key][j]) for key in WORLD_KEYS}
            out.append(w)
        meta = {key: z[key] for key in z.files if key.startswith("meta_")}
    return out, meta


def save_worlds_npz(path, worlds, **meta):
    d = {"n_worlds": np.int32(len(worlds))}
    for key in WORLD_KEYS:
        d[key] = np.stack([np.asarray(w[key]) for w in worlds])
    for k, v in meta.items():
        d["meta_" + k] = np.asarray(v)
    np.savez_compressed(path, **d)
