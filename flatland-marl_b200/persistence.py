"""Loader for the reference's saved levels: `RailEnvPersister.save(env, "Level_x.pkl",
save_distance_maps=True)` files as written by solution/debug-environments/generate_test_cases.py:64-68
(flatland-rl/flatland/envs/persistence.py:21-66, 196-217) -> a world dict (worlds.py) ready for `BatchedRailEnv`.

The pickle references the reference's own classes (`flatland.envs.agent_utils.Agent`, `SpeedCounter`,
`TrainStateMachine`, ...).  flatland-rl need not be installed: a restricted unpickler maps every `flatland.*` global
to an inert stand-in that only records the pickled state, and refuses every other global except an exact list of
(module, name) pairs: the ndarray / dtype / scalar reconstructors of numpy and the builtin containers — a level file is
data, not code.

What a level file holds, and what it does not (persistence.py:196-217): grid, the agents with their timetable
(earliest_departure / latest_arrival), `max_episode_steps`, the malfunction *parameters* (rate, min, max) and optionally
the distance map.  The malfunction draws themselves come from the environment's generator at run time, so the schedule
is drawn here from `malfunction_seed` with the reference's formula (worlds.draw_schedule).
"""
import io
import pickle
from collections import namedtuple

import numpy as np

from .worlds import draw_schedule

# field order of flatland.envs.agent_utils.Agent (agent_utils.py:18-34)
Agent = namedtuple("Agent", ["initial_position", "initial_direction", "direction", "target", "moving", "earliest_departure",
                             "latest_arrival", "handle", "position", "arrival_time", "old_direction", "old_position",
                             "speed_counter", "action_saver", "state_machine", "malfunction_handler"])
# flatland.envs.malfunction_generators.MalfunctionProcessData / MalfunctionParameters
MalfunctionParameters = namedtuple("MalfunctionParameters", ["malfunction_rate", "min_duration", "max_duration"])


class _Inert:
    """Stand-in for a pickled flatland object: keeps the state dict, runs no flatland code."""

    def __init__(self, *args, **kwargs):
        self._args = args

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):   # slots form
            state = {**(state[0] or {}), **state[1]}
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


def _inert_enum(value):
    return int(value)


_SAFE_BUILTINS = {"list", "dict", "tuple", "set", "frozenset", "int", "float", "bool", "str", "bytes", "complex", "slice", "range"}


# exactly what a pickled ndarray / numpy scalar needs (numpy < 2 writes numpy.core.*, numpy >= 2 numpy._core.*); everything
# else under numpy is refused: numpy.testing / numpy.f2py / numpy.load hold helpers that run code
_SAFE_NUMPY = {(m, n) for m in ("numpy.core.multiarray", "numpy._core.multiarray") for n in ("_reconstruct", "scalar")} | \
    {("numpy", "ndarray"), ("numpy", "dtype")} | \
    {("numpy", n) for n in ("bool_", "int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64", "float16",
                            "float32", "float64", "intc", "uintc", "longlong", "ulonglong")}


class _LevelUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("flatland."):
            if name == "Agent":
                return Agent
            if name in ("MalfunctionProcessData", "MalfunctionParameters"):
                return MalfunctionParameters
            if name in ("Grid4TransitionsEnum", "TrainState", "RailEnvActions"):
                return _inert_enum
            return type(name, (_Inert,), {})
        if (module, name) in _SAFE_NUMPY:
            return super().find_class(module, name)
        if module in ("builtins", "__builtin__") and name in _SAFE_BUILTINS:
            return super().find_class(module, name)
        if module == "collections" and name == "OrderedDict":
            return super().find_class(module, name)
        raise pickle.UnpicklingError("level file references %s.%s: refused" % (module, name))


def load_env_dict(path):
    """RailEnvPersister.load_env_dict (persistence.py:132-161) for .pkl files, without importing flatland.

    .mpk (msgpack) files are refused: the reference itself can neither write nor read them with its own pinned
    dependencies — `RailEnvPersister.save` packs the env dict with `msgpack.packb`, which has no encoder for the
    SpeedCounter / TrainStateMachine objects inside the Agent tuples, and `load_env_dict` calls
    `msgpack.unpackb(..., encoding="utf-8")`, a keyword msgpack >= 1.0 (requirements_dev.txt:11 pins >=1.0,<2) no longer
    accepts (tests/test_persistence.py::test_reference_cannot_read_mpk shows both on the unmodified reference)."""
    if not str(path).endswith(".pkl"):
        raise ValueError("only .pkl level files are supported (%s); see load_env_dict's docstring for .mpk" % path)
    with open(path, "rb") as f:
        d = _LevelUnpickler(io.BytesIO(f.read())).load()
    if not isinstance(d, dict) or "grid" not in d or "agents" not in d:
        raise ValueError("%s is not a RailEnvPersister level file" % path)
    return d


def _speed_of(speed_counter):
    s = getattr(speed_counter, "_speed", None)
    if s is None:
        s = getattr(speed_counter, "speed", None)
    if s is None:
        raise ValueError("level file: agent without a speed")
    return float(s)


def world_from_env_dict(d, sched_rows=None, malfunction_seed=0):
    """env_dict -> world dict (keys of worlds.WORLD_KEYS + sched, mal_params and, when saved, dist_f64)."""
    grid = np.asarray(d["grid"], dtype=np.uint16)
    ags = d["agents"]
    if any(not isinstance(a, Agent) for a in ags):
        raise ValueError("level file: legacy agent tuples are not supported")
    H, W = grid.shape
    T = d.get("max_episode_steps")
    if T is None:
        raise ValueError("level file has no max_episode_steps")
    w = dict(
        H=int(H), W=int(W), N=len(ags), T=int(T), grid=grid,
        init_pos=np.array([a.initial_position for a in ags], dtype=np.int16).reshape(-1, 2),
        init_dir=np.array([int(a.initial_direction) for a in ags], dtype=np.uint8),
        target=np.array([a.target for a in ags], dtype=np.int16).reshape(-1, 2),
        speed=np.array([_speed_of(a.speed_counter) for a in ags], dtype=np.float64),
        earliest=np.array([a.earliest_departure for a in ags], dtype=np.int32),
        latest=np.array([a.latest_arrival for a in ags], dtype=np.int32),
    )
    mp = d.get("malfunction")
    mp = MalfunctionParameters(*mp) if mp is not None else MalfunctionParameters(0.0, 0, 0)
    w["mal_params"] = np.array([float(mp.malfunction_rate), float(mp.min_duration), float(mp.max_duration)])
    rows = int(sched_rows or w["T"])
    w["sched"] = draw_schedule(np.random.RandomState(malfunction_seed), rows, w["N"], float(mp.malfunction_rate),
                               int(mp.min_duration), int(mp.max_duration))
    if d.get("distance_map") is not None:
        w["dist_f64"] = np.asarray(d["distance_map"], dtype=np.float64)     # [N, H, W, 4], inf = unreachable
    return w


def load_level(path, sched_rows=None, malfunction_seed=0):
    """One saved level -> world dict."""
    return world_from_env_dict(load_env_dict(path), sched_rows=sched_rows, malfunction_seed=malfunction_seed)


class RailEnvPersister:
    """`flatland.envs.persistence.RailEnvPersister` for the hot path's consumers (solution/demo.py:89:
    `env, _ = RailEnvPersister.load_new(args.env); env.obs_builder = TreeCutils(...)`): the level file becomes a
    `flatland_marl_b200.RailEnv` on the GPU.  `load_new` returns (env, env_dict) like the reference
    (persistence.py:105-129).  The reference's loaded env regenerates its timetable on the next default `reset()` from an
    unseeded RandomState (rail_env.py:260-330 with `rail_from_file` / `line_from_file`), i.e. differently on every run;
    this env keeps the timetable stored in the file."""

    @classmethod
    def load_new(cls, filename, load_from_package=None, device="cuda:0", malfunction_seed=0):
        if load_from_package is not None:
            raise ValueError("load_from_package is not supported: pass a file path")
        from .rail_env import RailEnv
        env_dict = load_env_dict(filename)
        world = world_from_env_dict(env_dict, malfunction_seed=malfunction_seed)
        return RailEnv.from_world(world, device=device), env_dict

    @classmethod
    def load_env_dict(cls, filename, load_from_package=None):
        return load_env_dict(filename)
