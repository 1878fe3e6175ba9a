"""Reference harness — TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference (flatland-rl Python simulator from /root/reference and the
flatland_cutils C++ observation builder compiled by oracle/Makefile into oracle/_ref/) so that
golden vectors can be generated and the C restatement (oracle/flatland_oracle.c) can be pinned
against the real thing.  /root/reference only exists in the build container: everything here
raises ``ReferenceUnavailable`` elsewhere, and nothing in the product package imports this file.

What is patched, and why (nothing under /root/reference is modified):
  * eight import-time stub modules (oracle/stubs/) for dependencies absent from this image;
  * ``TrainState.__str__`` is restored to the pre-3.11 ``"TrainState.NAME"`` form, because
    flatland_cutils/src/loader.cpp:10 parses ``str(agent.state)`` through the string table in
    tool.h:219-228 and would otherwise read every agent as WAITING on Python >= 3.11.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("FLATLAND_REFERENCE_ROOT", "/root/reference")


class ReferenceUnavailable(RuntimeError):
    pass


_loaded = {}


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "flatland-rl", "flatland"))


def load():
    """Returns a namespace dict with the reference classes; imports them on first use."""
    if _loaded:
        return _loaded
    if not available():
        raise ReferenceUnavailable("reference tree not found at %s" % REFERENCE_ROOT)
    for p in (os.path.join(HERE, "_ref"), os.path.join(REFERENCE_ROOT, "flatland-rl"),
              os.path.join(HERE, "stubs")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore")
    from flatland.envs.step_utils.states import TrainState
    TrainState.__str__ = lambda s: "TrainState." + s.name
    from flatland.envs.rail_env import RailEnv
    from flatland.envs.rail_generators import SparseRailGen, rail_from_grid_transition_map
    from flatland.envs.line_generators import SparseLineGen
    from flatland.envs.malfunction_generators import MalfunctionParameters, ParamMalfunctionGen
    from flatland.envs.agent_chains import MotionCheck
    try:
        import flatland_cutils
    except ImportError as e:  # pragma: no cover
        raise ReferenceUnavailable("oracle/_ref/flatland_cutils not built: run `make -C oracle ref` (%s)" % e)
    _loaded.update(dict(TrainState=TrainState, RailEnv=RailEnv, SparseRailGen=SparseRailGen,
                        SparseLineGen=SparseLineGen, MalfunctionParameters=MalfunctionParameters,
                        ParamMalfunctionGen=ParamMalfunctionGen, MotionCheck=MotionCheck,
                        rail_from_grid_transition_map=rail_from_grid_transition_map,
                        TreeCutils=flatland_cutils.TreeObsForRailEnv))
    return _loaded


# Flatland-3 round-2 configurations named by BASELINE.json (source of truth:
# solution/debug-environments/parameters_flatland_round_2_new.csv; width=x_dim, height=y_dim).
CONFIGS = {
    "Test_00": dict(n_agents=7, width=30, height=30, n_cities=2, mal_interval=540),
    "Test_02": dict(n_agents=20, width=30, height=30, n_cities=3, mal_interval=1800),
    "Test_03": dict(n_agents=50, width=30, height=35, n_cities=3, mal_interval=4500),
    "Test_08": dict(n_agents=80, width=60, height=60, n_cities=17, mal_interval=7200),
    "Test_14": dict(n_agents=425, width=158, height=158, n_cities=41, mal_interval=36000),
}
SPEED_RATIOS = {1.0: 0.25, 0.5: 0.25, 0.33: 0.25, 0.25: 0.25}


def csv_seed(test, level):
    """Seed of Test_<test>/Level_<level> from the reference's config table."""
    import csv
    path = os.path.join(REFERENCE_ROOT, "solution", "debug-environments",
                        "parameters_flatland_round_2_new.csv")
    with open(path) as f:
        for row in csv.DictReader(f):
            if row["test_id"] == "Test_%d" % test and row["env_id"] == "Level_%d" % level:
                return int(row["random_seed"])
    raise KeyError((test, level))


def make_env(cfg, seed, mal_interval=None, max_nodes=31, max_pred_depth=500):
    """Builds a reference RailEnv the way solution/debug-environments/generate_test_cases.py:48-62
    and solution/demo.py:20-42 do, with the flatland_cutils tree observation attached."""
    ref = load()
    c = dict(CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg)
    if mal_interval is not None:
        c["mal_interval"] = mal_interval
    rate = 0.0 if not c["mal_interval"] else 1.0 / c["mal_interval"]
    env = ref["RailEnv"](
        width=c["width"], height=c["height"],
        rail_generator=ref["SparseRailGen"](max_num_cities=c["n_cities"], grid_mode=False,
                                            max_rails_between_cities=2, max_rail_pairs_in_city=2),
        line_generator=ref["SparseLineGen"](speed_ratio_map=dict(SPEED_RATIOS)),
        number_of_agents=c["n_agents"],
        malfunction_generator=ref["ParamMalfunctionGen"](
            ref["MalfunctionParameters"](malfunction_rate=rate, min_duration=20, max_duration=50)),
        obs_builder_object=ref["TreeCutils"](max_nodes, max_pred_depth),
        random_seed=seed,
    )
    return env
