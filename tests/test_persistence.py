"""Saved-level loader (flatland-marl_b200/persistence.py) against a level written by the unmodified reference's
RailEnvPersister (tests/golden/level_t00.pkl, made by tests/golden/make_level_golden.py)."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_level_file_loads_without_flatland():
    """The loader must not need flatland-rl: run it in a clean interpreter and make sure nothing flatland got imported."""
    code = ("import sys; sys.path.insert(0, %r); from flatland_marl_b200 import persistence as p; "
            "w = p.load_level(%r); assert not any(m == 'flatland' or m.startswith('flatland.') for m in sys.modules); "
            "print(w['N'], w['H'], w['W'], w['T'])" % (ROOT, os.path.join(GOLD, "level_t00.pkl")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["7", "30", "30", "164"] or len(out.stdout.split()) == 4


def test_level_matches_reference_environment():
    from flatland_marl_b200 import persistence as p
    w = p.load_level(os.path.join(GOLD, "level_t00.pkl"), sched_rows=16, malfunction_seed=3)
    with np.load(os.path.join(GOLD, "level_t00_expected.npz")) as z:
        want = {k: z[k] for k in z.files}
    for k in ("H", "W", "N", "T"):
        assert int(w[k]) == int(want[k]), k
    for k in ("grid", "init_pos", "init_dir", "target", "speed", "earliest", "latest"):
        assert w[k].dtype == want[k].dtype and (w[k] == want[k]).all(), k
    assert w["dist_f64"].shape == want["dist_f64"].shape
    assert ((w["dist_f64"] == want["dist_f64"]) | (np.isinf(w["dist_f64"]) & np.isinf(want["dist_f64"]))).all()
    assert w["mal_params"][0] == float(want["mal_rate"])
    assert w["sched"].shape == (16, int(w["N"]))
    # same draws as ParamMalfunctionGen from the same RandomState (worlds.draw_schedule)
    from flatland_marl_b200.worlds import draw_schedule
    assert (w["sched"] == draw_schedule(np.random.RandomState(3), 16, int(w["N"]), w["mal_params"][0], 20, 50)).all()


def test_unpickler_refuses_foreign_globals(tmp_path):
    from flatland_marl_b200 import persistence as p
    bad = tmp_path / "evil.pkl"
    bad.write_bytes(pickle.dumps({"grid": [[0]], "agents": [], "x": os.system}))
    with pytest.raises(pickle.UnpicklingError):
        p.load_env_dict(str(bad))
    with pytest.raises(ValueError):
        p.load_env_dict(str(tmp_path / "level.mpk"))
    # numpy is not a free pass: only the ndarray / dtype / scalar reconstructors are let through, not helpers that run code
    class Gadget:
        def __reduce__(self):
            import numpy.testing._private.utils as u
            return (u.runstring, ("import os; os.environ['FL_PWNED'] = '1'", {}))
    bad.write_bytes(pickle.dumps({"grid": [[0]], "agents": [], "x": Gadget()}))
    with pytest.raises(pickle.UnpicklingError):
        p.load_env_dict(str(bad))
    assert "FL_PWNED" not in os.environ
    for mod, name in (("numpy", "load"), ("numpy.lib.npyio", "load"), ("numpy.f2py", "run_main"), ("numpy.core.multiarray", "frombuffer")):
        with pytest.raises(pickle.UnpicklingError):
            p._LevelUnpickler(__import__("io").BytesIO(b"")).find_class(mod, name)
    # what real level files need still loads
    ok = tmp_path / "ok.pkl"
    ok.write_bytes(pickle.dumps({"grid": np.zeros((2, 2), np.uint16), "agents": [], "n": np.int64(3), "f": np.float64(0.5)}))
    d = p.load_env_dict(str(ok))
    assert d["grid"].shape == (2, 2) and int(d["n"]) == 3


@pytest.mark.gpu
def test_loaded_level_runs_and_distance_map_matches_the_file():
    """The level's saved distance map (reference DistanceMap, float64, one map per agent) equals the BFS kernel's."""
    import torch
    import flatland_marl_b200 as fb
    from flatland_marl_b200 import persistence as p
    w = p.load_level(os.path.join(GOLD, "level_t00.pkl"))
    batch = fb.BatchedRailEnv([w, w])
    obs = batch.reset()
    _, slot = fb.unique_target_slots(w)
    got = batch.dist_numpy(1)
    for i in range(int(w["N"])):
        ref = np.where(np.isinf(w["dist_f64"][i]), 65535, w["dist_f64"][i]).astype(np.uint16)
        assert (got[slot[i]] == ref).all(), i
    for t in range(20):
        obs, rew, don = batch.step(torch.full((2, int(w["N"])), 2, dtype=torch.uint8, device=batch.device))
    assert int(batch.t["elapsed"][0].item()) == 20


@pytest.mark.reference
def test_reference_cannot_read_mpk(tmp_path):
    """Why .mpk level files are refused: the unmodified reference can neither read nor write them with the msgpack its own
    requirements pin (>= 1.0): load_env_dict passes `encoding=` to msgpack.unpackb (persistence.py:142), which msgpack 1.x
    rejects, and save() hands msgpack objects it has no encoder for."""
    import msgpack
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("needs the reference tree (build container)")
    rh.load()
    from flatland.envs.persistence import RailEnvPersister
    assert msgpack.version >= (1, 0, 0)
    f = tmp_path / "level.mpk"
    f.write_bytes(msgpack.packb({"grid": [[0]], "agents": []}))
    with pytest.raises(TypeError):
        RailEnvPersister.load_env_dict(str(f))
    env = rh.make_env("Test_00", rh.csv_seed(0, 0))
    env.reset()
    with pytest.raises(TypeError):
        RailEnvPersister.save(env, str(tmp_path / "out.mpk"))


@pytest.mark.gpu
def test_load_new_facade_runs_the_demo_calls():
    """solution/demo.py:89-113 with --env: RailEnvPersister.load_new, obs_builder assignment, reset, steps."""
    import flatland_marl_b200 as fb
    env, env_dict = fb.RailEnvPersister.load_new(os.path.join(GOLD, "level_t00.pkl"))
    env.obs_builder = fb.TreeObsForRailEnv(31, 500)
    assert set(env_dict) >= {"grid", "agents", "malfunction", "max_episode_steps"}
    obs, info = env.reset()
    n = env.get_num_agents()
    assert n == len(env_dict["agents"]) and env._max_episode_steps == int(env_dict["max_episode_steps"])
    for t in range(10):
        obs, rew, done, info = env.step({i: 2 for i in range(n)})
    assert np.array(obs[0]).shape == (n, 83) and not done["__all__"]
