"""Evaluator client shim (flatland/evaluators/client.py:228-321) against an in-process fake of the redis
server and the evaluation service (tests/fake_evaluator.py).

* build container: the fake is first pinned to the protocol by the UNMODIFIED reference client, then the shim and the
  reference client evaluate the same level under the same seed and must see the same local observations, rewards and dones
  (the shim's local environment is the façade over a test double with the C oracle as engine: no GPU here);
* GPU box: the shim drives a GPU-resident environment loaded from the level file."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from fake_evaluator import FakeRedis, FakeService

LEVEL = "level_t00.pkl"


def obs_arrays(obs):
    attr, (forest, adj, norder, eorder) = obs
    f = np.array(forest, np.float32)
    return [np.array(attr, np.float32), f, np.array(adj), np.array(norder), np.array(eorder)]


@pytest.mark.reference
def test_shim_matches_reference_client_on_the_same_level():
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("needs the reference tree (build container)")
    ref = rh.load()
    import flatland.evaluators.client as rc
    import flatland_marl_b200 as fb
    from test_dropin import OracleBatch

    def run(make_client, make_obs):
        r = FakeRedis()
        svc = FakeService(r, [LEVEL], [1234])
        svc.start()
        client = make_client(r)
        obs, info = client.env_create(make_obs())
        out = [obs_arrays(obs)]
        n = client.env.get_num_agents()
        rng = np.random.RandomState(5)
        for t in range(60):
            act = {i: int(rng.randint(0, 5)) for i in range(n) if rng.rand() < 0.9}
            obs, rew, done, info = client.env_step(act)
            out.append(obs_arrays(obs) + [np.array([rew[i] for i in range(n)]), np.array([done[i] for i in range(n)] + [done["__all__"]])])
            if done["__all__"]:
                break
        assert client.env_create(make_obs())[0] is False          # no more levels
        client.submit()
        svc.join(timeout=5)
        return out, svc, int(client.env._max_episode_steps)

    def ref_client(r):                                            # the unmodified reference client on the fake redis
        c = rc.FlatlandRemoteClient.__new__(rc.FlatlandRemoteClient)
        c.use_pickle, c.verbose, c.redis_conn = False, False, r
        c.namespace, c.service_id = "flatland-rl", "T12345"
        c.command_channel, c.error_channel = "flatland-rl::T12345::commands", "flatland-rl::T12345::errors"
        c.test_envs_root, c.current_env_path, c.env, c.stats, c.env_step_times = GOLDEN_DIR, None, None, {}, []
        c.ping_pong()
        return c

    def shim_client(r):
        def factory(world, obs_builder):
            return fb.RailEnv(width=int(world["W"]), height=int(world["H"]), number_of_agents=int(world["N"]), world=world,
                              batch=OracleBatch(world), index=0, obs_builder_object=obs_builder)
        return fb.FlatlandRemoteClient(test_env_folder=GOLDEN_DIR, redis_conn=r, env_factory=factory)

    want, svc_a, T_a = run(ref_client, lambda: ref["TreeCutils"](31, 500))
    got, svc_b, T_b = run(shim_client, lambda: fb.TreeObsForRailEnv(31, 500))
    assert T_a == T_b and svc_a.steps == svc_b.steps == len(want) - 1 and svc_a.actions == svc_b.actions
    assert len(got) == len(want)
    for t, (a, b) in enumerate(zip(got, want)):
        for k, (x, y) in enumerate(zip(a, b)):
            np.testing.assert_array_equal(x, y, err_msg="step %d item %d" % (t, k))


@pytest.mark.gpu
def test_shim_drives_a_gpu_environment():
    import flatland_marl_b200 as fb
    r = FakeRedis()
    svc = FakeService(r, [LEVEL], [77])
    svc.start()
    client = fb.FlatlandRemoteClient(test_env_folder=GOLDEN_DIR, redis_conn=r)
    obs, info = client.env_create(fb.TreeObsForRailEnv(31, 500))
    n = client.env.get_num_agents()
    direct = fb.RailEnv.from_world(fb.load_level(os.path.join(GOLDEN_DIR, LEVEL), malfunction_seed=77))
    dobs, _ = direct.reset()
    for x, y in zip(obs_arrays(obs), obs_arrays(dobs)):
        np.testing.assert_array_equal(x, y)
    rng = np.random.RandomState(1)
    for t in range(30):
        act = {i: int(rng.randint(0, 5)) for i in range(n)}
        obs, rew, done, info = client.env_step(act)
        dobs, drew, ddone, _ = direct.step(act)
        for x, y in zip(obs_arrays(obs), obs_arrays(dobs)):
            np.testing.assert_array_equal(x, y)
        assert rew == drew and done == ddone
    client.submit()
    svc.join(timeout=5)
    assert svc.steps == 30 and "file" in client.timetable_source
