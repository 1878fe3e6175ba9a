"""Reset pipeline: world sources on the CPU; on the GPU, swapping new worlds into slots of a running
batch must leave those slots bit-identical to a fresh batch of the new worlds and must not disturb the others."""
import os

import numpy as np
import pytest

from flatland_marl_b200.reset_pipeline import GeneratorPool, PackSource


def _toy_world(seed):
    return {"seed": int(seed), "grid": np.full((2, 2), seed, np.uint16)}


def _failing(seed):
    raise ValueError("no world for %d" % seed)


def test_pack_source_cycles():
    src = PackSource([{"k": i} for i in range(3)], start=2)
    assert [w["k"] for w in src.take(4)] == [2, 0, 1, 2]
    assert [w["k"] for w in src.take(1)] == [0]
    with pytest.raises(ValueError):
        PackSource([])


def test_generator_pool_runs_ahead_and_reports_failures():
    pool = GeneratorPool(_toy_world, first_seed=10, n_workers=2, ahead=4)
    try:
        got = []
        for _ in range(50):
            got += pool.take(6 - len(got), timeout=0.5)
            if len(got) >= 6:
                break
        assert len(got) == 6
        seeds = [w["seed"] for w in got]                  # completion order: any order, but no seed twice and none skipped far
        assert len(set(seeds)) == 6 and min(seeds) == 10 and max(seeds) < 10 + 6 + 4
        assert all((w["grid"] == w["seed"]).all() for w in got)
    finally:
        pool.close()
    bad = GeneratorPool(_failing, n_workers=1, ahead=1)
    try:
        with pytest.raises(RuntimeError):
            for _ in range(50):
                bad.take(1, timeout=0.5)
    finally:
        bad.close()


@pytest.mark.gpu
def test_replace_worlds_equals_fresh_batch(golden):
    import torch
    import bench
    import flatland_marl_b200 as fb
    worlds = bench.load_worlds("Test_03", 12)
    batch = fb.BatchedRailEnv(worlds[:8], reserve=0.5, min_slots=8)
    control = fb.BatchedRailEnv(worlds[:8], reserve=0.5, min_slots=8)       # never touched
    batch.reset(); control.reset()
    rng = np.random.RandomState(11)
    N = batch.N

    def actions(n):
        return np.where(rng.rand(n, N) < 0.7, 2, rng.randint(0, 5, (n, N))).astype(np.uint8)

    for t in range(25):
        a = torch.from_numpy(actions(8)).to(batch.device)
        batch.step(a); control.step(a)
    # slots 2 and 5 get worlds 8 and 9 in mid-run
    batch.replace_worlds([5, 2], [worlds[9], worlds[8]])
    fresh = fb.BatchedRailEnv([worlds[8], worlds[9]], reserve=0.5, min_slots=8)
    obs, fobs, cobs = batch.observe(), fresh.reset(), control.observe()
    for k in obs:
        assert (obs[k][2].cpu().numpy() == fobs[k][0].cpu().numpy()).all(), k
        assert (obs[k][5].cpu().numpy() == fobs[k][1].cpu().numpy()).all(), k
        for e in (0, 1, 3, 4, 6, 7):
            assert (obs[k][e].cpu().numpy() == cobs[k][e].cpu().numpy()).all(), (k, e)
    assert (batch.dist_numpy(2)[: fresh.n_slots] == fresh.dist_numpy(0)).all()
    for t in range(30):
        a8 = actions(8)
        a = torch.from_numpy(a8).to(batch.device)
        obs, rew, don = batch.step(a)
        cobs, crew, cdon = control.step(a)
        fobs, frew, fdon = fresh.step(torch.from_numpy(a8[[2, 5]]).to(batch.device))
        for k in obs:
            assert (obs[k][2].cpu().numpy() == fobs[k][0].cpu().numpy()).all(), (t, k)
            assert (obs[k][5].cpu().numpy() == fobs[k][1].cpu().numpy()).all(), (t, k)
            assert (obs[k][7].cpu().numpy() == cobs[k][7].cpu().numpy()).all(), (t, k)
        assert (rew[2].cpu().numpy() == frew[0].cpu().numpy()).all() and (don[5].cpu().numpy() == fdon[1].cpu().numpy()).all()
        s, fs = batch.state_numpy(2), fresh.state_numpy(0)
        for k in ("pos", "dir", "state", "ctr", "mal"):
            assert (s[k] == fs[k]).all(), (t, k)
    with pytest.raises(ValueError):
        batch.replace_worlds([1, 1], [worlds[0], worlds[1]])


@pytest.mark.gpu
def test_pipeline_replaces_finished_environments():
    """Every slot of the pipeline against the oracle of the world it currently holds: an episode that ends in step t
    is followed in step t+1 by the first observation of the NEXT world of the source (no host synchronisation on the
    stepping path: the done flags are read one step late), and that world then runs bit-exactly."""
    import torch
    import bench
    import flatland_marl_b200 as fb
    from oracle import oracle as orc
    from test_gpu_parity import compare_obs, compare_state
    worlds = bench.load_worlds("Test_00", 16)
    for w in worlds:
        w["T"] = 40                                  # short episodes so that resets happen within the test
    batch = fb.BatchedRailEnv(worlds[:6], auto_reset=True, reserve=1.0, min_slots=8)
    src = fb.PackSource(worlds[6:])
    pipe = fb.ResetPipeline(batch, src)
    obs = pipe.reset()
    envs = [orc.OracleEnv(w) for w in worlds[:6]]
    held = list(worlds[:6])
    rows = [0] * 6                                   # schedule row of each slot's current world
    for e in envs:
        e.reset()
    nxt = 0                                          # next world of the source (PackSource hands them out in order)
    rng = np.random.RandomState(2)
    episodes, pending = 0, []
    for t in range(130):
        act = rng.randint(0, 5, (6, batch.N)).astype(np.uint8)
        obs, rew, don = pipe.step(torch.from_numpy(act).to(batch.device))
        don = don.cpu().numpy()
        for e in pending:                            # finished last step: this call handed the slot the next world
            held[e] = worlds[6 + nxt % 10]
            nxt += 1
            envs[e] = orc.OracleEnv(held[e])
            envs[e].reset()
            rows[e] = 0
            compare_obs(batch, e, envs[e].obs(), "slot %d step %d: first observation of the new world" % (e, t))
            compare_state(batch, e, envs[e].state(), "slot %d step %d new world" % (e, t))
        fresh = set(pending)
        pending = []
        for e in range(6):
            if e in fresh:
                continue
            orew, odon = envs[e].step(act[e], held[e]["sched"][rows[e]])
            rows[e] += 1
            assert (rew[e].cpu().numpy() == orew).all() and (don[e] == odon).all(), (t, e)
            if t % 7 == 0 or odon[-1]:
                compare_obs(batch, e, envs[e].obs(), "slot %d step %d" % (e, t))
            if odon[-1]:
                episodes += 1
                pending.append(e)
        assert int(batch.t["status"].max().item()) & 5 == 0      # no step-after-done, no bad cell
    assert episodes >= 12 and pipe.replaced == episodes - len(pending)
    assert np.isfinite(obs["agent_attr"].cpu().numpy()).all()


@pytest.mark.gpu
def test_replace_worlds_refuses_oversized_world_and_leaves_the_batch_untouched():
    import torch
    import bench
    import flatland_marl_b200 as fb
    small = bench.load_worlds("Test_00", 4)
    batch = fb.BatchedRailEnv(small[:3], min_slots=8)              # no reserve: tables sized for these three worlds
    batch.reset()
    a = torch.full((3, batch.N), 2, dtype=torch.uint8, device=batch.device)
    for _ in range(5):
        batch.step(a)
    before = {k: v.clone() for k, v in batch.t.items()}
    big = dict(small[3])
    g = np.array(big["grid"], np.uint16).copy()
    g[g == 0] = 0x8421                                             # every empty cell becomes a crossing: far more rail states
    big["grid"] = g
    with pytest.raises(fb.FlatlandB200Error):
        batch.replace_worlds([1], [big])
    for k, v in before.items():
        assert torch.equal(batch.t[k], v), k + " changed although the replacement was refused"
    batch.step(a)                                                  # and the batch still steps
