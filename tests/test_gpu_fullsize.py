"""Parity at BASELINE.json's full batch sizes.  The oracle cannot step thousands of environments in
seconds, so each full-size batch is checked three ways: (1) a sample of its environments is stepped by
the oracle with the same worlds, actions and schedules and compared bit-exactly every few steps;
(2) size-independent properties: an environment's results do not depend on which batch, slice or chunk it
is computed in (same bytes from the full batch, from a batch holding only that environment, and from the
chunked host-buffer entry point), and the agent-step counter equals E*N*steps; (3) the sticky status
word stays clean (no bad cell, no step-after-done)."""
import os

import numpy as np
import pytest

from test_gpu_parity import assert_same, compare_obs, compare_state

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [  # config, envs, steps, sampled envs, oracle check every
    ("Test_03", 1024, 60, (0, 517, 1023), 5),
    ("Test_08", 512, 40, (3, 511), 8),
    ("Test_14", 64, 24, (63,), 8),
    ("Test_02", 8192, 50, (1, 4097, 8191), 5),
]


def _load(config, n):
    import bench
    return bench.load_worlds(config, n)


@pytest.mark.parametrize("config,n_envs,n_steps,sample,every", CASES)
def test_full_batch_matches_oracle_on_sampled_envs(config, n_envs, n_steps, sample, every):
    import torch
    import flatland_marl_b200 as fb
    from oracle import oracle as orc
    worlds = _load(config, n_envs)
    N = int(worlds[0]["N"])
    batch = fb.BatchedRailEnv(worlds)
    solo = fb.BatchedRailEnv([worlds[e] for e in sample])           # the same environments in a batch of their own
    envs = {e: orc.OracleEnv(worlds[e]) for e in sample}
    batch.reset(); solo.reset()
    for e, env in envs.items():
        env.reset()
        assert_same(batch.dist_numpy(e)[: env.dist_u16().shape[0]], env.dist_u16(), "%s env %d distance map" % (config, e))
        compare_obs(batch, e, env.obs(), "%s env %d reset obs" % (config, e))
    rng = np.random.RandomState(7)
    idx = torch.tensor(sample, device=batch.device)
    for t in range(n_steps):
        # mostly forward so that trains leave the stations and meet each other
        act = np.where(rng.rand(n_envs, N) < 0.7, 2, rng.randint(0, 5, (n_envs, N))).astype(np.uint8)
        a = torch.from_numpy(act).to(batch.device)
        _, rew, don = batch.step(a)
        solo.step(a[idx].contiguous())
        for j, e in enumerate(sample):
            orew, odon = envs[e].step(act[e], worlds[e]["sched"][t])
            what = "%s env %d step %d" % (config, e, t + 1)
            assert_same(rew[e].cpu().numpy(), orew, what + " rewards")
            assert_same(don[e].cpu().numpy(), odon, what + " dones")
            if t % every == 0 or t == n_steps - 1:
                compare_state(batch, e, envs[e].state(), what)
                compare_obs(batch, e, envs[e].obs(), what + " obs")
                for k in batch.obs:                                   # batch-composition independence, byte for byte
                    assert torch.equal(batch.obs[k][e], solo.obs[k][j]), what + " " + k + " differs between batch sizes"
    stats = batch.episode_stats().cpu().numpy()
    assert int(stats[3]) == n_envs * N * n_steps
    assert int(batch.t["status"].max()) == 0


def test_chunked_host_step_equals_device_step():
    """fl_step_observe_host with 1, 3 and 8 chunks (copies overlapped on a second stream) and fl_step_observe_host_compact
    (compact wire format, expanded by the library's host threads) return the same bytes as the device-resident path."""
    import torch
    import flatland_marl_b200 as fb
    worlds = _load("Test_03", 256)
    N = int(worlds[0]["N"])
    ref = fb.BatchedRailEnv(worlds)
    # (chunks, wire): the full copies, and the compact wire format packed on the device and expanded by host threads
    others = {c: fb.BatchedRailEnv(worlds) for c in ((1, "full"), (3, "full"), (8, "full"), (1, "compact"), (4, "compact"), (7, "compact"))}
    ref.reset()
    for b in others.values():
        b.reset()
    rng = np.random.RandomState(11)
    for t in range(25):
        act = np.where(rng.rand(256, N) < 0.7, 2, rng.randint(0, 5, (256, N))).astype(np.uint8)
        ref.step(torch.from_numpy(act).to(ref.device))
        for c, b in others.items():
            h = b.step_host(act, n_chunks=c[0], wire=c[1])
            if c[1] == "compact":
                assert 0 < b.last_wire_bytes < 0.6 * b.d2h_bytes_per_step
            for k in ref.obs:
                np.testing.assert_array_equal(h[k].numpy(), ref.obs[k].cpu().numpy(), err_msg="%s chunks=%s step %d" % (k, c, t))
            np.testing.assert_array_equal(h["rewards"].numpy(), ref.rewards.cpu().numpy())
            np.testing.assert_array_equal(h["dones"].numpy(), ref.dones.cpu().numpy())


@pytest.mark.parametrize("config,n_envs,n_steps", [("Test_03", 96, 260), ("Test_02", 256, 200), ("Test_08", 16, 320)])
def test_caches_do_not_change_a_byte(config, n_envs, n_steps):
    """FlBatch.tree_cache / path_cache keep per-agent intermediates of fl_observe under the agent's rail state as key.  Two
    batches of the same worlds step the same random actions through whole episodes with in-place resets, one with both caches
    on (the default), one with both off: every observation tensor, reward and done flag must be the same bytes at every step;
    then the cached batch gets new worlds in some slots (the tables change under the caches) and a mid-episode state upload."""
    import torch
    import flatland_marl_b200 as fb
    lib = fb._lib.lib()
    worlds = _load(config, n_envs)
    N = int(worlds[0]["N"])
    on, off = fb.BatchedRailEnv(worlds, auto_reset=True, reserve=0.5, min_slots=12), fb.BatchedRailEnv(worlds, auto_reset=True, reserve=0.5, min_slots=12)

    def both(fn):
        out = []
        for batch, flag in ((on, -1), (off, 0)):
            for knob in (b"treecache", b"pathcache"):
                assert lib.fl_observe_override(knob, flag) == 0
            out.append(fn(batch))
        return out

    def same(what):
        for k in on.obs:
            assert torch.equal(on.obs[k], off.obs[k]), "%s: %s differs with the caches on" % (what, k)

    try:
        both(lambda b: b.reset())
        same("reset")
        rng = np.random.RandomState(11)
        for t in range(n_steps):
            # phases of mostly-forward and uniform random actions: trains depart, move, stop and wait
            p_fwd = 0.8 if (t // 40) % 2 == 0 else 0.0
            act = torch.from_numpy(np.where(rng.rand(n_envs, N) < p_fwd, 2, rng.randint(0, 5, (n_envs, N))).astype(np.uint8)).to(on.device)
            (_, r1, d1), (_, r2, d2) = both(lambda b: b.step(act))
            assert torch.equal(r1, r2) and torch.equal(d1, d2), "step %d rewards / dones" % (t + 1)
            same("step %d" % (t + 1))
            if t == n_steps // 2:
                # new worlds in a few slots of both batches: the walk tables of those slots change under the caches
                ids = list(range(0, n_envs, max(n_envs // 5, 1)))
                new = _load(config, len(ids) + n_envs)[n_envs:]
                both(lambda b: b.replace_worlds(ids, new))
                both(lambda b: b.observe())
                same("after replace_worlds")
        assert int(on.t["status"].max()) & 4 == 0
    finally:
        for knob in (b"treecache", b"pathcache"):
            lib.fl_observe_override(knob, -1)
