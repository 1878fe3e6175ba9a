"""Multi-process path on CPU (gloo, world_size 2): environment sharding and the final episode-stat
reduction.  Each rank steps ITS shard of a small batch with the oracle standing in for the GPU (the
CUDA path itself is covered by the -m gpu tests); the reduced statistics must equal a single-process
run over the whole batch, and every environment must be owned by exactly one rank."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import flatland_marl_b200 as fb

N_ENVS, N_STEPS = 6, 120


def _episode_stats(worlds, lo, hi, n_steps):
    """[episodes, arrivals, reward_sum, agent_steps] of envs lo..hi-1 — what k_step accumulates."""
    from oracle import oracle as orc
    stats = np.zeros(4, np.int64)
    for k in range(lo, hi):
        w = worlds[k]
        env = orc.OracleEnv(w)
        env.reset()
        rng = np.random.RandomState(50 + k)
        n = int(w["N"])
        for t in range(n_steps):
            if env.done_all:
                env.reset()
            act = np.where(rng.rand(n) < 0.8, 2, rng.randint(0, 5, n)).astype(np.uint8)
            rew, don = env.step(act, np.zeros(n, np.uint8))
            stats[3] += n
            if don[-1]:
                stats[0] += 1
                stats[1] += int((env.state()["state"] == 6).sum())
                stats[2] += int(rew.sum())
    return stats


def _worlds():
    pack, _ = fb.load_worlds_npz(os.path.join(os.path.dirname(os.path.dirname(__file__)), "data", "worlds", "test_00.npz"))
    return pack[:N_ENVS]


def _worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        lo, hi = fb.shard_range(N_ENVS, rank, world_size)
        stats = torch.from_numpy(_episode_stats(_worlds(), lo, hi, N_STEPS))
        fb.reduce_episode_stats(stats)
        tmax = fb.max_over_ranks(10.0 + rank)
        owned = torch.zeros(N_ENVS, dtype=torch.int64)
        owned[lo:hi] = 1
        dist.all_reduce(owned)
        if rank == 0:
            out.put((stats.tolist(), tmax, owned.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_stat_reduction():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    stats, tmax, owned = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _episode_stats(_worlds(), 0, N_ENVS, N_STEPS)
    assert stats == want.tolist()
    assert stats[3] == N_ENVS * N_STEPS * 7
    assert tmax == 11.0
    assert owned == [1] * N_ENVS


def test_shard_ranges_partition_the_batch():
    for total in (1, 7, 64, 1024, 8192):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = fb.shard_range(total, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(total))
    with pytest.raises(ValueError):
        fb.shard_range(8, 2, 2)
    assert fb.weak_offset(1024, 3) == 3072


def test_final_metric():
    m = fb.final_metric(torch.tensor([4, 14, -80, 1000], dtype=torch.int64), 7)
    assert m["episodes"] == 4 and m["arrival_ratio"] == 0.5 and m["mean_total_reward"] == -20.0
    with pytest.raises(ValueError):
        fb.reduce_episode_stats(torch.zeros(3, dtype=torch.int64))
