"""Multi-GPU plumbing: environments are independent, so a batch shards by environment index with no
data-path collective.  The only exchange is the final episode-statistics reduction
(one all-reduce of four int64 — NCCL over NVLink on the GPU box, gloo in the CPU tests) and the
max-over-ranks of the timed duration."""
import torch
import torch.distributed as dist


def shard_range(n_envs_total, rank, world_size):
    """Contiguous slice [lo, hi) of environment indices owned by `rank` (strong scaling: a fixed global
    batch split over the ranks; sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    lo = n_envs_total * rank // world_size
    hi = n_envs_total * (rank + 1) // world_size
    return lo, hi


def weak_offset(envs_per_rank, rank):
    """First global environment index of `rank` under weak scaling (every rank runs envs_per_rank)."""
    return envs_per_rank * rank


def reduce_episode_stats(stats, group=None):
    """Sums [episodes, arrivals, reward_sum, agent_steps] (int64[4]) over the ranks, in place."""
    if stats.dtype != torch.int64 or stats.numel() != 4:
        raise ValueError("episode stats are int64[4]")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def max_over_ranks(value_ms, device="cpu", group=None):
    """Multi-GPU durations are the max over ranks of the device-measured time."""
    t = torch.tensor([float(value_ms)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def final_metric(stats, n_agents_per_env, max_steps=None):
    """eval_env.py:81-94 final_metric aggregated over all finished episodes: arrival ratio (the reference's predicate:
    position is None and state != READY_TO_DEPART, which also counts trains that never left), mean end-of-episode reward
    per episode and — when `max_steps` (T per environment, for per-environment stats [E, 4]) is given — the mean of
    norm_reward = 1 + total_reward / T / n_agents."""
    st = stats.reshape(-1, 4) if hasattr(stats, "reshape") else torch.as_tensor(stats).reshape(-1, 4)
    episodes, arrivals, reward_sum = (int(st[:, k].sum().item()) for k in range(3))
    if episodes == 0:
        return {"episodes": 0, "arrival_ratio": None, "mean_total_reward": None, "mean_norm_reward": None}
    out = {"episodes": episodes, "arrival_ratio": arrivals / (episodes * n_agents_per_env),
           "mean_total_reward": reward_sum / episodes, "mean_norm_reward": None}
    if max_steps is not None:
        T = torch.as_tensor(max_steps, dtype=torch.float64, device=st.device).reshape(-1)
        norm = st[:, 0].double() + st[:, 2].double() / T / n_agents_per_env
        out["mean_norm_reward"] = float(norm.sum().item()) / episodes
    return out
