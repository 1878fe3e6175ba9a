"""Times the policy forward pass (tcgen05 kernels) and the on-device rollout loop observation -> policy ->
action -> step on one GPU.  usage: python tools/policy_bench.py [config] [envs] [steps]
Prints one JSON line; with FL_POLICY_EVENTS=1 also a per-phase breakdown from CUDA events."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import flatland_marl_b200 as fb  # noqa: E402
from flatland_marl_b200.policy import BatchedActor  # noqa: E402

# multiply-accumulates per agent of Network.forward with n_real real nodes of which n_inner have children
def macs_per_agent(n_nodes, n_inner, n_agents):
    tree = n_nodes * 12 * 384 + n_inner * (384 * 384 + 384 * 128 + 3 * (128 * 128 + 12 * 128))
    mlp = 83 * 256 + 2 * 256 * 256 + 256 * 128
    tf = 3 * (256 * 768 + 256 * 256 + 512 * 256 + 2 * n_agents * 256)
    head = 512 * 512 + 2 * 256 * 128 + 6 * 128
    return tree + mlp + tf + head


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else "Test_03"
    E = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[config]["envs"]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    worlds = bench.load_worlds(config, E)
    env = fb.BatchedRailEnv(worlds, auto_reset=True)
    actor = BatchedActor(None, seed=0)
    N = env.N
    obs = env.reset()
    rng = np.random.RandomState(3)
    for _ in range(100):
        a = torch.from_numpy(rng.randint(0, 5, size=(E, N)).astype(np.uint8)).to(env.device)
        obs, _, _ = env.step(a)
    torch.cuda.synchronize()
    no = obs["node_order"]
    n_nodes = float((no >= 0).sum().item()) / (E * N)
    n_inner = float((no >= 1).sum().item()) / (E * N)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.device)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for _ in range(5):
        actor.get_actions(obs)
    torch.cuda.synchronize()
    for s, e in ev:
        flush.zero_()
        s.record()
        actor.get_actions(obs)
        e.record()
    torch.cuda.synchronize()
    fwd_ms = float(np.median([s.elapsed_time(e) for s, e in ev]))
    # rollout loop
    for _ in range(5):
        obs, _, _ = env.step(actor.get_actions(obs))
    torch.cuda.synchronize()
    for s, e in ev:
        flush.zero_()
        s.record()
        acts = actor.get_actions(obs)
        obs, rew, don = env.step(acts)
        e.record()
    torch.cuda.synchronize()
    loop_ms = float(np.median([s.elapsed_time(e) for s, e in ev]))
    macs = macs_per_agent(n_nodes, n_inner, N)
    out = {"config": config, "envs": E, "agents": N, "nodes_per_tree": n_nodes, "inner_nodes_per_tree": n_inner,
           "policy_forward_ms": fwd_ms, "policy_agent_steps_per_s": E * N / fwd_ms * 1e3,
           "policy_tflops": 2 * macs * E * N / fwd_ms / 1e9, "mmacs_per_agent": macs / 1e6,
           "rollout_ms_per_step": loop_ms, "rollout_agent_steps_per_s": E * N / loop_ms * 1e3,
           "policy_launches_per_forward": None}
    c0 = actor.launch_count()
    actor.get_actions(obs)
    out["policy_launches_per_forward"] = actor.launch_count() - c0
    print(json.dumps(out))


if __name__ == "__main__":
    main()
